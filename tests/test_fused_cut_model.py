"""The two claims the fused encoder's work distribution rests on (gcn10_b200/csrc/cn_deflate_fused.cuh), checked on a
plain-Python model of its greedy row parse -- no GPU, no product code: the kernel itself is covered bit for bit by
tests/test_gpu_deflate.py and tests/test_gpu_fullsize.py, this file pins the REASON its cuts are free.

The parse (fused_parse_row): at pixel x the candidates are the run of pixels equal to the row above (distance 256) and
the run of pixels equal to pixel x - 1 (distance 1); the longer one, if at least 3 long, becomes a match token,
otherwise pixel x becomes a literal.  A token never crosses the end of its item.

  1. Items.  A row cut at pixels that differ from the pixel above AND from the pixel to their left parses, item by
     item, into exactly the tokens of the uncut row (step 2 of the kernel chooses such pixels).
  2. Pieces.  The parse has no state: restarted at ANY of its own token boundaries it continues with the same tokens
     (what the checkpoints of the sizing pass rely on).
A cut at an arbitrary pixel, in contrast, generally changes the tokens -- which is why the fixed 64-pixel cuts of round 1
cost output bytes.
"""
import numpy as np
import pytest

W = 256


def parse(row, above, xa=0, xb=W):
    """tokens of pixels [xa, xb): (x, length, kind) with kind 'A' = match above, 'L' = run, 'l' = literal"""
    out = []
    x = xa
    while x < xb:
        la = 0
        if above is not None:
            while x + la < W and row[x + la] == above[x + la]:
                la += 1
        lr = 0
        if x > 0:
            while x + lr < W and row[x + lr] == row[x - 1]:
                lr += 1
        n = min(max(la, lr), xb - x)
        if n >= 3:
            out.append((x, n, "A" if la >= lr else "L"))
            x += n
        else:
            out.append((x, 1, "l"))
            x += 1
    return out


def rows(seed, n):
    """row pairs with the structure of Curve Number tiles: runs of a few values, partly copied from the row above"""
    rng = np.random.default_rng(seed)
    for _ in range(n):
        above = np.repeat(rng.integers(0, 6, 64), rng.integers(1, 9, 64))[:W]
        above = np.pad(above, (0, W - len(above)), constant_values=7)
        row = above.copy()
        for _ in range(int(rng.integers(0, 12))):              # stretches that differ from the row above
            a = int(rng.integers(0, W - 1))
            b = min(W, a + int(rng.integers(1, 40)))
            row[a:b] = np.repeat(rng.integers(0, 6, b - a), rng.integers(1, 6, b - a))[:b - a]
        yield row, (above if rng.random() > 0.1 else None)     # a tile's first row has no row above


@pytest.mark.parametrize("seed", range(4))
def test_cuts_at_pixels_no_match_can_cross_are_free(seed):
    nrows = ncut = 0
    for row, above in rows(seed, 150):
        free = [x for x in range(1, W) if row[x] != row[x - 1] and (above is None or row[x] != above[x])]
        whole = parse(row, above)
        starts = {t[0] for t in whole}
        assert all(x in starts for x in free)                  # such a pixel always starts a token
        # the kernel's choice: at least 8 pixels apart, at most 11 per row
        cuts, last = [], 0
        for x in free:
            if x - last >= 8 and len(cuts) < 11:
                cuts.append(x)
                last = x
        bounds = [0] + cuts + [W]
        pieces = [t for a, b in zip(bounds[:-1], bounds[1:]) for t in parse(row, above, a, b)]
        assert pieces == whole
        nrows += 1
        ncut += len(cuts)
    assert ncut > nrows                                         # the model rows do get cut


@pytest.mark.parametrize("step", [1, 2, 3, 5])
def test_the_parse_restarts_at_its_own_token_boundaries(step):
    for row, above in rows(10 + step, 120):
        whole = parse(row, above)
        marks = [t[0] for t in whole[step::step]]               # a checkpoint every `step` tokens
        bounds = [0] + marks + [W]
        pieces = [t for a, b in zip(bounds[:-1], bounds[1:]) for t in parse(row, above, a, b)]
        assert pieces == whole


def test_cuts_at_arbitrary_pixels_are_not_free():
    more = same = 0
    for row, above in rows(99, 200):
        whole = parse(row, above)
        pieces = [t for a in range(0, W, 64) for t in parse(row, above, a, a + 64)]
        assert len(pieces) >= len(whole)
        more += len(pieces) > len(whole)
        same += len(pieces) == len(whole)
    assert more > same // 4                                     # fixed 64-pixel cuts do add tokens on such rows
