"""Property-style GPU parity: seeded random block shapes, pixel sizes, HSG ratios, origin shifts, data
profiles and plane masks; every output byte must equal the oracle's."""
import numpy as np
import pytest

from gcn10_b200 import capi
from tests.cases import make_block, PX, PX_VRT

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", range(24))
def test_random_block(seed, gpu_ctx, port, tables):
    rng = np.random.default_rng(1000 + seed)
    w = int(rng.choice([1, 7, 16, 31, 255, 256, 257, 1023, 4095, 4096, 4097, 6000, int(rng.integers(1, 9000))]))
    h = int(rng.choice([1, 2, 11, 12, 13, 25, 255, 256, 257, int(rng.integers(1, 1200))]))
    px = float(rng.choice([PX, PX_VRT, 1.0 / 3600.0, 0.00025]))
    ratio = float(rng.choice([25.0, 25.0, 10.0, 3.0, 1.0, 0.5, 7.3, 60.0]))
    shift = (float(rng.uniform(0, 2)) * px * ratio, float(rng.uniform(0, 2)) * px * ratio)
    kw = dict(w=w, h=h, px=px, hsg_px=px * ratio, shift=shift, margin=int(rng.integers(0, 3)),
              lon0=float(rng.integers(-180, 177)), lat0=float(rng.integers(-57, 84)), seed=seed,
              profile=str(rng.choice(["worldcover", "random", "coastal"])),
              esa_patch=int(rng.choice([8, 48, 192])), hsg_patch=int(rng.choice([1, 3, 9])))
    b = make_block(**kw)
    mask = int(rng.choice([capi.MASK_ALL, capi.MASK_DRAINED, capi.MASK_UNDRAINED, int(rng.integers(1, 1 << 18))]))
    want = port.block_rows(b["esa"], b["gt"], b["hsg"], b["soil_gt"], tables)
    got = gpu_ctx.block(b["esa"], b["gt"], b["hsg"], b["soil_gt"], plane_mask=mask)
    for k in range(18):
        if mask & (1 << k):
            assert np.array_equal(got[k], want[k]), (kw, hex(mask), k)
        else:
            assert not got[k].any()
