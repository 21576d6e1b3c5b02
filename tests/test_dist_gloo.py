"""World-size-2 gloo test of the benchmark's multi-GPU plumbing (block sharding, max-over-ranks timing,
whole-job aggregation).  The data path has no collective, so this is all there is to test on CPU."""
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import json, os, sys
    sys.path.insert(0, %r)
    from gcn10_b200 import dist
    g = dist.Group("gloo")
    ids = list(range(100, 117))
    mine = dist.shard_blocks(ids, g.rank, g.world)
    g.barrier()
    elapsed = 1.0 + g.rank            # rank 1 is the slow one
    rate = dist.whole_job_rate(len(mine) * 10.0, g, elapsed)
    # the dynamic queue (bench.py e2e.queue64): rank 1 is slow, so rank 0 ends up with more blocks
    import time
    q = dist.BlockQueue(g, 23, "t")
    claimed = []
    while True:
        i = q.claim()
        if i is None:
            break
        claimed.append(i)
        time.sleep(0.002 + 0.02 * g.rank)
    counts = g.gather_ints(len(claimed))
    out = dict(rank=g.rank, world=g.world, mine=mine, tmax=g.max(elapsed), rate=rate, claimed=claimed, counts=counts)
    g.barrier()
    g.close()
    print("RESULT " + json.dumps(out), flush=True)
""") % ROOT


def test_two_rank_gloo_sharding_and_timing(tmp_path):
    import json
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), LOCAL_RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1",
                   MASTER_PORT="29533")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.PIPE, text=True))
    res = {}
    for p in procs:
        out, err = p.communicate(timeout=180)
        assert p.returncode == 0, err
        line = [ln for ln in out.splitlines() if ln.startswith("RESULT ")][0]
        d = json.loads(line[7:])
        res[d["rank"]] = d
    ids = list(range(100, 117))
    assert res[0]["mine"] == ids[0::2] and res[1]["mine"] == ids[1::2]          # main.c:171 round-robin
    assert sorted(res[0]["mine"] + res[1]["mine"]) == ids                       # every block exactly once
    assert res[0]["tmax"] == res[1]["tmax"] == 2.0                              # max over ranks
    assert res[0]["rate"] == res[1]["rate"] == 170.0 / 2.0                      # all units / slowest rank
    # dynamic queue: every block exactly once, the fast rank took more, both ranks agree on the counts
    assert sorted(res[0]["claimed"] + res[1]["claimed"]) == list(range(23))
    assert len(res[0]["claimed"]) > len(res[1]["claimed"]) >= 1
    assert res[0]["counts"] == res[1]["counts"] == [len(res[0]["claimed"]), len(res[1]["claimed"])]


def test_single_process_group_is_inert():
    sys.path.insert(0, ROOT)
    from gcn10_b200 import dist
    env = {k: os.environ.pop(k) for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE") if k in os.environ}
    try:
        g = dist.Group("gloo")
        assert (g.rank, g.world, g.active) == (0, 1, False)
        assert g.max(3.5) == 3.5 and dist.whole_job_rate(7.0, g, 2.0) == 3.5
        assert dist.shard_blocks([1, 2, 3], 0, 1) == [1, 2, 3]
        q = dist.BlockQueue(g, 3, "solo")
        assert [q.claim() for _ in range(5)] == [0, 1, 2, None, None] and g.gather_ints(7) == [7]
        g.close()
    finally:
        os.environ.update(env)
