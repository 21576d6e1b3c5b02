"""Multi-GPU plumbing for the benchmark harness: one process per GPU (torch.distributed), blocks
sharded across ranks with NO data-path collective -- GCN10 blocks are independent
(/root/reference/src/main.c:171; paper.md:165-166).  The only communication is the timing barrier and a
MAX reduction of the per-rank elapsed time.  The gcn10 executable itself does not use this module: it
runs one worker thread per GPU inside one process (gcn10_b200/host/host_pipeline.c)."""
from __future__ import annotations

import os


def env_world():
    """(rank, local_rank, world_size) from the torchrun environment (1 process if absent)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def shard_blocks(block_ids, rank: int, world: int):
    """The reference's static round-robin: rank r takes ids[r], ids[r+world], ... (main.c:171)."""
    return list(block_ids[rank::world])


class Group:
    """Barrier + max-over-ranks on either backend ('nccl' on GPUs, 'gloo' in the CPU tests)."""

    def __init__(self, backend: str, device=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank, self.local_rank, self.world = env_world()
        self.device = device
        self.active = self.world > 1
        if self.active:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("MASTER_PORT", "29511")
            kw = {"device_id": device} if backend == "nccl" and device is not None else {}
            dist.init_process_group(backend, rank=self.rank, world_size=self.world, **kw)

    def barrier(self):
        if self.active:
            self.dist.barrier()
        if self.device is not None and self.torch.cuda.is_available():
            self.torch.cuda.synchronize()

    def max(self, x: float) -> float:
        if not self.active:
            return float(x)
        t = self.torch.tensor([float(x)], dtype=self.torch.float64,
                              device=self.device if self.device is not None else "cpu")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum(self, x: float) -> float:
        if not self.active:
            return float(x)
        t = self.torch.tensor([float(x)], dtype=self.torch.float64,
                              device=self.device if self.device is not None else "cpu")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def gather_ints(self, value: int):
        """[value of rank 0, value of rank 1, ...] on every rank."""
        if not self.active:
            return [int(value)]
        t = self.torch.zeros(self.world, dtype=self.torch.int64,
                             device=self.device if self.device is not None else "cpu")
        t[self.rank] = int(value)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return [int(v) for v in t.tolist()]

    def close(self):
        if self.active:
            self.dist.destroy_process_group()


class BlockQueue:
    """The per-GPU block work queue of north_star item (3) across the ranks of one torchrun job.

    The reference hands blocks out statically, `for (i = rank; i < n; i += size)` (main.c:171); the gcn10 executable
    replaces that with an atomic counter shared by its worker threads (host_pipeline.c, loader_main).  Between
    PROCESSES the same counter lives in the rendezvous store: a claim is one atomic fetch-and-add
    (`Store.add`), so a rank that finishes early simply takes the next block.  No data moves between ranks."""

    def __init__(self, group: Group, n_blocks: int, name: str):
        self.n = int(n_blocks)
        self.key = f"gcn10/queue/{name}"
        self.local = 0
        self.store = None
        if group.active:
            from torch.distributed.distributed_c10d import _get_default_store
            self.store = _get_default_store()

    def claim(self):
        """Index of the next unclaimed block, or None when the queue is empty."""
        if self.store is not None:
            i = int(self.store.add(self.key, 1)) - 1
        else:
            i = self.local
            self.local += 1
        return i if i < self.n else None


def whole_job_rate(units_per_rank: float, group: Group, elapsed_s: float) -> float:
    """value = units all ranks processed / max-over-ranks time."""
    total = group.sum(units_per_rank)
    return total / group.max(elapsed_s)
