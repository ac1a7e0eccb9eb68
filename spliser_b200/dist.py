"""One process per GPU: the only cross-rank traffic of this path is the timing barrier and the max / sum
reductions of the benchmark numbers (the counting itself shards by genomic tile or by sample with no
exchange step).  Works over NCCL (GPU tensors) and gloo (CPU tensors, used by the CPU tests)."""
from __future__ import annotations

import os


class Ranks:
    def __init__(self, backend=None):
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        self.dist = None
        self.device = "cpu"
        if self.world > 1:
            import torch
            import torch.distributed as dist
            backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
            if backend == "nccl":
                torch.cuda.set_device(self.local)
                self.device = "cuda"
                dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
            else:
                dist.init_process_group("gloo")
            self.dist = dist

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()

    def _reduce(self, x, op):
        if self.dist is None:
            return float(x)
        import torch
        t = torch.tensor([float(x)], dtype=torch.float64, device=self.device)
        self.dist.all_reduce(t, op=op)
        return float(t.item())

    def max(self, x):
        return self._reduce(x, self.dist.ReduceOp.MAX) if self.dist is not None else float(x)

    def sum(self, x):
        return self._reduce(x, self.dist.ReduceOp.SUM) if self.dist is not None else float(x)

    def close(self):
        if self.dist is not None:
            self.dist.destroy_process_group()
            self.dist = None


def tile_of(rank, world, n_sites):
    """Owned site slice [lo, hi) of a rank when one sample is sharded by genomic tile (spl_set_tile)."""
    return n_sites * rank // world, n_sites * (rank + 1) // world


def samples_of(rank, world, n_samples):
    """Samples a rank re-counts when `combine` is sharded by sample."""
    return list(range(rank, n_samples, world))
