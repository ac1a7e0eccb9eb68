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


# ------------------------------------------------------------------------------------------------
# `process` of ONE sample sharded by genomic tile: tile k owns the k-th slice of the site table
# (spl_set_tile) and needs every alignment whose reference span can touch one of its sites; alignments that reach across
# a tile edge go to both tiles, whole (compSplicing looks at all junctions of a read).  No exchange step: the owned
# slices of the per-tile results concatenate to the single-context result.
# ------------------------------------------------------------------------------------------------
def max_reference_span(records):
    """Longest reference span (M/D/N/=/X lengths) of any record: how far to the left of a tile a read that still
    reaches it can start."""
    import numpy as np
    if len(records) == 0 or len(records.cigar) == 0:
        return 0
    op = records.cigar & 15
    adv = np.where((op == 0) | (op == 2) | (op == 3) | (op == 7) | (op == 8), records.cigar >> 4, 0).astype(np.int64)
    csum = np.concatenate([[0], np.cumsum(adv)])
    span = csum[records.cig_off[1:].astype(np.int64)] - csum[records.cig_off[:-1].astype(np.int64)]
    return int(span.max()) if len(span) else 0


def tile_position_ranges(table, n_chrom, tile, n_tiles):
    """Per chromosome, the [lowest, highest] position of the sites tile `tile` of `n_tiles` owns (None: owns nothing there).
    `table` is the site table in library order (api.build_site_table: exact in every strand regime)."""
    import numpy as np
    lo, hi = tile_of(tile, n_tiles, len(table))
    out = [None] * n_chrom
    if hi > lo:
        chrom, pos = np.asarray(table.chrom[lo:hi]), np.asarray(table.pos[lo:hi])
        for c in np.unique(chrom):
            p = pos[chrom == c]
            out[int(c)] = (int(p.min()), int(p.max()))
    return out


def tile_records(records, table, n_chrom, tile, n_tiles, max_span=None):
    """The records tile `tile` needs: per chromosome segment, the contiguous run of (coordinate-sorted) records that start
    no later than the tile's last site + 1 and no earlier than its first site - max_span.  Returns a Records."""
    import numpy as np
    from .api import Records
    if max_span is None:
        max_span = max_reference_span(records)
    ranges = tile_position_ranges(table, n_chrom, tile, n_tiles)
    pos_l, flag_l, cig_l, ncig_l, seg_chrom, seg_off = [], [], [], [], [], [0]
    total = 0
    for s in range(len(records.seg_chrom)):
        c = int(records.seg_chrom[s])
        r0, r1 = int(records.seg_off[s]), int(records.seg_off[s + 1])
        if c < 0 or c >= n_chrom or ranges[c] is None or r1 == r0:
            continue
        p = records.pos[r0:r1]
        if np.any(p[1:] < p[:-1]):                      # not coordinate-sorted: keep the whole segment
            a, b = 0, r1 - r0
        else:
            a = int(np.searchsorted(p, ranges[c][0] - max_span - 1, side="left"))
            b = int(np.searchsorted(p, ranges[c][1] + 1, side="right"))
        if b <= a:
            continue
        c0, c1 = int(records.cig_off[r0 + a]), int(records.cig_off[r0 + b])
        pos_l.append(p[a:b]); flag_l.append(records.flag[r0 + a:r0 + b]); cig_l.append(records.cigar[c0:c1])
        ncig_l.append(np.diff(records.cig_off[r0 + a:r0 + b + 1].astype(np.int64)))
        seg_chrom.append(c)
        total += b - a
        seg_off.append(total)
    if not pos_l:
        return Records(np.zeros(0, np.int32), np.zeros(0, np.uint16), np.zeros(1, np.uint32), np.zeros(0, np.uint32),
                       np.zeros(0, np.int32), np.zeros(1, np.int64))
    off = np.zeros(total + 1, np.uint32)
    np.cumsum(np.concatenate(ncig_l), out=off[1:])
    return Records(np.concatenate(pos_l), np.concatenate(flag_l), off, np.concatenate(cig_l), seg_chrom, seg_off)
