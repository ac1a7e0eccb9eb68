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


# ------------------------------------------------------------------------------------------------
# Host topology of a rank: which NUMA node its GPU hangs off and which CPUs the process may run on.  Page-locked
# staging memory is placed on the node of the thread that allocates it, so a rank running on the other socket uploads its
# records across the socket interconnect.  host_topology() only reports (bench.py prints it with every line);
# bind_to_device_node() restricts the calling process to the GPU's node -- opt-in (SPLISER_NUMA_BIND=1 in bench.py)
# until its effect on the end-to-end number has been measured on the target box.
# ------------------------------------------------------------------------------------------------
def parse_cpulist(text):
    """'0-3,8,10-11' (sysfs cpulist) -> sorted list of CPU numbers; malformed pieces are skipped."""
    cpus = set()
    for part in str(text).strip().split(","):
        part = part.strip()
        if not part:
            continue
        try:
            if "-" in part:
                a, b = part.split("-", 1)
                cpus.update(range(int(a), int(b) + 1))
            else:
                cpus.add(int(part))
        except ValueError:
            continue
    return sorted(cpus)


def _read(path):
    try:
        with open(path) as fh:
            return fh.read().strip()
    except OSError:
        return None


def host_topology(device, sysfs="/sys", bus_id=None):
    """-> dict(gpu_bus_id, gpu_numa_node, numa_nodes, node_cpus (of the GPU's node), cpus_allowed, cpus_allowed_on_gpu_node).
    Unknown pieces are None; never raises."""
    import subprocess
    info = dict(gpu_bus_id=None, gpu_numa_node=None, numa_nodes=None, node_cpus=None, cpus_allowed=None, cpus_allowed_on_gpu_node=None)
    try:
        allowed = sorted(os.sched_getaffinity(0))
        info["cpus_allowed"] = len(allowed)
        nodes = [d for d in os.listdir(os.path.join(sysfs, "devices/system/node")) if d.startswith("node") and d[4:].isdigit()]
        info["numa_nodes"] = len(nodes)
        if bus_id is None:
            res = subprocess.run(["nvidia-smi", "-i", str(device), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                                 capture_output=True, text=True, timeout=20)
            bus_id = res.stdout.strip().splitlines()[0].strip() if res.returncode == 0 and res.stdout.strip() else None
        if bus_id:
            info["gpu_bus_id"] = bus_id
            short = bus_id.lower()
            if len(short.split(":")[0]) == 8:                # nvidia-smi prints an 8-digit PCI domain, sysfs uses 4
                short = short[4:]
            node = _read(os.path.join(sysfs, "bus/pci/devices", short, "numa_node"))
            if node is not None and node.lstrip("-").isdigit():
                info["gpu_numa_node"] = int(node)
                if int(node) >= 0:
                    cl = _read(os.path.join(sysfs, "devices/system/node/node%d/cpulist" % int(node)))
                    if cl is not None:
                        cpus = parse_cpulist(cl)
                        info["node_cpus"] = len(cpus)
                        info["cpus_allowed_on_gpu_node"] = len(set(cpus) & set(allowed))
                        info["_bind"] = sorted(set(cpus) & set(allowed))
    except Exception:                                        # noqa: BLE001 -- reporting only
        pass
    return info


def bind_to_device_node(info):
    """Restricts the calling process to the allowed CPUs of its GPU's NUMA node (host_topology(...) result).  Returns True if
    the affinity was changed.  No-op when the node is unknown, the box has one node, or no allowed CPU sits on that node."""
    cpus = info.get("_bind") or []
    if not cpus or (info.get("numa_nodes") or 1) < 2 or len(cpus) == (info.get("cpus_allowed") or 0):
        return False
    try:
        os.sched_setaffinity(0, cpus)
        return True
    except OSError:
        return False


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
def _reference_spans(records):
    """Reference span (sum of M/D/N/=/X lengths) of every record."""
    import numpy as np
    if len(records) == 0:
        return np.zeros(0, np.int64)
    op = records.cigar & 15
    adv = np.where((op == 0) | (op == 2) | (op == 3) | (op == 7) | (op == 8), records.cigar >> 4, 0).astype(np.int64)
    csum = np.concatenate([[0], np.cumsum(adv)])
    return csum[records.cig_off[1:].astype(np.int64)] - csum[records.cig_off[:-1].astype(np.int64)]


def max_reference_span(records):
    """Longest reference span of any record: how far to the left of a tile a read that still reaches it can start."""
    span = _reference_spans(records)
    return int(span.max()) if len(span) else 0


def segment_max_spans(records):
    """Longest reference span per chromosome segment: one 500 kb intron on one chromosome must not widen the left edge of
    every tile on every other chromosome."""
    import numpy as np
    span = _reference_spans(records)
    return [int(span[int(records.seg_off[s]):int(records.seg_off[s + 1])].max()) if records.seg_off[s + 1] > records.seg_off[s] else 0
            for s in range(len(records.seg_chrom))]


def balanced_tiles(records, table, n_chrom, n_tiles):
    """Site-index cut points [0 = c_0 <= c_1 <= ... <= c_n = S] that give every tile about the same number of records
    (SURVEY 8(e)): the number of records starting at or before each site is a prefix sum over the per-chromosome record
    positions; tile k ends at the first site where that count reaches k/n of the sample."""
    import numpy as np
    S = len(table)
    if S == 0 or n_tiles <= 1:
        return [0] + [S] * max(1, n_tiles)
    per_chrom = [[] for _ in range(n_chrom)]
    for s in range(len(records.seg_chrom)):
        c = int(records.seg_chrom[s])
        if 0 <= c < n_chrom:
            per_chrom[c].append(np.asarray(records.pos[int(records.seg_off[s]):int(records.seg_off[s + 1])]))
    cum = np.zeros(S, np.int64)
    base = 0
    chrom, pos = np.asarray(table.chrom), np.asarray(table.pos)
    for c in range(n_chrom):
        sel = np.nonzero(chrom == c)[0]
        p = np.concatenate(per_chrom[c]) if per_chrom[c] else np.zeros(0, np.int32)
        if len(p) > 1 and np.any(p[1:] < p[:-1]):
            p = np.sort(p)
        if len(sel):
            cum[sel] = base + np.searchsorted(p, pos[sel], side="right")
        base += len(p)
    cum = np.maximum.accumulate(cum)
    total = max(base, 1)
    cuts = [0]
    for k in range(1, n_tiles):
        cuts.append(max(cuts[-1], int(np.searchsorted(cum, total * k // n_tiles, side="left"))))
    cuts.append(S)
    return cuts


def tile_position_ranges(table, n_chrom, tile, n_tiles, site_range=None):
    """Per chromosome, the [lowest, highest] position of the sites a tile owns (None: owns nothing there).  The tile is the
    `tile`-th equal slice of the site table, or the explicit site index range `site_range`.
    `table` is the site table in library order (api.build_site_table: exact in every strand regime)."""
    import numpy as np
    lo, hi = site_range if site_range is not None else tile_of(tile, n_tiles, len(table))
    out = [None] * n_chrom
    if hi > lo:
        chrom, pos = np.asarray(table.chrom[lo:hi]), np.asarray(table.pos[lo:hi])
        for c in np.unique(chrom):
            p = pos[chrom == c]
            out[int(c)] = (int(p.min()), int(p.max()))
    return out


def tile_records(records, table, n_chrom, tile, n_tiles, max_span=None, site_range=None, seg_spans=None):
    """The records a tile needs: per chromosome segment, the contiguous run of (coordinate-sorted) records that start
    no later than the tile's last site + 1 and no earlier than its first site - the segment's longest reference span
    (seg_spans, from segment_max_spans; or one global max_span).  Returns a Records."""
    import numpy as np
    from .api import Records
    if seg_spans is None:
        if max_span is None:
            seg_spans = segment_max_spans(records)
        else:
            seg_spans = [max_span] * len(records.seg_chrom)
    ranges = tile_position_ranges(table, n_chrom, tile, n_tiles, site_range)
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
            a = int(np.searchsorted(p, ranges[c][0] - seg_spans[s] - 1, side="left"))
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


# ------------------------------------------------------------------------------------------------
# The junction rows a tile needs (the site table + competing-site graph are then built per tile, not replicated).
# An owned site t needs: every row touching t (alpha, Partners in insertion order, PartnerCounts: S:341-355) and every
# row touching a partner p of t (p's PartnerCounts and alpha for findBeta2Counts, S:590-613; the partners of p are t's
# competitors, S:364-372).  A partner lies within the chromosome's longest junction of t, so the rows with an endpoint
# inside [first owned position - H, last owned position + H] are a superset.  The rows keep their BED order, so the
# sub-table's owned rows equal the full table's in every column.
# ------------------------------------------------------------------------------------------------
def is_dirty_regime(junctions, flags):
    """Stranded run with a strand byte other than + / - (or a zero-length row): the reference's site identity then depends
    on the path of its bisection through the WHOLE list, so the table is not cut."""
    import numpy as np
    if len(junctions) == 0:
        return False
    if np.any(junctions.left == junctions.right):
        return True
    return bool(flags & 1) and bool(np.any((junctions.strand != 43) & (junctions.strand != 45)))


def tile_junctions(junctions, table, n_chrom, site_range, flags):
    """(rows kept as indices into `junctions`, sub-table Junctions, owned slice [lo', hi') of the sub-table's site table).
    `table`: the full site table in library order (api.build_site_table)."""
    import numpy as np
    from .api import Junctions
    lo, hi = site_range
    J = len(junctions)
    if is_dirty_regime(junctions, flags) or hi <= lo:
        return np.arange(J, dtype=np.int64), junctions, (lo, hi)
    ranges = tile_position_ranges(table, n_chrom, 0, 1, site_range=(lo, hi))
    keep = np.zeros(J, bool)
    jc, jl, jr = junctions.chrom, junctions.left.astype(np.int64), junctions.right.astype(np.int64)
    # sites sit at both ends of a row; the alpha of a site counts rows by position, whatever the orientation of the row
    a, b = np.minimum(jl, jr), np.maximum(jl, jr)
    for c in range(n_chrom):
        if ranges[c] is None:
            continue
        on = jc == c
        if not on.any():
            continue
        H = int((b[on] - a[on]).max()) + 1
        w0, w1 = ranges[c][0] - H, ranges[c][1] + H
        keep |= on & (((a >= w0) & (a <= w1)) | ((b >= w0) & (b <= w1)))
    idx = np.nonzero(keep)[0].astype(np.int64)
    sub = Junctions(junctions.chrom[idx], junctions.left[idx], junctions.right[idx], junctions.score[idx], junctions.strand[idx])
    # the owned sites keep their order in the sub-table and every site at an owned position is present: the slice starts
    # after the sub-table sites that sort before the first owned (chromosome, position), plus the same-position sites before it
    chrom, pos = np.asarray(table.chrom), np.asarray(table.pos)
    c0, p0 = int(chrom[lo]), int(pos[lo])
    first_same = lo
    while first_same > 0 and int(chrom[first_same - 1]) == c0 and int(pos[first_same - 1]) == p0:
        first_same -= 1
    ends_c = np.concatenate([sub.chrom, sub.chrom]).astype(np.int64)
    ends_p = np.concatenate([sub.left, sub.right]).astype(np.int64)
    stranded = bool(flags & 1)
    if stranded:
        ends_s = np.concatenate([sub.strand, sub.strand]).astype(np.int64)
        key = (ends_c << 40) | (ends_p << 8) | ends_s
    else:
        key = (ends_c << 40) | (ends_p << 8)
    before = (ends_c < c0) | ((ends_c == c0) & (ends_p < p0))
    n_before = len(np.unique(key[before]))
    lo2 = n_before + (lo - first_same)
    return idx, sub, (lo2, lo2 + (hi - lo))


_PART_COLS = ("chrom", "pos", "strand", "first_line", "alpha", "beta1", "beta2simple", "beta2cryptic", "beta2weighted", "sse")


def owned_part(table, local_range, rows_kept=None):
    """The owned rows of a tile's result as a dict of arrays (every column; CSR columns as lengths + values).
    rows_kept maps the sub-table's BED row numbers (first_line) back to the full junction table's."""
    import numpy as np
    lo, hi = local_range
    out = {k: np.asarray(getattr(table, k))[lo:hi].copy() for k in _PART_COLS}
    if rows_kept is not None and hi > lo:
        out["first_line"] = np.asarray(rows_kept)[out["first_line"].astype(np.int64)].astype(out["first_line"].dtype)
    po, co = np.asarray(table.partner_off).astype(np.int64), np.asarray(table.comp_off).astype(np.int64)
    out["partner_len"] = np.diff(po[lo:hi + 1])
    out["partner_pos"] = np.asarray(table.partner_pos)[po[lo]:po[hi]].copy()
    out["partner_cnt"] = np.asarray(table.partner_cnt)[po[lo]:po[hi]].copy()
    out["comp_len"] = np.diff(co[lo:hi + 1])
    out["comp_pos"] = np.asarray(table.comp_pos)[co[lo]:co[hi]].copy()
    return out


def concat_parts(parts, like):
    """Table dict (the SiteTable column names as keys) from the owned parts of every tile, in tile order; dtypes as in `like`."""
    import numpy as np
    out = {k: np.concatenate([p[k] for p in parts]).astype(np.asarray(like[k]).dtype) for k in _PART_COLS}
    for name, ln in (("partner", "partner_len"), ("comp", "comp_len")):
        lens = np.concatenate([p[ln] for p in parts]).astype(np.int64)
        off = np.zeros(len(lens) + 1, np.asarray(like[name + "_off"]).dtype)
        np.cumsum(lens, out=off[1:])
        out[name + "_off"] = off
    out["partner_pos"] = np.concatenate([p["partner_pos"] for p in parts]).astype(np.asarray(like["partner_pos"]).dtype)
    out["partner_cnt"] = np.concatenate([p["partner_cnt"] for p in parts]).astype(np.asarray(like["partner_cnt"]).dtype)
    out["comp_pos"] = np.concatenate([p["comp_pos"] for p in parts]).astype(np.asarray(like["comp_pos"]).dtype)
    return out
