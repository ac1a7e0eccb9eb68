#!/usr/bin/env python
"""Benchmark of the SpliSER counting path (BASELINE.json metric: aligned reads/s to per-site SSE).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c1|c3|small]
    torchrun --nproc-per-node N ... bench.py --gpus N ...          (one rank per GPU, weak scaling)

One "step" = one pass of the hot path over one sample: read SoA -> per-site alpha, beta1, beta2Simple,
beta2Cryptic, SSE.  Per rank the workload is BASELINE.json configs[1] (full A. thaliana genome, 40M
150 bp PE records, --isStranded -s rf) generated synthetically with a per-rank seed.

  value     reads/s with the SoA already resident in HBM (CUDA events on the library's stream)
  e2e       reads/s through the C ABI from (pinned) host record arrays: junction table -> site graph,
            H2D copies, record expansion, counting, D2H of the per-site results, all inside the timed region
  roofline  dominant kernel (k_beta1_stab): algorithmic bytes / its CUDA-event time vs measured HBM peak
  cpu_baseline / --impl reference: the oracle's C port of the reference algorithm on the host cores,
            on a bounded genomic sub-region of the same workload (same coverage density)
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "aligned reads/s to per-site SSE"
CACHE = os.environ.get("SPLISER_BENCH_CACHE", os.path.join(tempfile.gettempdir(), "spliser_bench_cache"))


def workload_config(name, reads, rank):
    from spliser_b200 import synth
    if name == "c2":
        cfg = synth.config_c2(reads or 40_000_000)
        desc = "configs[1]: full A. thaliana genome (TAIR10 contig lengths), %d 150 bp PE records, --isStranded -s rf" % cfg.n_records
    elif name == "c1":
        cfg = synth.config_c1()
        if reads:
            cfg.n_records = reads
        desc = "configs[0]: Chr1-sized contig, %d 100 bp SE records, unstranded" % cfg.n_records
    elif name == "c3":
        cfg = synth.config_c3_tile(reads or 25_000_000, tile=rank % 8)
        desc = "configs[2]: one genomic tile (3 contigs) of a GRCh38-scale sample, %d 150 bp PE records, stranded rf" % cfg.n_records
    elif name == "small":
        cfg = synth.config_small(reads or 200_000, seed=3, stranded=True, paired=True)
        desc = "small test workload, %d records" % cfg.n_records
    else:
        raise SystemExit("unknown workload %s" % name)
    cfg.seed += 7919 * rank          # every rank counts its own sample (weak scaling)
    return cfg, desc


def region_sample(w, max_reads):
    """First `max_reads` records of the first chromosome segment + the junctions inside that region:
    a genomic sub-region with the workload's own coverage density."""
    from spliser_b200 import Junctions, Records
    r, j = w.records, w.junctions
    n = int(min(max_reads, r.seg_off[1] - r.seg_off[0])) if len(r.seg_chrom) else 0
    if n == 0:
        return r, j
    cut = int(r.pos[n - 1])
    c0 = int(r.seg_chrom[0])
    rec = Records(r.pos[:n], r.flag[:n], r.cig_off[:n + 1], r.cigar[:int(r.cig_off[n])], [c0], [0, n])
    keep = (j.chrom == c0) & (j.right <= cut)
    return rec, Junctions(j.chrom[keep], j.left[keep], j.right[keep], j.score[keep], j.strand[keep])


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(device), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def window(self, t0, t1):
        sm, mx, reasons = [], 0.0, set()
        rows = [r for t, r in self.rows if t0 - 0.05 <= t <= t1 + 0.15] or [r for _, r in self.rows[-3:]]
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            try:
                sm.append(float(f[0])); mx = max(mx, float(f[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}

    def stop(self):
        if self.proc:
            self.proc.terminate()


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    return 6650.0, "B200_PROFILING.md fallback 6.65 TB/s (of fallback)"


def traffic_from_profile():
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get("k_beta1_stab_dram_bytes_per_launch")
        except (ValueError, OSError):
            return None
    return None


def python_reference_note():
    """The unmodified (Python) reference cannot run on the GPU box; oracle/time_reference.py timed it in the authoring
    container on configs[0] and recorded how much faster the C port timed here is (profiles/r1_reference_python_c1.json)."""
    p = os.path.join(ROOT, "profiles", "r1_reference_python_c1.json")
    try:
        d = json.load(open(p))
        return {"reads_per_s": d["reference_reads_per_s"], "cores": 1, "workload": "configs[0], authoring container, samtools fork replaced by an in-process indexed store (its time excluded)",
                "c_port_over_python_reference_1_thread": d["c_port_over_reference"]["threads_1"], "source": "profiles/r1_reference_python_c1.json"}
    except (OSError, ValueError, KeyError):
        return None


def cpu_port_run(rec, junc, n_chrom, flags, threads):
    from oracle import c_oracle
    t0 = time.perf_counter()
    c_oracle.process(rec, n_chrom, junc, flags, threads=threads)
    return time.perf_counter() - t0


def synthetic_gff(w, path):
    """A gene every ~4 kb on both strands (seeded): only the Gene column of the TSV depends on it."""
    rng = np.random.default_rng(20260099)
    with open(path, "w") as fh:
        for c, n in zip(w.chroms, w.chrom_len):
            k = max(1, int(n) // 4000)
            starts = np.sort(rng.integers(1, max(2, int(n) - 6000), size=k))
            lens = rng.integers(500, 6000, size=k)
            strand = rng.integers(0, 2, size=k)
            fh.writelines("%s\tsynthetic\tgene\t%d\t%d\t.\t%s\t.\tID=%s_G%06d;Name=n%d\n" % (c, a, a + b, "+-"[d], c, i, i)
                          for i, (a, b, d) in enumerate(zip(starts.tolist(), lens.tolist(), strand.tolist())))


def cli_process_run(ctx, w, bam, args, reads):
    """`spliser_b200.cli.process` on files, timed as a whole and per stage (host wall clock)."""
    from spliser_b200 import cli
    from spliser_b200.genes import load_annotation
    stem = os.path.join(CACHE, "bench_%s_%d" % (args.workload, len(w.records)))
    bed, gff, outp = stem + ".bed", stem + ".gff", stem + ".out"
    if not os.path.exists(bed):
        open(bed + ".tmp", "w").write(w.bed12_text())
        os.replace(bed + ".tmp", bed)
    if not os.path.exists(gff):
        synthetic_gff(w, gff + ".tmp")
        os.replace(gff + ".tmp", gff)
    stranded = bool(w.flags & 1)
    kw = dict(annotationFile=gff, isStranded=stranded, strandedType="rf" if stranded else None, ctx=ctx)
    cli.process(bam, bed, outp, **kw)                               # warm-up (page cache, allocations)
    ts = []
    for _ in range(max(1, args.e2e_steps)):
        a = time.perf_counter()
        cli.process(bam, bed, outp, **kw)
        ts.append(time.perf_counter() - a)
    a = time.perf_counter()
    ann = load_annotation(gff)
    b = time.perf_counter()
    with open(bed) as fh:
        chroms, table, sstr = cli.process_table(ctx, bam, fh, annotation=ann, is_stranded=stranded, stranded_type=kw["strandedType"])
    c = time.perf_counter()
    ms_count = ctx.stats()["ms_total"]
    cli.write_process_tsv(outp + ".SpliSER.tsv", chroms, table, sstr, annotation=ann, is_stranded=stranded)
    d = time.perf_counter()
    n_rows = sum(1 for _ in open(outp + ".SpliSER.tsv")) - 1
    step = float(np.mean(ts))
    return {"value": reads / step, "unit": "reads/s", "ms_per_step": 1e3 * step, "rows_written": n_rows, "tsv_bytes": os.path.getsize(outp + ".SpliSER.tsv"),
            "genes": sum(len(g) for g in ann.genes),
            "breakdown_ms": {"annotation": round(1e3 * (b - a), 2), "bed_parse+spl_process": round(1e3 * (c - b), 2), "of_which_spl_process": round(ms_count, 2),
                             "gene_column+tsv_writer": round(1e3 * (d - c), 2)},
            "note": "cli.process(BAM, BED12, GFF) -> .SpliSER.tsv: annotation and BED12 parse (native), spl_process(bam) on the device, "
                    "Gene column + TSV writer (native); byte-identical output is pinned by tests/test_cli_*.py"}


def emit(line):
    """The ONE JSON line goes to the real stdout; everything else a library prints (NCCL banner, warnings) was
    redirected to stderr at start-up."""
    os.write(_REAL_STDOUT, (line + "\n").encode())


_REAL_STDOUT = 1


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)                      # fd 1 -> stderr for the rest of the process (C libraries included)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--reads", type=int, default=0, help="records per GPU (default: the named config's size)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-sample", type=int, default=2_000_000, help="records in the CPU baseline's genomic sub-region")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-bam", action="store_true", help="skip the from-a-BAM-file measurement")
    ap.add_argument("--profile", action="store_true", help="resident passes only (for ncu): no e2e, no CPU baseline")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: one sample per GPU (default); strong: ONE sample sharded over the GPUs by genomic tile "
                         "(each rank gets the records of its tile, edge-spanning reads duplicated, and counts the sites it owns)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    ncores = os.cpu_count() or 1

    from spliser_b200 import synth

    # ------------------------------------------------------------------ reference arm: CPU port
    if args.impl == "reference":
        if rank != 0:
            return
        cfg, desc = workload_config(args.workload, args.reads, 0)
        w = synth.generate(cfg, cache_dir=CACHE)
        rec, junc = region_sample(w, args.cpu_sample)
        times = []
        for i in range(args.warmup + args.steps):
            dt = cpu_port_run(rec, junc, len(w.chroms), w.flags, ncores)
            if i >= args.warmup:
                times.append(dt)
        tot = float(sum(times))
        val = len(rec) * len(times) / tot
        sample = "first %d records of %s (one genomic sub-region, same coverage) + its %d junctions, per step" % (len(rec), w.chroms[int(rec.seg_chrom[0])], len(junc))
        emit(json.dumps({
            "impl": "reference", "metric": METRIC, "value": val, "unit": "reads/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * tot / len(times), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int64", "data": "synthetic",
            "config": {"workload": desc, "note": "reference is single-threaded Python + one samtools fork per site; this arm times the oracle's C port of its algorithm (per-site read fetch + per-CIGAR-op state machine) with OpenMP over sites"},
            "cpu_baseline": {"value": val, "unit": "reads/s", "cores": ncores, "kind": "port", "sample": sample,
                             "python_reference": python_reference_note()},
            "e2e": {"value": val, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}))
        return

    # ------------------------------------------------------------------ our arm
    import spliser_b200
    from spliser_b200.api import Records, pinned_empty
    from spliser_b200.dist import Ranks
    from spliser_b200.dist import bind_to_device_node, host_topology
    topo = host_topology(local)
    topo["bound_to_gpu_node"] = bind_to_device_node(topo) if os.environ.get("SPLISER_NUMA_BIND") == "1" else False
    topo.pop("_bind", None)
    ranks = Ranks("nccl" if world > 1 else None)
    strong = args.scaling == "strong" and world > 1
    cfg, desc = workload_config(args.workload, args.reads, 0 if strong else rank)
    w = synth.generate(cfg, cache_dir=CACHE)
    n_chrom = len(w.chroms)
    n_sample = len(w.records)
    if strong:
        # the same sample on every rank; rank r keeps the records of genomic tile r and owns the r-th slice of the site table
        from spliser_b200 import api, dist
        table0 = api.build_site_table(n_chrom, w.junctions, w.flags)
        w.records = dist.tile_records(w.records, table0, n_chrom, rank, world)
        desc += " -- ONE sample sharded by genomic tile over %d GPUs" % world
        ctx = spliser_b200.Context(local, tile=(rank, world))
    else:
        ctx = spliser_b200.Context(local)
    barrier, max_over_ranks, sum_over_ranks = ranks.barrier, ranks.max, ranks.sum

    sampler = ClockSampler(local)
    # ---- resident: SoA in HBM -> SSE in HBM
    ctx.resident_load(w.records, n_chrom, w.junctions, w.flags)
    t_load = time.time()
    ctx.resident_count(args.warmup)
    barrier()
    t0 = time.time()
    st = ctx.resident_count(args.steps)
    t1 = time.time()
    barrier()
    ms_total = max_over_ranks(st["ms_total"])
    reads_rank = st["n_aligned"]
    reads_all = float(n_sample) if strong else sum_over_ranks(reads_rank)      # strong: the sample counts once, duplicates do not
    value = reads_all * args.steps / (ms_total * 1e-3)
    nA, nB, nJ, nS, S, E = (st[k] for k in ("n_mblocks_a", "n_mblocks_b", "n_junc_ops", "n_spliced", "n_sites", "n_edges"))
    peak, peak_src = measured_peak()
    # algorithmic bytes per pass of the resident layout: 8 B per M block of the bin-partitioned
    # stream; the junction kernels read the distinct-junction table (24 B), the grouped simple instances (8 B) and
    # the complex instances (16 B incl. their read's junction list entry)
    D, n_simple, n_complex = st["n_distinct_junc"], st["n_simple_junc"], st["n_complex_junc"]
    kern = {"k_beta1_stab": (8.0 * (nA + nB), st["ms_beta1"] / args.steps),
            "k_junc_*": (24.0 * D + 8.0 * n_simple + 16.0 * n_complex, st["ms_spliced"] / args.steps)}
    dom = max(kern, key=lambda k: kern[k][1])
    k3_bytes, k3_ms = kern[dom]
    k3_gbs = k3_bytes / (k3_ms * 1e-3) / 1e9 if k3_ms > 0 else 0.0
    path_bytes = kern["k_beta1_stab"][0] + kern["k_junc_*"][0] + 25.0 * S + 12.0 * E
    path_ms = st["ms_total"] / args.steps
    soa_mb = path_bytes / 1e6

    if args.profile:
        emit(json.dumps({"profile_run": True, "ms_per_step": ms_total / args.steps, "kernel_ms": {"beta1_stab": kern["k_beta1_stab"][1], "junction_kernels": kern["k_junc_*"][1], "final": st["ms_final"] / args.steps}}))
        sampler.stop()
        ctx.close()
        ranks.close()
        return
    # ---- e2e through the C ABI from pinned host arrays
    r = w.records
    pinned = {}
    for k in ("pos", "flag", "cig_off", "cigar"):
        a = getattr(r, k)
        b = pinned_empty(len(a), a.dtype)
        b[:] = a
        pinned[k] = b
    pr = Records(pinned["pos"], pinned["flag"], pinned["cig_off"], pinned["cigar"], r.seg_chrom, r.seg_off)
    ctx.process_records(pr, n_chrom, w.junctions, w.flags)       # warm-up (allocations)
    barrier()
    e2e_t = []
    stats = None
    table = None
    for _ in range(max(1, args.e2e_steps)):
        table = None                     # the previous sample's result is released before the next call (its pinned arena is reused)
        barrier()
        a = time.perf_counter()
        table = ctx.process_records(pr, n_chrom, w.junctions, w.flags)
        e2e_t.append(time.perf_counter() - a)
        stats = ctx.stats()
    e2e_step = max_over_ranks(float(np.mean(e2e_t)))
    e2e_val = reads_all / e2e_step
    # the timed passes last a few ms, shorter than nvidia-smi's sampling period: the clock record covers the whole GPU-active
    # stretch around them (warm-up passes, timed passes, end-to-end calls)
    clocks = sampler.window(t_load, time.time())
    clocks["window"] = "warm-up + timed passes + e2e calls"

    out = {
        "metric": METRIC, "value": value, "unit": "reads/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None,
        "dtype": "u32", "data": "synthetic",
        "config": {"workload": desc, "records_per_gpu": int(reads_rank), "sites_per_gpu": int(S), "junction_rows": len(w.junctions),
                   "l2": "no flush needed: the streamed SoA is %.0f MB per pass, larger than the 126 MB L2" % soa_mb,
                   "host": topo,
                   "timing": "CUDA events on the library's stream around %d passes; max over ranks" % args.steps,
                   "value_scope": "one counting pass (alpha reduce, beta1 stabbing, junction span + exceptions, beta2 gather, SSE) over the "
                                  "counting layout resident in HBM; building that layout from the raw records is load-time work: see from_records "
                                  "(device time incl. it) and e2e (host arrays -> host table, everything included)"},
        "clocks": clocks,
        "e2e": {"value": e2e_val, "unit": "reads/s", "h2d_bytes_per_step": int(stats["h2d_bytes"]), "d2h_bytes_per_step": int(stats["d2h_bytes"]),
                "ms_per_step": 1e3 * e2e_step,
                "breakdown_ms": {k: round(stats[k], 3) for k in ("ms_total", "ms_graph", "ms_upload", "ms_expand", "ms_count")},
                "graph_on_device": bool(stats["graph_on_device"]),
                "note": "host wall clock around spl_process_records: junction table -> site table + graph (device sort/unique in the clean regime), pinned H2D of the records, expansion, counting, D2H"},
        "from_records": {"note": "same metric with the RAW record arrays resident in HBM instead of the counting layout: adds the load-time kernels "
                                 "(record expansion, bin partition, junction grouping; CUDA events) to one counting pass; the site table + graph build "
                                 "(about 1 ms of small sort kernels) runs on a second stream under the record upload and is not included",
                         "ms_load_kernels": round(stats["ms_expand"], 3), "ms_count_pass": round(ms_total / args.steps, 4),
                         "value": reads_rank / ((stats["ms_expand"] + ms_total / args.steps) * 1e-3) * world, "unit": "reads/s"},
        "gpu_launches": int(st["launches"]) * args.steps,
        "roofline": {"bound": "hbm", "kernel": dom, "achieved": k3_gbs, "peak": peak, "unit": "GB/s",
                     "frac": k3_gbs / peak, "traffic": traffic_from_profile(), "algorithmic_bytes_per_launch": k3_bytes,
                     "ms_per_launch": k3_ms, "peak_source": peak_src},
        "roofline_path": {"algorithmic_bytes_per_pass": path_bytes, "ms_per_pass": path_ms,
                          "achieved_gbs": path_bytes / (path_ms * 1e-3) / 1e9, "frac": path_bytes / (path_ms * 1e-3) / 1e9 / peak,
                          "kernel_ms": {"beta1_stab": kern["k_beta1_stab"][1], "junction_kernels": kern["k_junc_*"][1], "alpha+scan+finalize": st["ms_final"] / args.steps},
                          "kernel_gbs": {k: (v[0] / (v[1] * 1e-3) / 1e9 if v[1] > 0 else 0.0) for k, v in kern.items()}},
        "kernel_path": {"n_mblocks_a": int(nA), "n_mblocks_b": int(nB), "n_junction_ops": int(nJ), "n_spliced_reads": int(nS), "n_edges": int(E),
                        "n_distinct_junctions": int(D), "n_simple_junction_instances": int(n_simple), "n_complex_junction_instances": int(n_complex)},
        "checksum": {"beta1": int(table.beta1.sum()), "beta2simple": int(table.beta2simple.sum()), "alpha": int(table.alpha.sum())},
    }
    # ---- the same call from a BAM FILE (spl_process): BGZF inflate + record parse on the device vs the host reader
    if world == 1 and not args.no_bam:
        try:
            bam = os.path.join(CACHE, "bench_%s_%d.bam" % (args.workload, len(r)))
            if not os.path.exists(bam):
                os.makedirs(CACHE, exist_ok=True)
                r.write_bam(bam + ".tmp", w.chroms, w.chrom_len)
                os.replace(bam + ".tmp", bam)
            table = None
            ctx.process_bam(bam, w.chroms, w.junctions, w.flags)       # warm-up (allocations, page cache)
            tb = []
            table_b = None
            for _ in range(max(1, args.e2e_steps)):
                table_b = None                    # release the previous result first (its pinned arena is reused)
                a = time.perf_counter()
                table_b = ctx.process_bam(bam, w.chroms, w.junctions, w.flags)
                tb.append(time.perf_counter() - a)
                sb = ctx.stats()
            os.environ["SPLISER_HOST_BAM"] = "1"
            a = time.perf_counter()
            th = ctx.process_bam(bam, w.chroms, w.junctions, w.flags)
            host_s = time.perf_counter() - a
            sh = ctx.stats()
            del os.environ["SPLISER_HOST_BAM"]
            out["bam_e2e"] = {"value": reads_rank / float(np.mean(tb)), "unit": "reads/s", "ms_per_step": 1e3 * float(np.mean(tb)),
                              "bam_bytes": os.path.getsize(bam), "bam_on_device": bool(sb["bam_on_device"]), "ms_ingest": round(sb["ms_decode"], 3),
                              "host_reader": {"ms_per_step": 1e3 * host_s, "ms_decode": round(sh["ms_decode"], 3), "threads": ncores},
                              "identical_to_host_reader": bool(np.array_equal(th.beta1, table_b.beta1) and np.array_equal(th.beta2simple, table_b.beta2simple)
                                                                and np.array_equal(th.sse, table_b.sse)),
                              "note": "spl_process(bam_path): file image -> pinned -> H2D, BGZF inflate + BAM parse on the device, then the same path as e2e; "
                                      "synthetic BAM without SEQ/QUAL (records only), so the file is much smaller than a sequencer's"}
        except Exception as ex:                                      # never lose the main line over the extra measurement
            out["bam_e2e"] = {"error": repr(ex)}
        # ---- the whole `process` command: BAM + BED12 + GFF files -> .SpliSER.tsv on disk (what a SpliSER user runs)
        try:
            out["cli_e2e"] = cli_process_run(ctx, w, bam, args, reads_rank)
        except Exception as ex:
            out["cli_e2e"] = {"error": repr(ex)}
    sampler.stop()
    if rank == 0 and not args.no_cpu_baseline and world == 1:
        rec_s, junc_s = region_sample(w, args.cpu_sample)
        dt = min(cpu_port_run(rec_s, junc_s, n_chrom, w.flags, ncores) for _ in range(2))
        out["cpu_baseline"] = {"value": len(rec_s) / dt, "unit": "reads/s", "cores": ncores, "kind": "port",
                               "sample": "first %d records of %s (genomic sub-region, same coverage) + its %d junctions; C port of the reference algorithm, OpenMP over sites" % (len(rec_s), w.chroms[int(rec_s.seg_chrom[0])], len(junc_s)),
                               "python_reference": python_reference_note()}
    elif rank == 0:
        out["cpu_baseline"] = None
    if rank == 0:
        emit(json.dumps(out))
    ctx.close()
    ranks.close()


if __name__ == "__main__":
    main()
