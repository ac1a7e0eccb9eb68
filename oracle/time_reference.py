"""TEST INFRASTRUCTURE ONLY -- never imported by the product package.

Times the UNMODIFIED reference (`/root/reference/SpliSER_v0_1_8.py process`) on BASELINE.json configs[0] -- the
reference's own CPU-runnable case: synthetic A. thaliana Chr1-sized sample, 2M 100 bp SE reads, ~20k BED12 junctions,
unstranded -- in the authoring container, and pins the oracle against it at that size.

The reference forks `samtools view <bam> chr:t-(t+1)` once per site (SpliSER_v0_1_8.py:422); samtools is not installed
here, so the fork is answered in-process by an INDEXED read store (coordinate-sorted arrays + binary search, the work
a .bai lookup does) that yields the SAM lines htslib's overlap rule selects (oracle/ref_runner.py states the rule).
The time spent inside that stand-in is measured separately, so the figure reported for the reference is its own
Python work per site (line split, regex CIGAR walk, branch chain, findBeta2Counts, calculateSSE, TSV) -- a LOWER bound
of what a user waits for: a real run adds a process spawn, a BAM open and an index load per site.

Outputs (committed, with this script as their recipe):
  profiles/r1_reference_python_c1.json   timing of the reference, of the C port and their ratio
  tests/golden/c1_full_reference.json    digest of the reference's own per-site result for configs[0] at full size:
                                         `tests/test_oracle.py` holds the C oracle to it on the CPU and
                                         `tests/test_gpu_parity.py` the CUDA path on the B200

  tests/golden/reference_digests.json    (--shapes) full-table digests (every column, Partners / Competitors included) of the
                                         reference on the config-shaped workloads of the GPU parity tests

    python oracle/time_reference.py [--records N]
    python oracle/time_reference.py --shapes
"""
from __future__ import annotations

import argparse
import hashlib
import io
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import c_oracle, ref_runner  # noqa: E402

LETTERS = "MIDNSHP=X"


class IndexedStore:
    """Stand-in for `samtools view` on a coordinate-sorted, indexed BAM: records of one chromosome sorted by position,
    a region query = binary search over (pos, running maximum of the end) + the overlap test of htslib."""

    def __init__(self, w):
        r = w.records
        self.seconds = 0.0
        self.lines = 0
        self.calls = 0
        self.chrom = {}
        ops = r.cigar & 15
        ln = (r.cigar >> 4).astype(np.int64)
        ref = np.where((ops == 0) | (ops == 2) | (ops == 3) | (ops == 7) | (ops == 8), ln, 0)
        cs = np.concatenate([[0], np.cumsum(ref)])
        reflen = cs[r.cig_off[1:].astype(np.int64)] - cs[r.cig_off[:-1].astype(np.int64)]
        reflen = np.maximum(reflen, 1)
        for s in range(len(r.seg_chrom)):
            a, b = int(r.seg_off[s]), int(r.seg_off[s + 1])
            pos0 = r.pos[a:b].astype(np.int64) - 1
            assert np.all(np.diff(pos0) >= 0), "records must be coordinate-sorted"
            end = pos0 + reflen[a:b]
            self.chrom[w.chroms[int(r.seg_chrom[s])]] = (a, pos0, end, np.maximum.accumulate(end))
        self.rec = r
        self.text = {}

    def _cigar(self, i):
        t = self.text.get(i)
        if t is None:
            r = self.rec
            t = "".join("%d%s" % (o >> 4, LETTERS[o & 15]) for o in r.cigar[int(r.cig_off[i]):int(r.cig_off[i + 1])].tolist())
            self.text[i] = t
        return t

    def view(self, region):
        t0 = time.perf_counter()
        chrom, rng = region.rsplit(":", 1)
        beg_s, end_s = rng.split("-")
        beg0, end = int(beg_s) - 1, int(end_s)
        out = []
        ent = self.chrom.get(chrom)
        if ent is not None:
            base, pos0, rend, rmax = ent
            hi = int(np.searchsorted(pos0, end, side="left"))          # pos0 < end
            lo = int(np.searchsorted(rmax, beg0, side="right"))        # first record whose running max end > beg0
            if lo < hi:
                idx = np.nonzero(rend[lo:hi] > beg0)[0] + lo
                r = self.rec
                for k in idx.tolist():
                    i = base + k
                    out.append(("r\t%d\t%s\t%d\t255\t%s\t*\t0\t0\t*\t*\n" % (int(r.flag[i]), chrom, int(r.pos[i]), self._cigar(i))).encode("ascii"))
        self.lines += len(out)
        self.calls += 1
        self.seconds += time.perf_counter() - t0
        return out


class _Popen:
    def __init__(self, store, args):
        assert args[0] == "samtools" and args[1] == "view", args
        self.stdout = store.view(args[3])


def run_reference(w, cryptic=False):
    """-> (site rows of the reference, seconds total, seconds inside the samtools stand-in, SAM lines, sites)."""
    store = IndexedStore(w)
    mod = ref_runner.load_reference(ref_runner.ReadStore())
    mod.subprocess.Popen = lambda args, stdout=None, **kw: _Popen(store, args)
    old_argv, old_out = sys.argv, sys.stdout
    with tempfile.TemporaryDirectory() as td:
        bed = os.path.join(td, "j.bed")
        with open(bed, "w") as fh:
            fh.write(w.bed12_text())
        sys.argv = ["SpliSER", "process"]
        sys.stdout = io.StringIO()
        stranded = bool(w.flags & 1)
        t0 = time.perf_counter()
        try:
            mod.process("x.bam", bed, os.path.join(td, "out"), "All", "All", 0, None, "gene", stranded, "rf" if stranded else None, bool(cryptic))
        finally:
            sys.argv, sys.stdout = old_argv, old_out
        total = time.perf_counter() - t0
    rows = ref_runner._site_dump(mod)
    return rows, total, store.seconds, store.lines, store.calls


def digest_columns(chroms, pos, strand, alpha, beta1, beta2s, sse):
    """Order-sensitive digest of a per-site result: sha256 over the little-endian column bytes.
    strand: first byte of the strand text per site (uint8); sse: float64, compared bit for bit."""
    h = hashlib.sha256()
    for a, dt in ((chroms, np.int32), (pos, np.int32), (strand, np.uint8), (alpha, np.int64), (beta1, np.int64), (beta2s, np.int64), (sse, np.float64)):
        h.update(np.ascontiguousarray(np.asarray(a), dtype=dt).tobytes())
    return h.hexdigest()


def digest_of_table(t):
    """t: dict of arrays (oracle.c_oracle.process / c_oracle.table_dict of a spliser_b200.SiteTable)."""
    return digest_columns(t["chrom"], t["pos"], t["strand"], t["alpha"], t["beta1"], t["beta2simple"], t["sse"])


def digest_of_rows(chrom_names, rows):
    ci = {c: i for i, c in enumerate(chrom_names)}
    return digest_columns([ci[r["chrom"]] for r in rows], [r["pos"] for r in rows], [ord(r["strand"][0]) if r["strand"] else 0 for r in rows],
                          [r["alpha"] for r in rows], [r["beta1"] for r in rows], [r["beta2s"] for r in rows],
                          [float.fromhex(r["sse"]) for r in rows])


FULL_FIELDS = (("chrom", np.int32), ("pos", np.int32), ("strand", np.uint8), ("alpha", np.int64), ("beta1", np.int64),
               ("beta2simple", np.int64), ("beta2cryptic", np.int64), ("beta2weighted", np.float64), ("sse", np.float64),
               ("partner_off", np.int64), ("partner_pos", np.int32), ("partner_cnt", np.int64), ("comp_off", np.int64), ("comp_pos", np.int32))


def full_digest_of_table(t):
    """Every column of the per-site result incl. the Partners / Competitors CSR (floats bit for bit), order-sensitive."""
    h = hashlib.sha256()
    for k, dt in FULL_FIELDS:
        h.update(np.ascontiguousarray(np.asarray(t[k]), dtype=dt).tobytes())
    return h.hexdigest()


def table_of_rows(chrom_names, rows):
    """Rows dumped from the reference's Site objects (ref_runner._site_dump) -> the column layout of the oracle / product table."""
    ci = {c: i for i, c in enumerate(chrom_names)}
    p_off, c_off = [0], [0]
    pp, pc, cp = [], [], []
    for r in rows:
        for pos, cnt in r["partners"]:
            pp.append(pos); pc.append(cnt)
        cp.extend(r["competitors"])
        p_off.append(len(pp)); c_off.append(len(cp))
    return dict(chrom=[ci[r["chrom"]] for r in rows], pos=[r["pos"] for r in rows],
                strand=[ord(r["strand"][0]) if r["strand"] else 0 for r in rows], alpha=[r["alpha"] for r in rows],
                beta1=[r["beta1"] for r in rows], beta2simple=[r["beta2s"] for r in rows], beta2cryptic=[r["beta2c"] for r in rows],
                beta2weighted=[float.fromhex(r["beta2w"]) for r in rows], sse=[float.fromhex(r["sse"]) for r in rows],
                partner_off=p_off, partner_pos=pp, partner_cnt=pc, comp_off=c_off, comp_pos=cp)


def shape(name):
    """The config-shaped workloads of tests/test_gpu_parity.py::test_config_shaped_workloads_vs_c_oracle -> (SynthConfig, extra flags)."""
    from spliser_b200 import synth
    if name == "c1_full":
        return synth.config_c1(), 0
    if name == "c3_tile":
        return synth.config_c3_tile(300_000, tile=3), 0
    if name == "c5_dense_locus":
        cfg = synth.config_c5()
        cfg.n_records = 150_000
        return cfg, 4
    if name == "c5_full":                                  # configs[4] at its full size: 1M records on the dense locus, --beta2Cryptic
        return synth.config_c5(), 4
    if name in ("c2_stranded", "c2_dirty_strands"):
        return synth.config_c2(500_000), 0
    raise KeyError(name)


def shape_workload(name):
    """-> (Workload, flags).  c2_dirty_strands: the stranded configs[1] shape with every 12th BED row's strand replaced by '?'
    (what regtools writes for junctions without an XS tag): the dirty regime of SURVEY.md 8(a), where a '?' row joins whichever
    same-position site the reference's bisection lands on."""
    from spliser_b200 import Junctions, synth
    cfg, extra = shape(name)
    w = synth.generate(cfg)
    if name == "c2_dirty_strands":
        j = w.junctions
        strand = j.strand.copy()
        strand[np.random.default_rng(12).permutation(len(j))[:len(j) // 12]] = ord("?")
        w.junctions = Junctions(j.chrom, j.left, j.right, j.score, strand)
    return w, w.flags | extra


SHAPES = ("c1_full", "c3_tile", "c5_dense_locus", "c2_stranded", "c2_dirty_strands", "c5_full")


def make_shape_digests():
    """tests/golden/reference_digests.json: full-table digests of the unmodified reference on the config-shaped workloads."""
    from spliser_b200 import synth
    out = {}
    for name in SHAPES:
        w, flags = shape_workload(name)
        rows, total, in_store, lines, calls = run_reference(w, cryptic=bool(flags & 4))
        ref = table_of_rows(w.chroms, rows)
        port = c_oracle.process(w.records, len(w.chroms), w.junctions, flags, threads=0)
        d_ref, d_port = full_digest_of_table(ref), full_digest_of_table(port)
        print("%-16s %8d records %7d sites  reference %.1f s (%.1f s in the read store)  C port == reference: %s" %
              (name, len(w.records), len(rows), total, in_store, d_ref == d_port), flush=True)
        if d_ref != d_port:
            raise SystemExit("C port differs from the reference on %s: %s" % (name, c_oracle.diff_tables(
                {k: np.asarray(v) for k, v in dict(ref, first_line=port["first_line"]).items()}, port)))
        out[name] = {"records": len(w.records), "junction_rows": len(w.junctions), "sites": len(rows), "flags": int(flags), "digest": d_ref,
                     "sums": {k: int(np.sum(ref[k])) for k in ("alpha", "beta1", "beta2simple", "beta2cryptic")},
                     "reference_seconds_own_python": round(total - in_store, 2)}
    doc = {"made_by": "oracle/time_reference.py --shapes (unmodified reference, authoring container)",
           "digest_of": "sha256 over " + " ".join("%s(%s)" % (k, np.dtype(dt).name) for k, dt in FULL_FIELDS) + "; floats by their bits; site order of the reference",
           "shapes": out}
    with open(os.path.join(ROOT, "tests", "golden", "reference_digests.json"), "w") as fh:
        json.dump(doc, fh, indent=1)
        fh.write("\n")


def main():
    from spliser_b200 import synth
    if "--shapes" in sys.argv:
        if not ref_runner.reference_available():
            raise SystemExit("reference not mounted at %s" % ref_runner.REF_DIR)
        return make_shape_digests()
    ap = argparse.ArgumentParser()
    ap.add_argument("--records", type=int, default=0, help="records of the configs[0] shape (default: its full 2M)")
    ap.add_argument("--no-write", action="store_true")
    args = ap.parse_args()
    if not ref_runner.reference_available():
        raise SystemExit("reference not mounted at %s" % ref_runner.REF_DIR)
    cfg = synth.config_c1()
    if args.records:
        cfg.n_records = args.records
    w = synth.generate(cfg)
    n = len(w.records)
    rows, total, in_store, lines, calls = run_reference(w)
    ncores = os.cpu_count() or 1
    t0 = time.perf_counter()
    port = c_oracle.process(w.records, len(w.chroms), w.junctions, w.flags, threads=ncores)
    port_s = time.perf_counter() - t0
    t0 = time.perf_counter()
    c_oracle.process(w.records, len(w.chroms), w.junctions, w.flags, threads=1)
    port1_s = time.perf_counter() - t0
    d_ref, d_port = digest_of_rows(w.chroms, rows), digest_of_table(port)
    own = total - in_store
    out = {
        "what": "unmodified SpliSER_v0_1_8.py `process` under the HTSeq / samtools stand-ins, authoring container (no GPU, no samtools)",
        "workload": "configs[0]: Chr1-sized contig, %d 100 bp SE records, %d BED12 junctions, unstranded; synth seed %d" % (n, len(w.junctions), cfg.seed),
        "sites": len(rows), "samtools_calls": calls, "sam_lines_parsed": lines,
        "reference_seconds_total": round(total, 2), "seconds_inside_samtools_stand_in": round(in_store, 2),
        "reference_seconds_own_python": round(own, 2),
        "reference_reads_per_s": n / own, "reference_sites_per_s": len(rows) / own, "reference_sam_lines_per_s": lines / own,
        "cores_reference": 1,
        "c_port_seconds": {"threads_%d" % ncores: round(port_s, 4), "threads_1": round(port1_s, 4)},
        "c_port_reads_per_s": {"threads_%d" % ncores: n / port_s, "threads_1": n / port1_s},
        "c_port_over_reference": {"threads_%d" % ncores: own / port_s, "threads_1": own / port1_s},
        "digest_reference": d_ref, "digest_c_port": d_port, "c_port_equals_reference": d_ref == d_port,
        "note": "reference_seconds_own_python excludes the read fetch; a real run adds one samtools fork + BAM open + index load per site "
                "(estimated 5-30 ms each, SURVEY.md section 6), so the reference's wall time on real files is larger. "
                "bench.py's cpu_baseline / --impl reference time the C port (kind: port); divide by c_port_over_reference for the Python reference.",
    }
    print(json.dumps(out, indent=1))
    if d_ref != d_port:
        raise SystemExit("C port differs from the reference on this workload")
    if not args.no_write and not args.records:
        with open(os.path.join(ROOT, "profiles", "r1_reference_python_c1.json"), "w") as fh:
            json.dump(out, fh, indent=1)
            fh.write("\n")
        gold = {"workload": out["workload"], "config": "synth.config_c1()", "sites": len(rows), "digest": d_ref,
                "digest_of": "sha256 over chrom(int32) pos(int32) strand-byte(uint8) alpha beta1 beta2Simple(int64) SSE(float64 bits), site order of the reference",
                "sums": {"alpha": int(sum(r["alpha"] for r in rows)), "beta1": int(sum(r["beta1"] for r in rows)), "beta2simple": int(sum(r["beta2s"] for r in rows))},
                "made_by": "oracle/time_reference.py (unmodified reference, authoring container)"}
        with open(os.path.join(ROOT, "tests", "golden", "c1_full_reference.json"), "w") as fh:
            json.dump(gold, fh, indent=1)
            fh.write("\n")


if __name__ == "__main__":
    main()
