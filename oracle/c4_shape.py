"""TEST INFRASTRUCTURE ONLY -- never imported by the product package.

BASELINE configs[3] at reduced size, driven through the unmodified reference: 6 samples (3 conditions x 2 replicates) of one
synthetic genome -- every sample has its own reads and its own junction subset (a condition drops a share of the junctions,
a replicate a few more), so the merged table has gaps in every sample -- are each run through `process`, then all through
`combine` (re-count of the sites a sample lacks, SpliSER_v0_1_8.py:742-917), stranded rf.

    python oracle/c4_shape.py        # authoring container: runs the reference, writes tests/golden/c4_shape_reference.json
    python oracle/c4_shape.py c4x48  # the same with configs[3]'s own sample count, 48 = 6 conditions x 8 replicates, on a smaller
                                     # genome -> tests/golden/c4_48_samples_reference.json

The golden holds the sha256 of every `.SpliSER.tsv` and of the `.combined.tsv` the reference wrote; the CLI tests rebuild the
same inputs (`build_samples`, deterministic), run spliser_b200.cli on BAM / BED12 files and compare the bytes' digests
(tests/test_cli_cpu.py with the oracle-backed stand-in context, tests/test_cli_gpu.py with the CUDA path).
"""
from __future__ import annotations

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

N_COND, N_REP = 3, 2
N_RECORDS = 1_200_000
SEED = 20260004
SHALLOW = (4, 6, 0.05)            # combineShallow -m / -r / -e of the second merge
GOLDEN = os.path.join(ROOT, "tests", "golden", "c4_shape_reference.json")
CONTIGS = (("R1", 4_000_000), ("R2", 2_500_000), ("R3", 1_200_000))


class Shape:
    """One configs[3]-shaped workload: conditions x replicates samples of one synthetic genome."""

    def __init__(self, name, n_cond, n_rep, n_records, seed, shallow, contigs, golden):
        self.name, self.n_cond, self.n_rep, self.n_records, self.seed = name, n_cond, n_rep, n_records, seed
        self.shallow, self.contigs, self.golden = shallow, contigs, golden


SIX = Shape("c4", N_COND, N_REP, N_RECORDS, SEED, SHALLOW, CONTIGS, GOLDEN)
# the sample count configs[3] names (6 conditions x 8 replicates = 48 samples), on a smaller genome so that the reference's
# merge (literal_eval of every cell of 48 tables) and the CPU test stay in the tens of seconds
FORTY_EIGHT = Shape("c4x48", 6, 8, 1_440_000, 20260048, (24, 4, 0.05), (("R1", 1_500_000), ("R2", 900_000)),
                    os.path.join(ROOT, "tests", "golden", "c4_48_samples_reference.json"))
SHAPES = {s.name: s for s in (SIX, FORTY_EIGHT)}


def _subset(r, keep):
    from spliser_b200 import Records
    ncig = np.diff(r.cig_off.astype(np.int64))
    off = np.zeros(int(keep.sum()) + 1, np.uint32)
    np.cumsum(ncig[keep], out=off[1:])
    csum = np.concatenate([[0], np.cumsum(keep)])
    return Records(r.pos[keep], r.flag[keep], off, r.cigar[np.repeat(keep, ncig)], r.seg_chrom, csum[r.seg_off])


def build_samples(shape=SIX):
    """-> (chromosome names, chromosome lengths, [(title, Records, BED12 text)] in samples-file order)."""
    from spliser_b200 import Junctions, synth
    from spliser_b200.synth import Workload
    w = synth.generate(synth.SynthConfig(name="c4", seed=shape.seed, contigs=shape.contigs, n_records=shape.n_records,
                                         read_len=100, paired=True, stranded=True, genes_per_mb=165.0))
    rng = np.random.default_rng(shape.seed)
    r, j = w.records, w.junctions
    N_COND, N_REP = shape.n_cond, shape.n_rep
    n_s = N_COND * N_REP
    owner = rng.integers(0, n_s, len(r))                       # every record belongs to one sample
    cond_drop = rng.random((N_COND, len(j))) < 0.08
    out = []
    for c in range(N_COND):
        for rep in range(N_REP):
            k = c * N_REP + rep
            keep_j = ~cond_drop[c] & (rng.random(len(j)) >= 0.04)
            score = np.maximum(1, (j.score[keep_j] + k) // n_s)
            jk = Junctions(j.chrom[keep_j], j.left[keep_j], j.right[keep_j], score, j.strand[keep_j])
            wk = Workload(w.cfg, w.chroms, w.chrom_len, _subset(r, owner == k), jk)
            out.append(("cond%d_rep%d" % (c + 1, rep + 1), wk.records, wk.bed12_text()))
    return w.chroms, w.chrom_len, out


def sha(text) -> str:
    return hashlib.sha256(text.encode() if isinstance(text, str) else text).hexdigest()


def run_cli(cli, ctx, tmp, shape=SIX):
    """The product CLI on files: `process` per sample, then `combine` -> ([digest of each .SpliSER.tsv], digest of .combined.tsv)."""
    chroms, chrom_len, samples = build_samples(shape)
    SHALLOW = shape.shallow
    lines, digests = [], []
    for i, (title, rec, bed_text) in enumerate(samples):
        bam, bed, out = os.path.join(tmp, "s%d.bam" % i), os.path.join(tmp, "s%d.bed" % i), os.path.join(tmp, "s%d" % i)
        rec.write_bam(bam, chroms, chrom_len)
        with open(bed, "w") as fh:
            fh.write(bed_text)
        cli.process(bam, bed, out, isStranded=True, strandedType="rf", ctx=ctx)
        digests.append(sha(open(out + ".SpliSER.tsv", "rb").read()))
        lines.append("%s\t%s\t%s\n" % (title, out + ".SpliSER.tsv", bam))
    sf = os.path.join(tmp, "samples.tsv")
    with open(sf, "w") as fh:
        fh.writelines(lines)
    out = os.path.join(tmp, "combined")
    cli.combine(sf, out, isStranded=True, strandedType="rf", ctx=ctx)
    out2 = os.path.join(tmp, "shallow")
    cli.combineShallow(sf, out2, isStranded=True, minSamples=SHALLOW[0], minReads=SHALLOW[1], minSSE=SHALLOW[2], strandedType="rf", ctx=ctx)
    return digests, sha(open(out + ".combined.tsv", "rb").read()), sha(open(out2 + ".combined.tsv", "rb").read())


def main(shape=SIX):
    from oracle import ref_runner
    from oracle.time_reference import IndexedStore, _Popen
    from spliser_b200.synth import Workload
    if not ref_runner.reference_available():
        raise SystemExit("reference not mounted at %s" % ref_runner.REF_DIR)
    chroms, chrom_len, samples = build_samples(shape)
    SHALLOW, N_COND, N_REP, N_RECORDS, GOLDEN = shape.shallow, shape.n_cond, shape.n_rep, shape.n_records, shape.golden
    stores, tsvs, secs = {}, [], []
    old_argv, old_out = sys.argv, sys.stdout
    with tempfile.TemporaryDirectory() as td:
        lines = []
        for i, (title, rec, bed_text) in enumerate(samples):
            bam = "s%d.bam" % i
            store = IndexedStore(Workload(None, chroms, chrom_len, rec, None))
            stores[bam] = store
            mod = ref_runner.load_reference(ref_runner.ReadStore())
            mod.subprocess.Popen = lambda args, stdout=None, _s=store, **kw: _Popen(_s, args)
            bed = os.path.join(td, "s%d.bed" % i)
            with open(bed, "w") as fh:
                fh.write(bed_text)
            out = os.path.join(td, "s%d" % i)
            sys.argv, sys.stdout = ["SpliSER", "process"], io.StringIO()
            t0 = time.perf_counter()
            try:
                mod.process(bam, bed, out, "All", "All", 0, None, "gene", True, "rf", False)
            finally:
                sys.argv, sys.stdout = old_argv, old_out
            secs.append(time.perf_counter() - t0)
            tsvs.append(open(out + ".SpliSER.tsv").read())
            lines.append("%s\t%s\t%s\n" % (title, out + ".SpliSER.tsv", bam))
            print("process %s: %d records, %d rows, %.1f s" % (title, len(rec), tsvs[-1].count("\n") - 1, secs[-1]), flush=True)
        sf = os.path.join(td, "samples.tsv")
        with open(sf, "w") as fh:
            fh.writelines(lines)
        mod = ref_runner.load_reference(ref_runner.ReadStore())
        mod.subprocess.Popen = lambda args, stdout=None, **kw: _Popen(stores[args[2]], args)
        n_gap = [0]
        orig = mod.checkBam

        def counted(*a, **kw):
            n_gap[0] += 1
            return orig(*a, **kw)
        mod.checkBam = counted
        out = os.path.join(td, "combined")
        sys.argv, sys.stdout = ["SpliSER", "combine"], io.StringIO()
        t0 = time.perf_counter()
        try:
            mod.combine(sf, out, "All", True, "rf", False)
        finally:
            sys.argv, sys.stdout = old_argv, old_out
        t_comb = time.perf_counter() - t0
        combined = open(out + ".combined.tsv").read()
        n_comb_gaps = n_gap[0]
        mod = ref_runner.load_reference(ref_runner.ReadStore())              # module state is global: a fresh copy for the second merge
        mod.subprocess.Popen = lambda args, stdout=None, **kw: _Popen(stores[args[2]], args)
        orig2 = mod.checkBam
        n_gap2 = [0]

        def counted2(*a, **kw):
            n_gap2[0] += 1
            return orig2(*a, **kw)
        mod.checkBam = counted2
        out2 = os.path.join(td, "shallow")
        sys.argv, sys.stdout = ["SpliSER", "combineShallow"], io.StringIO()
        t0 = time.perf_counter()
        try:
            mod.combineShallow(sf, out2, "All", True, SHALLOW[0], SHALLOW[1], SHALLOW[2], "rf", False)
        finally:
            sys.argv, sys.stdout = old_argv, old_out
        t_shallow = time.perf_counter() - t0
        shallow = open(out2 + ".combined.tsv").read()
    print("combine: %d rows, %d re-counted gaps, %.1f s" % (combined.count("\n") - 1, n_comb_gaps, t_comb))
    print("combineShallow -m %d -r %d -e %g: %d rows, %d re-counted gaps, %.1f s" % (SHALLOW + (shallow.count("\n") - 1, n_gap2[0], t_shallow)))
    doc = {"made_by": "oracle/c4_shape.py%s (unmodified reference, authoring container)" % ("" if shape is SIX else " " + shape.name),
           "workload": "configs[3] shape at reduced size: %d samples (%d conditions x %d replicates) of one synthetic genome (%s), %d records in all, stranded rf"
                       % (len(samples), N_COND, N_REP, ", ".join(chroms), N_RECORDS),
           "titles": [s[0] for s in samples], "records": [len(s[1]) for s in samples],
           "process_rows": [t.count("\n") - 1 for t in tsvs], "process_sha256": [sha(t) for t in tsvs],
           "combined_rows": combined.count("\n") - 1, "recounted_gaps": n_comb_gaps, "combined_sha256": sha(combined),
           "shallow_settings": {"minSamples": SHALLOW[0], "minReads": SHALLOW[1], "minSSE": SHALLOW[2]},
           "shallow_rows": shallow.count("\n") - 1, "shallow_recounted_gaps": n_gap2[0], "shallow_sha256": sha(shallow),
           "reference_seconds": {"process_per_sample": [round(x, 2) for x in secs], "combine": round(t_comb, 2), "combineShallow": round(t_shallow, 2)}}
    with open(GOLDEN, "w") as fh:
        json.dump(doc, fh, indent=1)
        fh.write("\n")


if __name__ == "__main__":
    main(SHAPES[sys.argv[1]] if len(sys.argv) > 1 else SIX)
