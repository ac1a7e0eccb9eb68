"""TEST INFRASTRUCTURE ONLY -- never imported by the product package.

BASELINE configs[1]'s flow at reduced size, WITH the GFF annotation (gene assignment, SpliSER_v0_1_8.py:76-173 and S:313-329),
driven through the unmodified reference: one stranded-rf sample of a three-contig synthetic genome and a dense annotation
(a gene every ~3 kb on either strand, lengths up to 7 kb, so genes overlap and nest -- the cases where the reference's
bisection over gene starts has its quirks).  Five runs: all genes stranded; all genes unstranded; the same annotation with
its lines shuffled (createGenes inserts by bisection, S:93-110); `-g GENE -c CHROM -m 20000` (the window filter); and `-t mRNA` (no effect in the reference, S:82).

    python oracle/c2_annotated.py        # authoring container: runs the reference, writes tests/golden/c2_annotated_reference.json

The golden holds the sha256 of each `.SpliSER.tsv` the reference wrote; the CLI tests rebuild the same inputs
(`build_inputs`, deterministic) and compare the digests (tests/test_cli_cpu.py with the oracle-backed stand-in context,
tests/test_zz_added_after_gpu_budget.py with the CUDA path).
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

SEED = 20260021
N_RECORDS = 400_000
CONTIGS = (("A1", 3_000_000), ("A2", 1_800_000), ("A3", 700_000))
GOLDEN = os.path.join(ROOT, "tests", "golden", "c2_annotated_reference.json")


def gff_lines(chroms, chrom_len):
    rng = np.random.default_rng(SEED + 1)
    lines = []
    for c, n in zip(chroms, chrom_len):
        k = max(1, int(n) // 3000)
        starts = np.sort(rng.integers(1, max(2, int(n) - 7000), size=k))
        lens = rng.integers(300, 7000, size=k)
        strand = rng.integers(0, 2, size=k)
        for i, (a, b, d) in enumerate(zip(starts.tolist(), lens.tolist(), strand.tolist())):
            lines.append("%s\tsynthetic\tgene\t%d\t%d\t.\t%s\t.\tID=%s_G%05d;Name=n%d\n" % (c, a, a + b, "+-"[d], c, i, i))
            if i % 3 == 0:                                        # non-gene features are ignored for aType "gene" (S:83)
                lines.append("%s\tsynthetic\tmRNA\t%d\t%d\t.\t%s\t.\tID=%s_G%05d.1\n" % (c, a, a + b, "+-"[d], c, i))
    return lines


def build_inputs():
    """-> (Workload, {variant name: (gff text, kwargs of process)})."""
    from spliser_b200 import synth
    w = synth.generate(synth.SynthConfig(name="c2a", seed=SEED, contigs=CONTIGS, n_records=N_RECORDS, read_len=150, paired=True,
                                         stranded=True, genes_per_mb=165.0))
    lines = gff_lines(w.chroms, w.chrom_len)
    rng = np.random.default_rng(SEED + 2)
    shuffled = [lines[i] for i in rng.permutation(len(lines))]
    genes = [ln.split("\t") for ln in lines if ln.split("\t")[2] == "gene"]
    j = w.junctions                                                # the query gene: the one holding most same-strand junction starts
    best, pick = -1, genes[0]
    for g in genes[::7]:
        c = w.chroms.index(g[0])
        n_in = int(np.count_nonzero((j.chrom == c) & (j.left >= int(g[3])) & (j.left <= int(g[4])) & (j.strand == ord(g[6]))))
        if n_in > best:
            best, pick = n_in, g
    qgene, qchrom = pick[8].split(";")[0].split("=", 1)[1], pick[0]
    text = "##gff-version 3\n" + "".join(lines)
    variants = {
        "stranded": (text, dict(isStranded=True, strandedType="rf")),
        "unstranded": (text, dict(isStranded=False, strandedType=None)),
        "shuffled_annotation": ("".join(shuffled), dict(isStranded=True, strandedType="rf")),
        "query_gene": (text, dict(isStranded=True, strandedType="rf", qGene=qgene, qChrom=qchrom, maxIntronSize=20000)),
        "atype_mRNA": (text, dict(isStranded=True, strandedType="rf", aType="mRNA")),         # -t mRNA: the reference tests line.type == 'gene' whatever -t says (S:82)
    }
    return w, variants


def sha(text) -> str:
    return hashlib.sha256(text.encode() if isinstance(text, str) else text).hexdigest()


def run_cli(cli, ctx, tmp):
    """The product CLI on files for every variant -> {variant: digest of its .SpliSER.tsv}."""
    w, variants = build_inputs()
    bam, bed = os.path.join(tmp, "s.bam"), os.path.join(tmp, "s.bed")
    w.records.write_bam(bam, w.chroms, w.chrom_len)
    with open(bed, "w") as fh:
        fh.write(w.bed12_text())
    out = {}
    for name, (gff_text, kw) in variants.items():
        gff = os.path.join(tmp, name + ".gff")
        with open(gff, "w") as fh:
            fh.write(gff_text)
        o = os.path.join(tmp, name)
        cli.process(bam, bed, o, annotationFile=gff, ctx=ctx, **kw)
        out[name] = sha(open(o + ".SpliSER.tsv", "rb").read())
    return out


def main():
    from oracle import ref_runner
    from oracle.time_reference import IndexedStore, _Popen
    if not ref_runner.reference_available():
        raise SystemExit("reference not mounted at %s" % ref_runner.REF_DIR)
    w, variants = build_inputs()
    store = IndexedStore(w)
    doc = {"made_by": "oracle/c2_annotated.py (unmodified reference, authoring container)",
           "workload": "configs[1] flow at reduced size with the GFF annotation: %d records, stranded rf, contigs %s" % (len(w.records), ", ".join(w.chroms)),
           "variants": {}}
    old_argv, old_out = sys.argv, sys.stdout
    with tempfile.TemporaryDirectory() as td:
        bed = os.path.join(td, "s.bed")
        with open(bed, "w") as fh:
            fh.write(w.bed12_text())
        for name, (gff_text, kw) in variants.items():
            gff = os.path.join(td, name + ".gff")
            with open(gff, "w") as fh:
                fh.write(gff_text)
            mod = ref_runner.load_reference(ref_runner.ReadStore())
            mod.subprocess.Popen = lambda args, stdout=None, _s=store, **k2: _Popen(_s, args)
            o = os.path.join(td, name)
            sys.argv, sys.stdout = ["SpliSER", "process"], io.StringIO()
            t0 = time.perf_counter()
            try:
                mod.process("s.bam", bed, o, kw.get("qGene", "All"), kw.get("qChrom", "All"), kw.get("maxIntronSize", 0), gff, kw.get("aType", "gene"),
                            kw["isStranded"], kw["strandedType"], False)
            finally:
                sys.argv, sys.stdout = old_argv, old_out
            dt = time.perf_counter() - t0
            tsv = open(o + ".SpliSER.tsv").read()
            rows = tsv.count("\n") - 1
            named = sum(1 for ln in tsv.splitlines()[1:] if ln.split("\t")[3] not in ("NA", ""))
            print("%s: %d rows, %d with a gene, %.1f s" % (name, rows, named, dt), flush=True)
            doc["variants"][name] = {"rows": rows, "rows_with_gene": named, "sha256": sha(tsv), "reference_seconds": round(dt, 2),
                                     "options": {k: v for k, v in kw.items()}}
    with open(GOLDEN, "w") as fh:
        json.dump(doc, fh, indent=1)
        fh.write("\n")


if __name__ == "__main__":
    main()
