"""The SpliSER-compatible CLI end to end on the GPU: .SpliSER.tsv and .combined.tsv must be byte-identical
to what the unmodified reference wrote for the same inputs (golden fixtures)."""
import os

import pytest

from common import load_golden

pytestmark = pytest.mark.gpu


def _write_inputs(tmp, case, chroms_extra=()):
    from spliser_b200 import Records
    bed = os.path.join(tmp, "j.bed")
    open(bed, "w").write(case["bed"])
    reads = [tuple(r) for r in case["reads"]]
    refs = sorted({r[0] for r in reads} | set(chroms_extra)) or ["C"]
    bam = os.path.join(tmp, "x.bam")
    Records.from_reads(refs, reads).write_bam(bam, refs)
    gff = None
    if case.get("gff"):
        gff = os.path.join(tmp, "a.gff")
        open(gff, "w").write(case["gff"])
    return bam, bed, gff


def test_process_cli_known_answers(ctx, tmp_path):
    from spliser_b200 import cli
    g = load_golden("appendix_a.json.gz")
    for k, case in enumerate(g["process"]):
        d = tmp_path / ("p%d" % k)
        d.mkdir()
        bam, bed, gff = _write_inputs(str(d), case)
        out = str(d / "out")
        cli.process(bam, bed, out, qGene=case.get("qgene", "All"), qChrom=case.get("qchrom", "All"),
                    maxIntronSize=case.get("max_intron", 0), annotationFile=gff, isStranded=case["stranded"],
                    strandedType=case["stype"], isbeta2Cryptic=case["cryptic"], ctx=ctx)
        assert open(out + ".SpliSER.tsv").read() == case["tsv"], case["name"]


def test_process_cli_fuzz_text(ctx, tmp_path):
    from spliser_b200 import cli
    cases = load_golden("process_fuzz.json.gz")[::6]
    for k, case in enumerate(cases):
        d = tmp_path / ("f%d" % k)
        d.mkdir()
        bam, bed, _ = _write_inputs(str(d), case)
        out = str(d / "out")
        cli.process(bam, bed, out, isStranded=case["stranded"], strandedType=case["stype"], isbeta2Cryptic=case["cryptic"], ctx=ctx)
        assert open(out + ".SpliSER.tsv").read() == case["tsv"], case["seed"]


def test_missing_query_gene_raises_like_the_reference(ctx, tmp_path):
    from spliser_b200 import cli
    case = [c for c in load_golden("appendix_a.json.gz")["process"] if c["name"] == "A.4-locus"][0]
    bam, bed, gff = _write_inputs(str(tmp_path), case)
    with pytest.raises(AttributeError):
        cli.process(bam, bed, str(tmp_path / "o"), qGene="NOPE", qChrom="C", maxIntronSize=50, annotationFile=gff, ctx=ctx)
    with pytest.raises(UnboundLocalError):
        cli.process(bam, bed, str(tmp_path / "o"), isStranded=True, strandedType="xx", ctx=ctx)


def _run_combine(ctx, tmp, case, cryptic=False):
    from spliser_b200 import Records, cli
    lines = []
    for i, s in enumerate(case["samples"]):
        p = os.path.join(tmp, "s%d.SpliSER.tsv" % i)
        open(p, "w").write(s["tsv"])
        bam = os.path.join(tmp, "s%d.bam" % i)
        reads = [tuple(r) for r in s["reads"]]
        Records.from_reads(["C"], reads).write_bam(bam, ["C"])
        lines.append("%s\t%s\t%s\n" % (s["title"], p, bam))
    sf = os.path.join(tmp, "samples.tsv")
    open(sf, "w").writelines(lines)
    out = os.path.join(tmp, "out")
    cli.combine(sf, out, isStranded=case["stranded"], strandedType=case["stype"], isbeta2Cryptic=cryptic, ctx=ctx)
    return open(out + ".combined.tsv").read()


def test_combine_cli_matches_reference(ctx, tmp_path):
    cases = load_golden("appendix_a.json.gz")["combine"] + load_golden("combine_fuzz.json.gz")
    for k, case in enumerate(cases):
        d = tmp_path / ("c%d" % k)
        d.mkdir()
        got = _run_combine(ctx, str(d), case)
        assert got == case["combined"], case.get("name", case.get("seed"))


def test_combine_over_regions_genes_and_cryptic(ctx, tmp_path):
    """combine_wide golden cases (several regions, annotation + -g, --beta2Cryptic): cli.process per sample and
    cli.combine, both byte-identical to the reference's files."""
    from spliser_b200 import cli
    from test_cli_cpu import _run_wide_combine
    for k, case in enumerate(load_golden("combine_wide.json.gz")):
        d = tmp_path / ("w%d" % k)
        d.mkdir()
        _run_wide_combine(cli, ctx, case, d)


def test_c4_shaped_process_and_combine_equal_the_reference_files(ctx, tmp_path):
    """configs[3] shape at reduced size (6 samples of one genome, 1.2M records, 14k sites each, 11k re-counted gaps): the CLI
    on BAM / BED12 files with the CUDA path must write the bytes the unmodified reference wrote (oracle/c4_shape.py)."""
    import json
    from oracle import c4_shape
    from spliser_b200 import cli
    gold = json.load(open(c4_shape.GOLDEN))
    per_sample, combined, shallow = c4_shape.run_cli(cli, ctx, str(tmp_path))
    assert per_sample == gold["process_sha256"]
    assert combined == gold["combined_sha256"]
    assert shallow == gold["shallow_sha256"]                  # combineShallow -m 4 -r 6 -e 0.05 over the same samples


def test_combine_sharded_by_sample_over_contexts(ctx, tmp_path):
    """Sample-sharded re-count with one context per host thread.  On a one-GPU box both contexts sit on device 0 (what is
    exercised is the concurrency of independent contexts); on a multi-GPU box they sit on devices 0 and 1."""
    import ctypes
    from spliser_b200 import cli
    from test_cli_cpu import _run_wide_combine
    n_dev = ctypes.c_int(0)
    ctypes.CDLL("libcuda.so.1").cuDeviceGetCount(ctypes.byref(n_dev))       # the driver is initialised (ctx fixture)
    devices = [0, 1] if n_dev.value >= 2 else [0, 0]
    cases = [c for c in load_golden("combine_wide.json.gz") if len(c["samples"]) >= 3][:10]
    for k, case in enumerate(cases):
        d = tmp_path / ("m%d" % k)
        d.mkdir()
        _run_wide_combine(cli, ctx, case, d, devices=devices)


def test_combine_shallow_matches_reference(ctx, tmp_path):
    """combineShallow (S:920-1167) with the CUDA re-count behind it: the crafted quirk cases and every second run of the
    -m / -r / -e grid the unmodified reference wrote (tests/golden/combine_shallow.json.gz), byte for byte."""
    from spliser_b200 import cli
    from test_cli_cpu import run_shallow_cases
    n, bad = run_shallow_cases(cli, ctx, tmp_path, stride=2)
    assert n >= 200
    assert not bad, "%d/%d differ; first: %r" % (len(bad), n, bad[0][0])
