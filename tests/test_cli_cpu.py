"""The host layer of the SpliSER-compatible CLI on the CPU: BED parsing, locus filters, gene assignment, TSV writers and the
`combine` merge driver (order dependence, gap collection) run for real; only the counting calls behind the C ABI are
replaced by a stand-in context that answers them with the oracle (test infrastructure -- the product never does this).
Outputs must be byte-identical to what the unmodified reference wrote (golden fixtures).  The same tests with the CUDA
path behind the ABI are in test_cli_gpu.py."""
import os

import pytest

from common import load_golden
from oracle import c_oracle
from spliser_b200 import Records, api


class OracleContext:
    """Answers process_bam / recount_bam like spliser_b200.Context, with the C oracle instead of the GPU."""

    def process_bam(self, bam_path, chrom_names, junctions, flags):
        rec = Records.from_bam(bam_path, chrom_names)
        return api.SiteTable(**c_oracle.process(rec, len(chrom_names), junctions, flags))

    def process_records(self, records, n_chrom, junctions, flags):
        return api.SiteTable(**c_oracle.process(records, n_chrom, junctions, flags))

    def recount_bam(self, bam_path, chrom_names, gaps, flags):
        rec = Records.from_bam(bam_path, chrom_names)
        return c_oracle.recount(rec, len(chrom_names), gaps, flags)

    def recount_records(self, records, n_chrom, gaps, flags):
        return c_oracle.recount(records, n_chrom, gaps, flags)

    def close(self):
        pass


def _write_inputs(tmp, case):
    bed = os.path.join(tmp, "j.bed")
    open(bed, "w").write(case["bed"])
    reads = [tuple(r) for r in case["reads"]]
    refs = sorted({r[0] for r in reads}) or ["C"]
    bam = os.path.join(tmp, "x.bam")
    Records.from_reads(refs, reads).write_bam(bam, refs)
    gff = None
    if case.get("gff"):
        gff = os.path.join(tmp, "a.gff")
        open(gff, "w").write(case["gff"])
    return bam, bed, gff


def test_process_cli_text_on_cpu(tmp_path, built_library):
    from spliser_b200 import cli
    cases = load_golden("appendix_a.json.gz")["process"] + load_golden("process_fuzz.json.gz")[::9]
    ctx = OracleContext()
    for k, case in enumerate(cases):
        d = tmp_path / ("p%d" % k)
        d.mkdir()
        bam, bed, gff = _write_inputs(str(d), case)
        out = str(d / "out")
        cli.process(bam, bed, out, qGene=case.get("qgene", "All"), qChrom=case.get("qchrom", "All"),
                    maxIntronSize=case.get("max_intron", 0), annotationFile=gff, isStranded=case["stranded"],
                    strandedType=case["stype"], isbeta2Cryptic=case["cryptic"], ctx=ctx)
        assert open(out + ".SpliSER.tsv").read() == case["tsv"], case.get("name", case.get("seed"))


def test_combine_merge_driver_on_cpu(tmp_path, built_library):
    from spliser_b200 import cli
    cases = load_golden("appendix_a.json.gz")["combine"] + load_golden("combine_fuzz.json.gz")
    ctx = OracleContext()
    for k, case in enumerate(cases):
        d = tmp_path / ("c%d" % k)
        d.mkdir()
        lines = []
        for i, s in enumerate(case["samples"]):
            p = str(d / ("s%d.SpliSER.tsv" % i))
            open(p, "w").write(s["tsv"])
            bam = str(d / ("s%d.bam" % i))
            Records.from_reads(["C"], [tuple(r) for r in s["reads"]]).write_bam(bam, ["C"])
            lines.append("%s\t%s\t%s\n" % (s["title"], p, bam))
        sf = str(d / "samples.tsv")
        open(sf, "w").writelines(lines)
        out = str(d / "out")
        cli.combine(sf, out, isStranded=case["stranded"], strandedType=case["stype"], ctx=ctx)
        assert open(out + ".combined.tsv").read() == case["combined"], case.get("name", case.get("seed"))


def _run_wide_combine(cli, ctx, case, d, **combine_kw):
    """One combine_wide golden case through cli.process (per sample, with the annotation) and cli.combine."""
    lines = []
    gff = str(d / "a.gff")
    open(gff, "w").write(case["gff"])
    for i, s in enumerate(case["samples"]):
        reads = [tuple(r) for r in s["reads"]]
        bam = str(d / ("s%d.bam" % i))
        Records.from_reads(case["chroms"], reads).write_bam(bam, case["chroms"])
        bed = str(d / ("s%d.bed" % i))
        open(bed, "w").write(s["bed"])
        out = str(d / ("s%d" % i))
        cli.process(bam, bed, out, annotationFile=gff, isStranded=case["stranded"],
                    strandedType=case["stype"] if case["stranded"] else None, isbeta2Cryptic=case["cryptic"], ctx=ctx)
        assert open(out + ".SpliSER.tsv").read() == s["tsv"], (case["seed"], i)
        lines.append("%s\t%s\t%s\n" % (s["title"], out + ".SpliSER.tsv", bam))
    sf = str(d / "samples.tsv")
    open(sf, "w").writelines(lines)
    out = str(d / "out")
    if combine_kw:
        ctx = None
    cli.combine(sf, out, qGene=case["qgene"], isStranded=case["stranded"], strandedType=case["stype"],
                isbeta2Cryptic=case["cryptic"], ctx=ctx, **combine_kw)
    assert open(out + ".combined.tsv").read() == case["combined"], case["seed"]


def test_combine_over_regions_genes_and_cryptic_on_cpu(tmp_path, built_library):
    """Several regions (some absent from a sample), annotation + -g filter, --beta2Cryptic (str(float) column)."""
    from spliser_b200 import cli
    ctx = OracleContext()
    for k, case in enumerate(load_golden("combine_wide.json.gz")):
        d = tmp_path / ("w%d" % k)
        d.mkdir()
        _run_wide_combine(cli, ctx, case, d)


def test_combine_sharded_by_sample_over_contexts_on_cpu(tmp_path, built_library):
    """The re-count of `combine` sharded by sample: three stand-in contexts on three host threads, samples dealt
    round-robin (dist.samples_of); the combined table does not depend on who re-counted which sample."""
    from spliser_b200 import cli
    made = []

    def factory(device):
        made.append(device)
        return OracleContext()
    cases = [c for c in load_golden("combine_wide.json.gz") if len(c["samples"]) >= 3][:12]
    assert cases
    for k, case in enumerate(cases):
        d = tmp_path / ("m%d" % k)
        d.mkdir()
        _run_wide_combine(cli, OracleContext(), case, d, devices=[0, 1, 2], context_factory=factory)
    assert set(made) <= {0, 1, 2} and len(set(made)) >= 2


def run_shallow_cases(cli, ctx, tmp_path, stride=1):
    """combine_shallow golden cases (oracle/make_golden.py --shallow): cli.combineShallow on the samples of the referenced
    combine_wide / combine_fuzz case -> list of (case id, got text, wanted text) that differ."""
    cases = load_golden("combine_shallow.json.gz")[::stride]
    bases, dirs, bad = {}, {}, []
    for n, g in enumerate(cases):
        if g["base"] is None:                                 # crafted known-answer case: samples stored inline
            base, key = g, ("crafted", g["name"])
            g = dict(g, index=g["name"])
        else:
            base, key = bases.setdefault(g["base"], load_golden(g["base"]))[g["index"]], (g["base"], g["index"])
        if key not in dirs:                                   # the sample files of a base case are written once
            d = tmp_path / ("b%d" % len(dirs))
            d.mkdir()
            chroms = base.get("chroms", ["C"])
            lines = []
            for i, smp in enumerate(base["samples"]):
                p = str(d / ("s%d.SpliSER.tsv" % i))
                open(p, "w").write(smp["tsv"])
                bam = str(d / ("s%d.bam" % i))
                Records.from_reads(chroms, [tuple(r) for r in smp["reads"]]).write_bam(bam, chroms)
                lines.append("%s\t%s\t%s\n" % (smp["title"], p, bam))
            open(str(d / "samples.tsv"), "w").writelines(lines)
            dirs[key] = d
        d = dirs[key]
        out = str(d / ("out%d" % n))
        cli.combineShallow(str(d / "samples.tsv"), out, qGene=g["qgene"], isStranded=base["stranded"], minSamples=g["min_samples"],
                           minReads=g["min_reads"], minSSE=g["min_sse"], strandedType=base["stype"], isbeta2Cryptic=g["cryptic"], ctx=ctx)
        got = open(out + ".combined.tsv").read()
        if got != g["combined"]:
            bad.append(((g["base"], g["index"], g["min_samples"], g["min_reads"], g["min_sse"], g["qgene"]), got, g["combined"]))
    return len(cases), bad


def test_combine_shallow_equals_the_reference_files_on_cpu(tmp_path, built_library):
    """combineShallow (S:920-1167): 400 runs of the unmodified reference over a grid of -m / -r / -e settings, with and without
    -g and --beta2Cryptic, on the samples of the combine goldens; the merge driver's shallow mode must write the same bytes."""
    from spliser_b200 import cli
    n, bad = run_shallow_cases(cli, OracleContext(), tmp_path)
    assert n >= 400
    assert not bad, "%d/%d differ; first: %r\n--- got\n%s\n--- want\n%s" % (len(bad), n, bad[0][0], bad[0][1][:1500], bad[0][2][:1500])


def test_c4_shaped_process_and_combine_equal_the_reference_files_on_cpu(tmp_path, built_library):
    """configs[3] shape at reduced size (6 samples of one genome, 1.2M records, 14k sites each, 11k re-counted gaps): the
    unmodified reference wrote the six .SpliSER.tsv and the .combined.tsv in the authoring container (oracle/c4_shape.py);
    the CLI must write the same bytes from BAM / BED12 files."""
    import json
    from oracle import c4_shape
    from spliser_b200 import cli
    gold = json.load(open(c4_shape.GOLDEN))
    per_sample, combined, shallow = c4_shape.run_cli(cli, OracleContext(), str(tmp_path))
    assert per_sample == gold["process_sha256"]
    assert combined == gold["combined_sha256"]
    assert shallow == gold["shallow_sha256"]                  # combineShallow -m 4 -r 6 -e 0.05 over the same samples


def test_c4_with_48_samples_equals_the_reference_files_on_cpu(tmp_path, built_library):
    """configs[3]'s own sample count: 48 samples (6 conditions x 8 replicates) of one genome, 1.44M records in all; the
    unmodified reference wrote 48 .SpliSER.tsv, the 210,576-row .combined.tsv (24,131 re-counted gaps) and the
    combineShallow -m 24 -r 4 -e 0.05 table (oracle/c4_shape.py c4x48); the CLI must write the same bytes."""
    import json
    from oracle import c4_shape
    from spliser_b200 import cli
    shape = c4_shape.FORTY_EIGHT
    gold = json.load(open(shape.golden))
    assert len(gold["titles"]) == 48
    per_sample, combined, shallow = c4_shape.run_cli(cli, OracleContext(), str(tmp_path), shape)
    assert per_sample == gold["process_sha256"]
    assert combined == gold["combined_sha256"]
    assert shallow == gold["shallow_sha256"]
    # the same merge with the re-count sharded by sample over 8 devices: one context per device on its own host
    # thread, six samples each; the bytes do not depend on the sharding
    made = []

    def factory(device):
        made.append(device)
        return OracleContext()
    out = str(tmp_path / "sharded")
    cli.combine(str(tmp_path / "samples.tsv"), out, isStranded=True, strandedType="rf", devices=list(range(8)), context_factory=factory)
    assert sorted(made) == list(range(8))
    assert c4_shape.sha(open(out + ".combined.tsv", "rb").read()) == gold["combined_sha256"]


def test_annotated_process_equals_the_reference_files_on_cpu(tmp_path, built_library):
    """configs[1]'s flow with the GFF annotation at reduced size (400k records, 10k sites, ~1.8k overlapping / nested genes on
    both strands): stranded, unstranded, the annotation with its lines shuffled, -g / -c / -m, and -t mRNA; the unmodified
    reference wrote the five .SpliSER.tsv (oracle/c2_annotated.py), the CLI must write the same bytes -- Gene column included."""
    import json
    from oracle import c2_annotated
    from spliser_b200 import cli
    gold = json.load(open(c2_annotated.GOLDEN))["variants"]
    got = c2_annotated.run_cli(cli, OracleContext(), str(tmp_path))
    assert sorted(got) == sorted(gold)
    for name, digest in got.items():
        assert digest == gold[name]["sha256"], name


def test_cli_errors_mirror_the_reference(tmp_path, built_library):
    from spliser_b200 import cli
    case = [c for c in load_golden("appendix_a.json.gz")["process"] if c["name"] == "A.4-locus"][0]
    bam, bed, gff = _write_inputs(str(tmp_path), case)
    with pytest.raises(AttributeError):                       # -g gene absent from the annotation (S:283)
        cli.process(bam, bed, str(tmp_path / "o"), qGene="NOPE", qChrom="C", maxIntronSize=50, annotationFile=gff, ctx=OracleContext())
    with pytest.raises(UnboundLocalError):                    # --isStranded with a type other than fr / rf (S:378-406)
        cli.process(bam, bed, str(tmp_path / "o"), isStranded=True, strandedType="xx", ctx=OracleContext())


def test_command_line_mirrors_the_reference_options(tmp_path, built_library, monkeypatch):
    """Sub-commands and option names of S:1300-1361 (process / combine / combineShallow), the two argument checks of S:1349-1355,
    and the dispatch with the parsed values."""
    from spliser_b200 import cli
    with pytest.raises(SystemExit):                           # --isStranded requires -s (S:1354-1355)
        cli.main(["combineShallow", "-S", "s.tsv", "-o", "out", "--isStranded"])
    with pytest.raises(SystemExit):                           # --gene requires --annotationFile (S:1350-1353)
        cli.main(["process", "-B", "x.bam", "-b", "x.bed", "-o", "out", "-g", "G1"])
    seen = {}
    monkeypatch.setattr(cli, "combineShallow", lambda **kw: seen.update(kw))
    cli.main(["combineShallow", "-S", "s.tsv", "-o", "out", "-g", "G7", "-m", "3", "-r", "12", "-e", "0.25", "--isStranded", "-s", "rf",
              "--beta2Cryptic", "--gpus", "2"])
    assert seen == dict(samplesFile="s.tsv", outputPath="out", qGene="G7", isStranded=True, minSamples=3, minReads=12, minSSE=0.25,
                        strandedType="rf", isbeta2Cryptic=True, devices=[0, 1])
    seen.clear()
    cli.main(["combineShallow", "-S", "s.tsv", "-o", "out"])  # the reference's defaults: -m 0, -r 10, -e 0.0
    assert (seen["minSamples"], seen["minReads"], seen["minSSE"], seen["qGene"], seen["strandedType"]) == (0, 10, 0.0, "All", None)

