"""TEST INFRASTRUCTURE ONLY -- never imported by the product package.

Runs the UNMODIFIED reference (`/root/reference/SpliSER_v0_1_8.py`) in-process under two
shims, so that golden vectors can be generated in the authoring container:

  * an ``HTSeq`` stand-in exposing ``GFF_Reader`` (the reference imports HTSeq at module
    top, SpliSER_v0_1_8.py:11, and uses only ``GFF_Reader`` at :81);
  * a ``subprocess.Popen`` stand-in that answers ``samtools view <bam> chr:beg-end``
    (SpliSER_v0_1_8.py:422) from an in-memory read store with htslib's region-overlap rule
    (0-based half-open region ``[beg-1, end)``; a record overlaps when
    ``pos0 < end and pos0 + max(reflen, 1) > beg - 1``).

`/root/reference` does not exist on the GPU box: this module is used only by
`oracle/make_golden.py` and by CPU tests that skip when the reference is absent.
Neither samtools nor HTSeq is installed here, so parity is pinned on the reference's own
arithmetic executed under these two stand-ins ("Oracle").
"""
from __future__ import annotations

import importlib.util
import io
import os
import re
import sys
import tempfile
import types
from dataclasses import dataclass, field

REF_DIR = os.environ.get("SPLISER_REFERENCE_DIR", "/root/reference")
REF_MAIN = os.path.join(REF_DIR, "SpliSER_v0_1_8.py")

_CIG = re.compile(r"(\d+)([MIDNSHP=X])")


def reference_available() -> bool:
    return os.path.isfile(REF_MAIN)


def cigar_reflen(cigar: str) -> int:
    if cigar == "*":
        return 0
    return sum(int(n) for n, op in _CIG.findall(cigar) if op in "MDN=X")


@dataclass
class ReadStore:
    """bam path -> chrom -> list of (pos1, flag, cigar) in file order."""

    bams: dict = field(default_factory=dict)

    def add(self, bam: str, chrom: str, pos1: int, flag: int, cigar: str) -> None:
        self.bams.setdefault(bam, {}).setdefault(chrom, []).append((int(pos1), int(flag), cigar))

    def view(self, bam: str, region: str):
        chrom, rng = region.rsplit(":", 1)
        beg_s, end_s = rng.split("-")
        beg0, end = int(beg_s) - 1, int(end_s)
        for pos1, flag, cigar in self.bams.get(bam, {}).get(chrom, ()):
            pos0 = pos1 - 1
            rl = cigar_reflen(cigar)
            if rl == 0:
                rl = 1
            if pos0 < end and pos0 + rl > beg0:
                yield ("r\t%d\t%s\t%d\t255\t%s\t*\t0\t0\t*\t*\n" % (flag, chrom, pos1, cigar)).encode("ascii")


class _FakePopen:
    def __init__(self, store: ReadStore, args):
        assert args[0] == "samtools" and args[1] == "view", args
        self.stdout = store.view(args[2], args[3])


class _FakeSubprocess(types.SimpleNamespace):
    pass


def _htseq_shim() -> types.ModuleType:
    """HTSeq.GFF_Reader semantics used by the reference (SpliSER_v0_1_8.py:81-87):
    iv.start = GFF start - 1, iv.end = GFF end, name = value of the first attribute."""
    mod = types.ModuleType("HTSeq")

    class _IV:
        __slots__ = ("chrom", "start", "end", "strand")

    class _Feat:
        __slots__ = ("type", "name", "iv")

    def GFF_Reader(path):
        with open(path) as fh:
            for line in fh:
                if not line.strip() or line.startswith("#"):
                    continue
                f = line.rstrip("\n").split("\t")
                if len(f) < 9:
                    continue
                feat = _Feat()
                feat.type = f[2]
                iv = _IV()
                iv.chrom, iv.start, iv.end, iv.strand = f[0], int(f[3]) - 1, int(f[4]), f[6]
                feat.iv = iv
                # HTSeq.parse_GFF_attribute_string(..., extra_return_first_value=True), restated from its published source:
                # quote-safe split at ';', \s*([^\s=]+)[\s=]+(.*) on the first piece, one pair of enclosing quotes removed
                piece, in_quote = f[8], False
                for i, ch in enumerate(f[8]):
                    if ch == '"':
                        in_quote = not in_quote
                    elif ch == ";" and not in_quote:
                        piece = f[8][:i]
                        break
                if not piece.strip():
                    feat.name = "_unnamed_"
                else:
                    mo = re.match(r"\s*([^\s=]+)[\s=]+(.*)", piece, re.S)
                    val = mo.group(2) if mo else piece.strip()
                    if mo and len(val) >= 2 and val.startswith('"') and val.endswith('"'):
                        val = val[1:-1]
                    feat.name = val
                yield feat

    mod.GFF_Reader = GFF_Reader
    return mod


_counter = [0]


def load_reference(store: ReadStore):
    """Fresh copy of the reference module (its state lives in module globals,
    SpliSER_v0_1_8.py:23-47, so every run needs a new module object)."""
    if not reference_available():
        raise RuntimeError("reference not mounted at %s" % REF_DIR)
    sys.modules["HTSeq"] = _htseq_shim()
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    _counter[0] += 1
    spec = importlib.util.spec_from_file_location("_spliser_ref_%d" % _counter[0], REF_MAIN)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    fake = _FakeSubprocess()
    fake.PIPE = -1
    fake.Popen = lambda args, stdout=None, **kw: _FakePopen(store, args)
    mod.subprocess = fake
    return mod


def _site_dump(mod, sample=0):
    rows = []
    for ci, chrom in enumerate(mod.chrom_index):
        if ci >= len(mod.site2D_array):
            continue
        for s in mod.site2D_array[ci]:
            rows.append(dict(
                chrom=chrom, pos=s.getPos(), strand=s.getStrand(), gene=s.getGeneName(),
                alpha=int(s.alphaCounts[sample]), beta1=int(s.beta1Counts[sample]),
                beta2s=int(s.beta2SimpleCounts[sample]), beta2c=int(s.beta2CrypticCounts[sample]),
                beta2w=float(s.beta2Weighted[sample]).hex(), sse=float(s.SSEs[sample]).hex(),
                partners=[[int(k), int(v[sample])] for k, v in s.PartnerCounts.items()],
                competitors=[int(c) for c in s.CompetitorPos],
            ))
    return rows


def run_process(bed_text: str, reads, *, stranded=False, stype=None, cryptic=False,
                qchrom="All", qgene="All", max_intron=0, gff_text=None, bam_name="x.bam"):
    """reads: iterable of (chrom, pos1, flag, cigar). Returns (tsv_text, site_rows)."""
    store = ReadStore()
    for chrom, pos1, flag, cigar in reads:
        store.add(bam_name, chrom, pos1, flag, cigar)
    mod = load_reference(store)
    old_argv, old_out = sys.argv, sys.stdout
    with tempfile.TemporaryDirectory() as td:
        bed = os.path.join(td, "j.bed")
        with open(bed, "w") as fh:
            fh.write(bed_text)
        gff = None
        if gff_text is not None:
            gff = os.path.join(td, "a.gff")
            with open(gff, "w") as fh:
                fh.write(gff_text)
        out = os.path.join(td, "out")
        sys.argv = ["SpliSER", "process"]
        sys.stdout = io.StringIO()
        try:
            mod.process(bam_name, bed, out, qgene, qchrom, max_intron, gff, "gene",
                        stranded, stype, cryptic)
        finally:
            sys.argv, sys.stdout = old_argv, old_out
        with open(out + ".SpliSER.tsv") as fh:
            tsv = fh.read()
    return tsv, _site_dump(mod)


def run_combine(samples, *, stranded=False, stype="fr", cryptic=False, qgene="All", shallow=None):
    """samples: list of (title, spliser_tsv_text, reads) in samples-file order.
    shallow: None = `combine`; (minSamples, minReads, minSSE) = `combineShallow` (SpliSER_v0_1_8.py:920).
    Returns (.combined.tsv text, log of every checkBam gap call with its inputs and results)."""
    store = ReadStore()
    mod_store_names = []
    with tempfile.TemporaryDirectory() as td:
        lines = []
        for i, (title, tsv, reads) in enumerate(samples):
            bam = "s%d.bam" % i
            mod_store_names.append(bam)
            for chrom, pos1, flag, cigar in reads:
                store.add(bam, chrom, pos1, flag, cigar)
            p = os.path.join(td, "s%d.SpliSER.tsv" % i)
            with open(p, "w") as fh:
                fh.write(tsv)
            lines.append("%s\t%s\t%s\n" % (title, p, bam))
        sf = os.path.join(td, "samples.tsv")
        with open(sf, "w") as fh:
            fh.writelines(lines)
        mod = load_reference(store)
        gap_log = []
        orig_check = mod.checkBam

        def logged_check(bam, site, sample, is_stranded, stranded_type):
            rec = dict(sample=sample, chrom=site.getChromosome(), pos=site.getPos(), strand=site.getStrand(),
                       partners=[int(k) for k in site.getPartnerCounts().keys()],
                       competitors=[int(c) for c in site.getCompetitorPos()])
            orig_check(bam, site, sample, is_stranded, stranded_type)
            rec["beta1"] = int(site.getBeta1Count(sample))
            rec["beta2s"] = int(site.getBeta2SimpleCount(sample))
            gap_log.append(rec)

        mod.checkBam = logged_check
        out = os.path.join(td, "out")
        old_argv, old_out = sys.argv, sys.stdout
        sys.argv = ["SpliSER", "combine" if shallow is None else "combineShallow"]
        sys.stdout = io.StringIO()
        try:
            if shallow is None:
                mod.combine(sf, out, qgene, stranded, stype, cryptic)
            else:
                mod.combineShallow(sf, out, qgene, stranded, shallow[0], shallow[1], shallow[2], stype, cryptic)
        finally:
            sys.argv, sys.stdout = old_argv, old_out
        with open(out + ".combined.tsv") as fh:
            return fh.read(), gap_log


def bed_line(chrom, l, r, score, strand, anchor=10, name="J"):
    """BED12 line for junction (l, r) in the reference's site convention
    (SpliSER_v0_1_8.py:274-276: leftpos = start + blockSizes[0], rightpos = end - blockSizes[1])."""
    start, end = l - anchor, r + anchor
    return "\t".join(map(str, [chrom, start, end, name, score, strand, start, end, "255,0,0", 2,
                               "%d,%d" % (anchor, anchor), "0,%d" % (r - l + anchor)])) + "\n"
