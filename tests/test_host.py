"""CPU tests of the host side: the C ABI library loads and exports every symbol the header declares,
BED12 parsing mirrors the reference's text filters, BAM write/read round-trips (no GPU needed)."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(built_library):
    import ctypes
    from spliser_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "spliser_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(spl_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 35
    lib = ctypes.CDLL(built_library)
    missing = [n for n in sorted(declared) if not hasattr(lib, n)]
    assert not missing, missing
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert b"spliser_b200" in _lib.load().spl_version()


def test_no_gpu_means_loud_failure(built_library):
    from conftest import HAS_GPU
    import spliser_b200
    if HAS_GPU:
        pytest.skip("GPU present")
    with pytest.raises(spliser_b200.SpliserError, match="no CPU fallback"):
        spliser_b200.Context(0)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "spliser_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                hit = re.search(r"(import|from)\s+\.*oracle|oracle[/.]\w|liboracle|spliser_oracle", txt)
                assert hit is None, "%s reaches into oracle/: %r" % (f, hit.group(0))


def test_bed_parse_matches_oracle_parse():
    from oracle import fuzzgen
    from oracle import spliser_oracle as O
    from spliser_b200.bed import parse_bed12
    for seed in range(50):
        case = fuzzgen.gen_case(seed, n_chrom=2)
        lines = case["bed"].splitlines(True)
        for kw in (dict(), dict(qchrom="C1"), dict(qchrom="C0", qgene_bounds=(300, 500), max_intron=100)):
            chroms, junc = O.parse_bed(lines, **kw)
            c2, j2, sstr = parse_bed12(lines, **kw)
            assert chroms == c2
            got = [(int(j2.chrom[i]), int(j2.left[i]), int(j2.right[i]), int(j2.score[i]), sstr[i]) for i in range(len(j2))]
            assert got == junc


def test_bam_roundtrip(tmp_path, built_library):
    from spliser_b200 import Records, synth
    w = synth.generate(synth.config_small(30000, seed=5, stranded=True, paired=True))
    path = str(tmp_path / "x.bam")
    w.records.write_bam(path, w.chroms, w.chrom_len)
    back = Records.from_bam(path, w.chroms)
    for k in ("pos", "flag", "cig_off", "cigar", "seg_chrom", "seg_off"):
        assert np.array_equal(getattr(back, k), getattr(w.records, k)), k
    # a different chromosome order / unknown references are remapped or dropped
    back = Records.from_bam(path, [w.chroms[1]])
    n1 = int(w.records.seg_off[2] - w.records.seg_off[1])
    assert len(back) == n1 and list(back.seg_chrom) == [0]
    with pytest.raises(IOError):
        Records.from_bam(str(tmp_path / "missing.bam"), w.chroms)
    open(str(tmp_path / "bad.bam"), "wb").write(b"not a bam file at all, just bytes" * 4)
    with pytest.raises(IOError):
        Records.from_bam(str(tmp_path / "bad.bam"), w.chroms)
    # a member whose CRC32 field does not match its data is refused (as htslib does), and so is an ISIZE above the format's 64 KiB
    raw = bytearray(open(path, "rb").read())
    bsize = int.from_bytes(raw[16:18], "little") + 1                    # first member: BC subfield right behind the fixed header
    crc_at = bsize - 8
    raw[crc_at] ^= 0x5a
    open(str(tmp_path / "crc.bam"), "wb").write(raw)
    with pytest.raises(IOError):
        Records.from_bam(str(tmp_path / "crc.bam"), w.chroms)
    raw[crc_at] ^= 0x5a
    raw[bsize - 4:bsize] = (70000).to_bytes(4, "little")
    open(str(tmp_path / "isize.bam"), "wb").write(raw)
    with pytest.raises(IOError):
        Records.from_bam(str(tmp_path / "isize.bam"), w.chroms)


def test_records_from_reads_segments():
    from spliser_b200 import Records
    r = Records.from_reads(["A", "B"], [("A", 5, 0, "10M"), ("A", 7, 16, "5M3N5M"), ("Z", 1, 0, "4M"), ("B", 2, 0, "*")])
    assert list(r.seg_chrom) == [0, -1, 1] and list(r.seg_off) == [0, 2, 3, 4]
    assert list(r.cig_off) == [0, 1, 4, 5, 5]
    assert list(r.cigar[1:4]) == [(5 << 4) | 0, (3 << 4) | 3, (5 << 4) | 0]


def test_synth_is_deterministic_and_sorted():
    from spliser_b200 import synth
    a = synth.generate(synth.config_small(5000, seed=9))
    b = synth.generate(synth.config_small(5000, seed=9))
    assert np.array_equal(a.records.cigar, b.records.cigar) and np.array_equal(a.junctions.score, b.junctions.score)
    r = a.records
    for s in range(len(r.seg_chrom)):
        assert (np.diff(r.pos[r.seg_off[s]:r.seg_off[s + 1]]) >= 0).all()
    assert len(a.junctions) > 50 and int(((r.cigar & 15) == 3).sum()) > 500


def test_sorted_site_table_equals_line_by_line_emulation(monkeypatch, built_library):
    """Clean regime: the sort/unique builder must reproduce the emulated reference construction
    (list order, first-seen strand and row, Partners order, PartnerCounts, CompetitorPos)."""
    from oracle import fuzzgen
    from spliser_b200 import synth
    from spliser_b200.api import build_site_table
    from spliser_b200.bed import parse_bed12
    fields = ("chrom", "pos", "strand", "alpha", "first_line", "partner_off", "partner_pos", "partner_cnt", "comp_off", "comp_pos")

    def both(n_chrom, junc, flags):
        monkeypatch.delenv("SPLISER_FORCE_EMULATION", raising=False)
        a = build_site_table(n_chrom, junc, flags)
        monkeypatch.setenv("SPLISER_FORCE_EMULATION", "1")
        b = build_site_table(n_chrom, junc, flags)
        monkeypatch.delenv("SPLISER_FORCE_EMULATION", raising=False)
        for k in fields:
            assert np.array_equal(getattr(a, k), getattr(b, k)), k
        return a

    for seed in range(120):
        case = fuzzgen.gen_case(seed, n_chrom=1 + seed % 3, dirty=False)
        chroms, junc, _ = parse_bed12(case["bed"].splitlines(True))
        both(len(chroms), junc, 1 if case["stranded"] else 0)
    for stranded in (False, True):
        w = synth.generate(synth.config_small(40000, seed=70 + stranded, stranded=stranded, paired=stranded))
        t = both(len(w.chroms), w.junctions, w.flags)
        assert len(t) > 100 and int(t.alpha.sum()) == 2 * int(w.junctions.score.sum())


def test_site_table_matches_reference_golden(built_library):
    """Host-only site table (alpha, Partners, Competitors) against the reference's rows, incl. the dirty regime."""
    from common import load_golden
    from spliser_b200.api import build_site_table
    from spliser_b200.bed import parse_bed12
    cases = load_golden("process_fuzz.json.gz") + [c for c in load_golden("appendix_a.json.gz")["process"] if "qgene" not in c]
    for case in cases:
        chroms, junc, _ = parse_bed12(case["bed"].splitlines(True))
        t = build_site_table(len(chroms), junc, 1 if case["stranded"] else 0)
        got = [(chroms[int(t.chrom[i])], int(t.pos[i]), t.strand_str(i), int(t.alpha[i]), list(map(list, t.partners(i).items())), t.competitors(i))
               for i in range(len(t))]
        want = [(r["chrom"], r["pos"], r["strand"], r["alpha"], [list(p) for p in r["partners"]], r["competitors"]) for r in case["rows"]]
        assert got == want, case.get("seed", case.get("name"))


def test_gene_assignment_matches_reference_golden(tmp_path):
    """Gene column (createGenes + binary_gene_search, S:50-173) for the annotated known-answer scenario."""
    from common import load_golden
    from spliser_b200.genes import gene_name, load_annotation
    case = [c for c in load_golden("appendix_a.json.gz")["process"] if c["name"] == "A.4-annot"][0]
    p = tmp_path / "a.gff"
    p.write_text(case["gff"])
    ann = load_annotation(str(p))
    assert ann.chrom_index == ["C"] and [g.name for g in ann.genes[0]] == ["G1", "G2"]
    assert (ann.genes[0][0].left, ann.genes[0][0].right) == (89, 320)
    for r in case["rows"]:
        assert gene_name(ann, 0, r["pos"], r["strand"], False) == r["gene"]
    q = load_annotation(str(p), "G2")
    assert q.query_gene.name == "G2" and len(q.genes[0]) == 1
    assert load_annotation(str(p), "NOPE").query_gene is None


def test_device_inflate_decoder_matches_zlib_on_the_host():
    """csrc/inflate.h is the raw-DEFLATE decoder every GPU warp runs on a BGZF member; the same source compiled for
    the host must reproduce zlib byte for byte (stored, fixed and dynamic blocks, long codes) and reject corrupt input."""
    import ctypes as C
    import zlib
    import numpy as np
    from spliser_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(3)

    def inflate(comp, cap):
        src = np.frombuffer(comp, np.uint8).copy() if len(comp) else np.zeros(1, np.uint8)
        dst = np.full(cap + 8, 0xEE, np.uint8)
        n = C.c_uint32(0)
        rc = lib.spl_debug_inflate(src.ctypes.data_as(_lib.c_u8p), len(comp), dst.ctypes.data_as(_lib.c_u8p), cap, C.byref(n))
        return rc, dst[:n.value].tobytes(), dst
    n_cases = 0
    for level in (0, 1, 6, 9):
        for strategy in (zlib.Z_DEFAULT_STRATEGY, zlib.Z_FIXED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE, zlib.Z_FILTERED):
            for kind in range(6):
                n = int(rng.integers(0, 65281)) if kind else kind
                if kind == 1:
                    raw = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
                elif kind == 2:
                    raw = rng.choice(np.frombuffer(b"ACGT", np.uint8), n).tobytes()
                elif kind == 3:
                    raw = bytes(n)
                elif kind == 4:      # BAM-like: small little-endian integers
                    raw = rng.integers(0, 40, n // 4 + 1, dtype=np.int32).tobytes()[:n]
                else:
                    raw = (rng.integers(0, 256, 97, dtype=np.uint8).tobytes() * (n // 97 + 1))[:n]
                co = zlib.compressobj(level, zlib.DEFLATED, -15, 8, strategy)
                comp = co.compress(raw) + co.flush()
                rc, out, dst = inflate(comp, len(raw))
                assert rc == 0 and out == raw and dst[len(raw)] == 0xEE, (level, strategy, kind, n, rc)
                n_cases += 1
                if len(comp) > 8:    # truncated and corrupted members fail cleanly
                    rc, _, _ = inflate(comp[:len(comp) // 2], len(raw))
                    assert rc != 0
                    inflate(comp[:len(comp) // 3] + bytes([comp[len(comp) // 3] ^ 0x5A]) + comp[len(comp) // 3 + 1:], len(raw))
                if len(raw) > 1:     # output capacity is respected
                    rc, _, dst = inflate(comp, len(raw) - 1)
                    assert rc != 0 and dst[len(raw) - 1] == 0xEE
    assert n_cases == 120


def test_native_gene_search_equals_the_python_restatement(built_library):
    """spl_gene_search (csrc/host_text.cpp) against genes.binary_gene_search (the restatement pinned on the reference's
    golden rows above): overlapping / nested genes, both strands, '?' site strands, positions on the gene bounds, lists
    of 0-40 genes (the last-ditch window of S:162-169 and its skipped last gene matter for the short ones)."""
    import bisect
    import random
    from spliser_b200 import api, hosttext
    from spliser_b200.genes import Annotation, Gene, binary_gene_search
    rng = random.Random(11)
    n_checked = n_found = 0
    for trial in range(300):
        genes = []
        for k in range(rng.randint(0, 40)):
            left = rng.randrange(0, 3000, 10)
            g = Gene("C", "g%d_%d" % (trial, k), left, left + rng.choice([5, 40, 200, 900, 2500]), rng.choice("+-"))
            bisect.insort(genes, g)
        ann = Annotation(chrom_index=["C"], genes=[genes])
        n = 200
        edges = [g.left for g in genes] + [g.right for g in genes] or [0]
        pos = [rng.choice(edges) + rng.choice([-1, 0, 0, 1]) if rng.random() < 0.4 else rng.randrange(0, 6000) for _ in range(n)]
        pos = [max(p, 0) for p in pos]
        strands = [rng.choice(["+", "-", "?", ".", ""]) for _ in range(n)]
        for stranded in (False, True):
            vocab, sid = hosttext.strand_ids(strands)
            z = np.zeros(n, np.int64)
            table = api.SiteTable(chrom=np.zeros(n, np.int32), pos=np.array(pos, np.int32), strand=np.zeros(n, np.uint8),
                                  alpha=z, beta1=z, beta2simple=z, beta2cryptic=z, beta2weighted=np.zeros(n), sse=np.zeros(n),
                                  first_line=np.arange(n, dtype=np.int64), partner_off=np.zeros(n + 1, np.int64),
                                  partner_pos=np.zeros(0, np.int32), partner_cnt=np.zeros(0, np.int64),
                                  comp_off=np.zeros(n + 1, np.int64), comp_pos=np.zeros(0, np.int32))
            names, got = hosttext.assign_genes(ann, table, sid, vocab, stranded)
            want = [binary_gene_search(genes, p, s, stranded) for p, s in zip(pos, strands)]
            assert got.tolist() == want, (trial, stranded)
            n_checked += n
            n_found += sum(w >= 0 for w in want)
    assert n_checked == 300 * 2 * 200 and n_found > n_checked // 4


def test_combined_tsv_prints_floats_like_python(tmp_path, built_library):
    """beta2_weighted of a .combined.tsv is str(float(text)) (S:734): shortest round-trip digits, exponent form below 1e-4
    and from 1e16.  One sample, no gaps, so nothing needs the GPU."""
    from spliser_b200 import cli
    texts = ["0.00000", "0.00001", "0.00010", "0.00123", "0.10000", "0.33333", "1.00000", "2.50000", "123456.78901",
             "1e-7", "5e-324", "1.7976931348623157e308", "1e15", "1e16", "12345678901234567890", "0.1", "3", "1e22", "1.5e-5"]
    rows = ["Region\tSite\tStrand\tGene\tSSE\talpha_count\tbeta1_count\tbeta2Simple_count\tbeta2Cryptic_count\t"
            "beta2Cryptic_weighted\tPartners\tCompetitors\n"]
    for i, t in enumerate(texts):
        rows.append("C\t%d\t+\tNA\t0.500\t%d\t2\t1\t7\t%s\t{%d: 3, 5000: 1}\t[%d, 7000]\n" % (100 + i, i, t, 900 + i, 800 + i))
    p = tmp_path / "a.SpliSER.tsv"
    p.write_text("".join(rows))
    sf = tmp_path / "samples.tsv"
    sf.write_text("A\t%s\tnone.bam\n" % p)
    cli.combine(str(sf), str(tmp_path / "out"), isbeta2Cryptic=True)
    got = open(str(tmp_path / "out") + ".combined.tsv").read().splitlines()[1:]
    assert len(got) == len(texts)
    for i, (line, t) in enumerate(zip(got, texts)):
        v = line.split("\t")
        w = float(t)
        den = i + (2 + 1 + w)
        assert v[10] == str(w), (t, v[10])
        assert v[5] == "{0:.3f}".format(i / den if den > 0.0 else 0.0)
        assert v[11] == str({900 + i: 3, 5000: 1}) and v[12] == str([800 + i, 7000])


def test_combine_rejects_malformed_tables(tmp_path, built_library):
    from spliser_b200 import SpliserError, cli
    p = tmp_path / "bad.SpliSER.tsv"
    p.write_text("header\nC\t100\t+\tNA\t0.5\t1\t2\n")
    sf = tmp_path / "samples.tsv"
    sf.write_text("A\t%s\tnone.bam\n" % p)
    with pytest.raises(SpliserError, match="fewer than 12"):
        cli.combine(str(sf), str(tmp_path / "out"))
    sf.write_text("A\t%s\tnone.bam\n" % (tmp_path / "absent.tsv"))
    with pytest.raises(IOError):
        cli.combine(str(sf), str(tmp_path / "out"))
    sf.write_text("A\tonly-two-columns\n")
    with pytest.raises(Exception, match="exactly 3"):
        cli.combine(str(sf), str(tmp_path / "out"))


def _bed_reference_semantics(lines, chrom_index=(), qchrom="All", bounds=None, max_intron=0):
    """findAlphaCounts' text half (S:255-288) line by line, as plain Python: the checker for spl_bed_parse."""
    chroms, out = list(chrom_index), []
    for line in lines:
        v = str(line).split("\t")
        if len(v) != 12:
            continue
        if v[0] not in chroms:
            chroms.append(v[0])
        if not (qchrom == v[0] or qchrom == "All"):
            continue
        flank = v[10].split(",")
        left, right, score = int(v[1]) + int(flank[0]), int(v[2]) - int(flank[1]), int(v[4])
        if bounds is not None:
            lin = (left + max_intron >= bounds[0]) and (left <= bounds[1])
            rin = (right - max_intron <= bounds[1]) and (right >= bounds[0])
            if not (lin or rin):
                continue
        out.append((chroms.index(v[0]), left, right, score, v[5]))
    return chroms, out


def test_native_bed_parser_edge_cases(tmp_path, built_library):
    from spliser_b200.bed import parse_bed12

    def row(chrom, start, end, score, strand, sizes="10,10", tail="0,90"):
        return "\t".join([chrom, str(start), str(end), "J", str(score), strand, str(start), str(end), "255,0,0", "2", sizes, tail]) + "\n"
    lines = [
        "track name=junctions description=\"x\"\n",                       # not 12 fields
        row("chrB", 90, 310, 5, "+"),
        row("chrA", 90, 210, 3, "-"),
        row("chrB", 190, 410, " 7 ", "?", sizes=" 12 , 8 ,"),               # int() strips blanks; regtools' trailing comma
        row("chrC", 5, 60, "+4", ""),                                      # explicit sign, empty strand column
        row("chrB", 90, 310, 2, "+-odd"),                                  # multi-character strand text is kept verbatim
        row("chrA", 1000, 1300, 1, ".").rstrip("\n") + "\textra\n",        # 13 fields: skipped
        "chrD\t1\t2\n",                                                    # short line
        row("chrD", 400, 900, 9, "+", tail="0,490").rstrip("\n"),          # last line without a newline
    ]
    for kw in (dict(), dict(qchrom="chrB"), dict(qchrom="chrZ"), dict(chrom_index=["chrA", "chrQ"]),
               dict(qchrom="chrB", bounds=(250, 320), max_intron=0), dict(bounds=(250, 320), max_intron=100)):
        want_chroms, want = _bed_reference_semantics(lines, kw.get("chrom_index", ()), kw.get("qchrom", "All"), kw.get("bounds"), kw.get("max_intron", 0))
        chroms, j, sstr = parse_bed12(lines, kw.get("chrom_index"), kw.get("qchrom", "All"), kw.get("bounds"), kw.get("max_intron", 0))
        got = [(int(j.chrom[i]), int(j.left[i]), int(j.right[i]), int(j.score[i]), sstr[i]) for i in range(len(j))]
        assert chroms == want_chroms and got == want, kw
        assert [int(b) for b in j.strand] == [(ord(s[0]) if s else 0) for *_, s in want]
    # the same bytes through a file handle, with CRLF line ends (text mode translates them before the reference sees them)
    p = tmp_path / "crlf.bed"
    p.write_bytes("".join(lines).replace("\n", "\r\n").encode())
    with open(p) as fh:
        chroms, j, sstr = parse_bed12(fh)
    assert (chroms, len(j)) == (_bed_reference_semantics(lines)[0], len(_bed_reference_semantics(lines)[1]))
    with pytest.raises(ValueError):
        parse_bed12([row("c", "x1", 60, 1, "+")])                          # int('x1') raises in the reference too
    with pytest.raises(ValueError):
        parse_bed12([row("c", 5, 60, 1, "+", sizes="10")])                 # flankSize[1] does not exist
    with pytest.raises(OverflowError):
        parse_bed12([row("c", 2 ** 31, 2 ** 31 + 50, 1, "+")])
    assert len(parse_bed12([])[1]) == 0 and parse_bed12("")[0] == []


def test_native_combine_parser_edge_cases(tmp_path, built_library):
    """Rows as the merge loop reads them (S:836: rstrip() then split on tabs): CRLF line ends, trailing blanks, extra
    columns, a dict text with a repeated key, NA cryptic columns; two samples with and without the site."""
    from spliser_b200 import cli
    hdr = "Region\tSite\tStrand\tGene\tSSE\talpha_count\tbeta1_count\tbeta2Simple_count\tbeta2Cryptic_count\tbeta2Cryptic_weighted\tPartners\tCompetitors"
    a = [hdr, "C\t100\t+\tg1\t0.500\t4\t3\t1\tNA\tNA\t{300: 4, 300: 5}\t[250]\textra\tcolumns   ",
         "C\t300\t+\tg1\t1.000\t4\t0\t0\tNA\tNA\t{100: 4}\t[]"]
    b = [hdr, "C\t100\t+\tg1\t0.250\t1\t3\t0\tNA\tNA\t{ 250 : 1 }\t[ 300 ]\t"]
    (tmp_path / "a.tsv").write_bytes(("\r\n".join(a) + "\r\n").encode())
    (tmp_path / "b.tsv").write_bytes(("\n".join(b)).encode())                 # no final newline

    class NoReads:
        def recount_bam(self, bam, names, gaps, flags):
            assert names == ["C"] and [g[1] for g in gaps] == [300] and [g[3] for g in gaps] == [[100]]
            return np.array([7]), np.array([2])

        def close(self):
            pass
    sf = tmp_path / "samples.tsv"
    sf.write_text("A\t%s\tx.bam\nB\t%s\ty.bam\n" % (tmp_path / "a.tsv", tmp_path / "b.tsv"))
    cli.combine(str(sf), str(tmp_path / "out"), ctx=NoReads())
    got = open(str(tmp_path / "out") + ".combined.tsv").read().splitlines()[1:]
    assert got == ["A\tC\t100\t+\tg1\t0.500\t4\t3\t1\tNA\tNA\t{300: 5, 250: 0}\t[250, 300]",
                   "B\tC\t100\t+\tg1\t0.250\t1\t3\t0\tNA\tNA\t{300: 0, 250: 1}\t[250, 300]",
                   "A\tC\t300\t+\tg1\t1.000\t4\t0\t0\tNA\tNA\t{100: 4}\t[]",
                   "B\tC\t300\t+\tg1\t0.000\t0\t7\t2\tNA\tNA\t{100: 0}\t[]"]


def test_native_bed_parser_in_pieces(built_library):
    """A BED12 image of several MB is cut at line ends and parsed concurrently; the merged result (chromosome index and
    strand-text ids by first appearance, row order, the line number of the first bad line) equals one sequential pass."""
    import random
    from spliser_b200.bed import parse_bed12
    rng = random.Random(4)
    chrom_pool = ["chr%d" % i for i in range(1, 40)]
    lines = []
    for i in range(70000):
        c = chrom_pool[min(len(chrom_pool) - 1, (i * len(chrom_pool)) // 70000 + rng.choice([0, 0, 0, 1]))]
        s = rng.randrange(1, 10 ** 8)
        e = s + rng.randrange(60, 5000)
        strand = rng.choice(["+", "-", "?", ".", "+x", "st%d" % (i // 9000)])
        lines.append("%s\t%d\t%d\tJUNC%08d\t%d\t%s\t%d\t%d\t255,0,0\t2\t%d,%d\t0,%d\n"
                     % (c, s, e, i, rng.randrange(1, 999), strand, s, e, rng.randrange(8, 30), rng.randrange(8, 30), e - s - 8))
        if i % 5000 == 17:
            lines.append("# comment line %d\n" % i)
    text = "".join(lines)
    assert len(text) > 4 * (1 << 20)
    for kw in (dict(), dict(qchrom="chr7"), dict(chrom_index=["chr30", "zz"], bounds=(10 ** 7, 2 * 10 ** 7), max_intron=50000)):
        want_chroms, want = _bed_reference_semantics(lines, kw.get("chrom_index", ()), kw.get("qchrom", "All"), kw.get("bounds"), kw.get("max_intron", 0))
        chroms, j, sstr = parse_bed12(text, kw.get("chrom_index"), kw.get("qchrom", "All"), kw.get("bounds"), kw.get("max_intron", 0))
        assert chroms == want_chroms and len(j) == len(want)
        assert list(zip(j.chrom.tolist(), j.left.tolist(), j.right.tolist(), j.score.tolist(), list(sstr))) == want
    bad = list(lines)
    bad[61234] = bad[61234].replace("JUNC", "J").replace("\t255,0,0", "\t255,0,0", 1).split("\t")
    bad[61234][1] = "12x"
    bad[61234] = "\t".join(bad[61234])
    bad[69000] = bad[69000].replace("\t2\t", "\t2\tq", 1)
    with pytest.raises(ValueError, match="BED line 61235:"):
        parse_bed12("".join(bad))


def _create_genes_reference_semantics(text, qgene="All"):
    """createGenes (S:50-116) with HTSeq's GFF_Reader conventions restated in plain Python (iv.start = start - 1, iv.end = end,
    name = value of the first attribute): the checker for spl_genes_parse."""
    import bisect
    from spliser_b200.genes import Gene

    def _first_attribute(col9):
        # HTSeq.parse_GFF_attribute_string(..., extra_return_first_value=True), restated: quote-safe split at ';', then
        # \\s*([^\\s=]+)[\\s=]+(.*) on the first piece, one pair of enclosing quotes removed
        import re
        piece, in_quote = col9, False
        for i, ch in enumerate(col9):
            if ch == '"':
                in_quote = not in_quote
            elif ch == ";" and not in_quote:
                piece = col9[:i]
                break
        if not piece.strip():
            return "_unnamed_"
        mo = re.match(r"\s*([^\s=]+)[\s=]+(.*)", piece, re.S)
        if not mo:
            return piece.strip()
        val = mo.group(2)
        if len(val) >= 2 and val.startswith('"') and val.endswith('"'):
            val = val[1:-1]
        return val
    chrom_index, genes, query = [], [], None
    for line in text.splitlines():
        if not line.strip() or line.startswith("#"):
            continue
        f = line.split("\t")
        if len(f) < 9 or f[2] != "gene":
            continue
        g = Gene(f[0], _first_attribute(f[8]), int(f[3]) - 1, int(f[4]), f[6])
        if f[0] not in chrom_index:
            chrom_index.append(f[0])
            genes.append([])
        if qgene == "All":
            bisect.insort(genes[chrom_index.index(f[0])], g)
        elif g.name == qgene:
            query = g
            genes[chrom_index.index(f[0])].append(g)
    return chrom_index, genes, query


def test_native_annotation_parser_equals_the_python_restatement(tmp_path, built_library):
    import random
    from spliser_b200.genes import load_annotation
    rng = random.Random(21)
    for trial in range(40):
        lines = ["##gff-version 3\n", "#!comment\n"]
        chroms = ["c%d" % i for i in range(rng.randint(1, 4))]
        for k in range(rng.randint(0, 120)):
            c = rng.choice(chroms)
            start = rng.choice([10, 50, 50, 200, rng.randrange(1, 5000)])
            end = start + rng.randrange(0, 900)
            kind = rng.choice(["gene", "gene", "gene", "mRNA", "exon", "Gene", "gene "])
            name = rng.choice(["G%d" % rng.randrange(12), "AT%dG%05d" % (rng.randint(1, 5), k)])
            attr = rng.choice(["ID=%s;Name=x%d" % (name, k), "ID=%s" % name, " ID=%s ;Note=a=b" % name, 'gene_id "%s"; transcript_id "t%d";' % (name, k),
                               'gene_id  "%s"' % name, name, "Parent=p;ID=%s" % name, "ID=%s=tail;x" % name, ""])
            cols = [c, "src", kind, str(start) if rng.random() > 0.05 else " %d " % start, str(end), ".", rng.choice(["+", "-", ".", "?"]), ".", attr]
            if rng.random() < 0.1:
                cols.append("tenth column")
            if rng.random() < 0.05:
                cols = cols[:rng.randint(1, 8)]
            lines.append("\t".join(cols) + "\n")
            if rng.random() < 0.03:
                lines.append("\n")
        text = "".join(lines)
        if trial % 5 == 0:
            text = text.rstrip("\n")                                     # no newline at the end of the file
        p = tmp_path / ("a%d.gff" % trial)
        p.write_bytes((text.replace("\n", "\r\n") if trial % 7 == 3 else text).encode())
        for q in ("All", "G3", "G7", "absent"):
            want_idx, want_genes, want_q = _create_genes_reference_semantics(text, q)
            ann = load_annotation(str(p), q)
            assert ann.chrom_index == want_idx, (trial, q)
            assert ann.genes == want_genes, (trial, q)
            assert ann.query_gene == want_q, (trial, q)
    bad = tmp_path / "bad.gff"
    bad.write_text("c\ts\tgene\tten\t20\t.\t+\t.\tID=g\n")
    with pytest.raises(ValueError):
        load_annotation(str(bad))


def test_process_tsv_numbers_print_like_python(tmp_path, built_library):
    """SSE ('{:.3f}') and beta2Cryptic_weighted ('{:.5f}') of the native writer against Python's own formatting: random values,
    exact ties of the binary value (0.0625 -> '0.062', 0.5 ulp cases), large magnitudes, negative zero."""
    from spliser_b200 import api, hosttext
    rng = np.random.default_rng(8)
    vals = np.concatenate([rng.random(20000), rng.random(5000) * 1e6, rng.integers(0, 2 ** 20, 5000) / 2.0 ** 14,
                           np.array([0.0625, 0.0005, 0.0015, 0.5, 2.5, 1.0005, 0.1235, 1e15 + 0.5, 123456789.987654321, -0.0, 0.0,
                                     0.00049999999999999999, 0.9995, 0.99949999999999994, 1e-300, 7e22])])
    n = len(vals)
    z = np.zeros(n, np.int64)
    w = vals[::-1].copy()
    t = api.SiteTable(chrom=np.zeros(n, np.int32), pos=np.arange(1, n + 1, dtype=np.int32), strand=np.zeros(n, np.uint8), alpha=z, beta1=z,
                      beta2simple=z, beta2cryptic=z, beta2weighted=w, sse=vals, first_line=z, partner_off=np.zeros(n + 1, np.int64),
                      partner_pos=np.zeros(0, np.int32), partner_cnt=np.zeros(0, np.int64), comp_off=np.zeros(n + 1, np.int64),
                      comp_pos=np.zeros(0, np.int32))
    p = str(tmp_path / "n.tsv")
    hosttext.write_process_tsv(p, ["C"], t, ["+"], beta2_cryptic=True)
    rows = open(p).read().splitlines()[1:]
    assert len(rows) == n
    for i, line in enumerate(rows):
        f = line.split("\t")
        assert f[4] == "{0:.3f}".format(float(vals[i])) and f[9] == "{0:.5f}".format(float(w[i])), (i, vals[i], f[4], f[9])


def test_bed_parsed_alone_then_appended_equals_parse_after_the_annotation(built_library):
    """cli.process parses the BED12 while the annotation loads and appends the BED's chromosome numbers to the annotation's
    afterwards: same chromosome index (S:90-92 then S:265-268) and junction table as parsing with the annotation's index."""
    import types
    from spliser_b200.bed import parse_bed12
    from spliser_b200.cli import _append_bed_chroms
    rng = np.random.default_rng(8)
    names = ["Chr%d" % i for i in range(1, 9)] + ["ChrC", "scaffold_12"]
    for trial in range(60):
        ann = [names[i] for i in rng.permutation(len(names))[:int(rng.integers(0, 7))]]
        rows = []
        for i in range(int(rng.integers(0, 40))):
            c = names[int(rng.integers(0, len(names)))]
            s = int(rng.integers(0, 100000))
            rows.append("%s\t%d\t%d\tJ\t%d\t%s\t0\t0\t0\t2\t10,12\t0,90" % (c, s, s + 200, int(rng.integers(1, 99)), "+-?"[int(rng.integers(0, 3))]))
            if rng.random() < 0.1:
                rows.append("%s\t1\t2\tshort line" % c)                       # not 12 columns: ignored, registers nothing
        text = "\n".join(rows) + ("\n" if rows else "")
        qchrom = "All" if trial % 3 else names[trial % len(names)]
        want_chroms, want_j, want_s = parse_bed12(text, ann, qchrom, None, 0)
        alone_chroms, alone_j, alone_s = parse_bed12(text, None, qchrom, None, 0)
        got_chroms, got_j = _append_bed_chroms(types.SimpleNamespace(chrom_index=ann), alone_chroms, alone_j)
        assert got_chroms == want_chroms, trial
        for k in ("chrom", "left", "right", "score", "strand"):
            assert np.array_equal(getattr(got_j, k), getattr(want_j, k)), (trial, k)
        assert list(alone_s) == list(want_s)


def test_compact_records_round_trip():
    """api.CompactRecords (the host layout of spl_process_compact): packing and the numpy restatement of the device's unpack
    give the plain view back -- sorted and shuffled records, a GRCh38-shaped tile (wide strides, long introns in the 32-bit
    stream), empty input; records with more than 255 operators are refused."""
    import numpy as np
    import pytest
    from spliser_b200 import CompactRecords, Records, synth
    for cfg in (synth.config_small(30_000, seed=5, stranded=True, paired=True), synth.config_c3_tile(40_000, tile=2)):
        r = synth.generate(cfg).records
        for shuffled in (False, True):
            if shuffled:
                order = np.random.default_rng(2).permutation(len(r))
                nop = np.diff(r.cig_off.astype(np.int64))[order]
                off = np.concatenate([[0], np.cumsum(nop)])
                src = np.repeat(r.cig_off[:-1].astype(np.int64)[order], nop) + (np.arange(int(off[-1])) - np.repeat(off[:-1], nop))
                r = Records(r.pos[order], r.flag[order], off, r.cigar[src], [0], [0, len(r)])
            c = CompactRecords.from_records(r)
            back = c.to_records()
            assert np.array_equal(back.pos, r.pos) and np.array_equal(back.cig_off, r.cig_off) and np.array_equal(back.cigar, r.cigar)
            assert np.array_equal(back.flag, r.flag & (1 | 16 | 64))
            v = c.view()
            assert v.n16 + v.n32 == v.n_cigar == len(r.cigar) and int(c.idx16[-1]) == v.n16 and int(c.idx32[-1]) == v.n32
            assert np.all((r.cigar[np.repeat((c.flag8 & 8) == 0, c.n_op8)] >> 4) < 4096)
    e = CompactRecords.from_records(Records.from_reads(["A"], []))
    assert len(e) == 0 and len(e.idx16) == 1 and len(e.to_records()) == 0
    big = Records.from_reads(["A"], [("A", 10, 0, "1M1I" * 130)])
    with pytest.raises(ValueError):
        CompactRecords.from_records(big)


def test_first_attribute_follows_htseq(tmp_path, built_library):
    """Gene names as HTSeq.GFF_Reader gives them (first attribute value; parse_GFF_attribute_string restated, parity unpinned):
    quote-safe split at ';', key and value separated by blanks and / or '=', one pair of enclosing quotes removed."""
    from spliser_b200.genes import load_annotation
    cases = [("ID=AT1G01010;Name=x", "AT1G01010"), ('gene_id "G1"; transcript_id "t"', "G1"), ('gene_id "A=B"; x "y"', "A=B"),
             ("ID= X;Name=n", "X"), ('gene_id "a;b"; x', "a;b"), ("ID=g=tail;x", "g=tail"), ('Name "q" ', '"q" '), ("", "_unnamed_"),
             (" ;ID=late", "_unnamed_"), ("lonely", "lonely"), ("ID=", "")]
    text = "".join("c\ts\tgene\t%d\t%d\t.\t+\t.\t%s\n" % (10 * (i + 1), 10 * (i + 1) + 5, a) for i, (a, _) in enumerate(cases))
    p = tmp_path / "a.gff"
    p.write_text(text)
    ann = load_annotation(str(p))
    got = [g.name for g in ann.genes[0]]
    assert got == [w for _, w in cases]
    want_idx, want_genes, _ = _create_genes_reference_semantics(text)
    assert [g.name for g in want_genes[0]] == got
