"""Mutation fuzzer for the native host code behind the C ABI (csrc/host_text.cpp, csrc/bam_io.cpp, csrc/inflate.h).

Not collected by pytest.  Meant to be run against a sanitizer build of the library (tests/tools/asan_run.sh builds one
in a scratch copy of the repo and runs this file under AddressSanitizer + UBSan): every input below is malformed on
purpose, and the only requirement is that each entry point returns (a result or an error code / exception) without
reading or writing out of bounds.  Valid-input behaviour is pinned elsewhere (tests/test_host.py, tests/test_cli_*.py).

    python tests/tools/fuzz_native.py [seconds-per-target] [seed]
"""
from __future__ import annotations

import ctypes as C
import os
import sys
import tempfile
import time
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from spliser_b200 import _lib, synth  # noqa: E402
from spliser_b200.api import Records  # noqa: E402
from spliser_b200.bed import parse_bed12  # noqa: E402
from spliser_b200.genes import load_annotation  # noqa: E402
from spliser_b200.hosttext import CombineMerge  # noqa: E402

INTERESTING = [b"", b"\t", b"\n", b"\r", b"\x00", b",", b"-", b"+", b"?", b"-1", b"0", b"2147483647", b"2147483648",
               b"99999999999999999999", b"-99999999999999999999", b" ", b"1e5", b"0x10", b"{", b"}", b"[", b"]", b":", b"'",
               b"{1: 2}", b"[1, 2]", b"{}", b"[]", b"NA", b"nan", b"inf", b"\xff\xfe", b";", b"=", b"\"", b"ID=", b"gene"]


def mutate(rng, data: bytes, rounds=None) -> bytes:
    b = bytearray(data)
    for _ in range(rounds if rounds is not None else int(rng.integers(1, 6))):
        k = int(rng.integers(0, 8))
        n = len(b)
        i = int(rng.integers(0, n + 1)) if n else 0
        if k == 0 and n:                                  # flip a byte
            b[i % n] ^= 1 << int(rng.integers(0, 8))
        elif k == 1:                                      # insert an interesting token
            b[i:i] = INTERESTING[int(rng.integers(0, len(INTERESTING)))]
        elif k == 2 and n:                                # delete a span
            j = min(n, i + int(rng.integers(1, 40)))
            del b[i:j]
        elif k == 3 and n:                                # truncate
            del b[i:]
        elif k == 4 and n:                                # duplicate a span
            j = min(n, i + int(rng.integers(1, 200)))
            b[i:i] = b[i:j]
        elif k == 5 and n:                                # replace a field-sized span by a token
            j = min(n, i + int(rng.integers(1, 8)))
            b[i:j] = INTERESTING[int(rng.integers(0, len(INTERESTING)))]
        elif k == 6 and n:                                # random bytes
            j = min(n, i + int(rng.integers(1, 16)))
            b[i:j] = rng.integers(0, 256, j - i, dtype=np.uint8).tobytes()
        elif k == 7 and n:                                # swap separators
            b[i % n:i % n + 1] = b"\t" if b[i % n] == 10 else b"\n"
    return bytes(b)


def budget(seconds):
    end = time.time() + seconds
    while time.time() < end:
        yield


def fuzz_inflate(rng, seconds):
    lib = _lib.load()
    libc = C.CDLL(None)
    libc.malloc.restype, libc.malloc.argtypes, libc.free.argtypes = C.c_void_p, [C.c_size_t], [C.c_void_p]
    seeds = []
    for level, strat in ((1, zlib.Z_DEFAULT_STRATEGY), (6, zlib.Z_FIXED), (9, zlib.Z_DEFAULT_STRATEGY), (6, zlib.Z_HUFFMAN_ONLY), (0, 0)):
        raw = rng.integers(0, 24, 4000, dtype=np.int32).tobytes()
        co = zlib.compressobj(level, zlib.DEFLATED, -15, 8, strat)
        seeds.append((co.compress(raw) + co.flush(), len(raw)))
    n = 0
    for _ in budget(seconds):
        comp, cap = seeds[int(rng.integers(0, len(seeds)))]
        m = mutate(rng, comp) if rng.random() < 0.8 else rng.integers(0, 256, int(rng.integers(0, 600)), dtype=np.uint8).tobytes()
        cap = int(rng.choice([0, 1, cap // 2, cap, cap + 100, 65536]))
        # exact-size malloc'ed buffers (Python's small-object allocator has no red zones): an out-of-bounds access of
        # the decoder lands in ASan's red zone
        src = libc.malloc(max(1, len(m)))
        dst = libc.malloc(max(1, cap))
        C.memmove(src, m, len(m))
        out = C.c_uint32(0)
        rc = lib.spl_debug_inflate(C.cast(src, _lib.c_u8p), len(m), C.cast(dst, _lib.c_u8p), cap, C.byref(out))
        libc.free(src)
        libc.free(dst)
        assert rc != 0 or out.value <= cap
        n += 1
    return n


def _bed_seed(rng):
    rows = []
    for i in range(40):
        s = int(rng.integers(0, 1 << 20))
        a, b = int(rng.integers(1, 60)), int(rng.integers(1, 60))
        e = s + a + b + int(rng.integers(1, 5000))
        rows.append("Chr%d\t%d\t%d\tJ%d\t%d\t%s\t%d\t%d\t255,0,0\t2\t%d,%d\t0,%d" % (int(rng.integers(1, 4)), s, e, i, int(rng.integers(1, 900)),
                                                                                "+-?"[int(rng.integers(0, 3))], s, e, a, b, e - s - b))
    return ("\n".join(rows) + "\n").encode()


def fuzz_bed(rng, seconds):
    seed = _bed_seed(rng)
    n = 0
    for _ in budget(seconds):
        m = mutate(rng, seed)
        kw = {}
        if rng.random() < 0.3:
            kw = dict(qchrom="Chr1")
        if rng.random() < 0.3:
            kw.update(qgene_bounds=(int(rng.integers(0, 1 << 20)), int(rng.integers(0, 1 << 21))), max_intron=int(rng.integers(0, 1 << 20)))
        try:
            parse_bed12(m, chrom_index=["Chr2"] if rng.random() < 0.5 else None, **kw)
        except (ValueError, OverflowError, UnicodeDecodeError):
            pass
        n += 1
    return n


def _gff_seed(rng, gtf):
    rows = ["##gff-version 3"]
    for i in range(30):
        s = int(rng.integers(1, 1 << 20))
        attr = 'gene_id "G%d"; gene_name "x";' % i if gtf else "ID=G%d;Name=x" % i
        rows.append("Chr%d\tsrc\t%s\t%d\t%d\t.\t%s\t.\t%s" % (int(rng.integers(1, 4)), "gene" if i % 3 else "exon", s, s + int(rng.integers(1, 9000)), "+-."[int(rng.integers(0, 3))], attr))
    return ("\n".join(rows) + "\n").encode()


def fuzz_genes(rng, seconds, tmp):
    seeds = [_gff_seed(rng, False), _gff_seed(rng, True)]
    path = os.path.join(tmp, "a.gff")
    n = 0
    for _ in budget(seconds):
        with open(path, "wb") as f:
            f.write(mutate(rng, seeds[n & 1]))
        try:
            load_annotation(path, "G%d" % int(rng.integers(0, 40)) if rng.random() < 0.5 else "All")
        except (ValueError, OverflowError, UnicodeDecodeError, KeyError, IOError, RuntimeError):
            pass
        n += 1
    return n


def _tsv_seed(rng, cryptic):
    head = "Region\tSite\tStrand\tGene\tSSE\talpha_count\tbeta1_count\tbeta2Simple_count\tbeta2Cryptic_count\tbeta2_weighted\tPartners\tCompetitors"
    rows = [head]
    pos = 100
    for i in range(40):
        pos += int(rng.integers(1, 500))
        partners = "{%s}" % ", ".join("%d: %d" % (pos + 50 * (k + 1), int(rng.integers(1, 30))) for k in range(int(rng.integers(0, 3))))
        comps = "[%s]" % ", ".join(str(pos + 7 * (k + 1)) for k in range(int(rng.integers(0, 3))))
        rows.append("Chr%d\t%d\t%s\tNA\t%.3f\t%d\t%d\t%d\t%d\t%s\t%s\t%s" % (1 + i // 20, pos, "+-"[i & 1], rng.random(), int(rng.integers(0, 50)), int(rng.integers(0, 50)),
                                                                      int(rng.integers(0, 50)), int(rng.integers(0, 9)), repr(float(rng.random())) if cryptic else "0.0", partners, comps))
    return ("\n".join(rows) + "\n").encode()


def fuzz_combine(rng, seconds, tmp):
    seeds = [_tsv_seed(rng, False), _tsv_seed(rng, True)]
    paths = [os.path.join(tmp, "s%d.SpliSER.tsv" % k) for k in range(3)]
    out = os.path.join(tmp, "out.combined.tsv")
    n = 0
    for _ in budget(seconds):
        for k, p in enumerate(paths):
            with open(p, "wb") as f:
                f.write(mutate(rng, seeds[(n + k) & 1]) if rng.random() < 0.7 else seeds[(n + k) & 1])
        try:
            with CombineMerge() as m:
                m.add_samples(["a", "b", "c"], paths, threads=int(rng.integers(0, 3)))
                names = m.region_names()
                order = list(range(len(names)))
                rng.shuffle(order)
                if n % 3 == 2:
                    m.merge_shallow(order, qgene="All" if rng.random() < 0.7 else "NA", is_stranded=bool(rng.random() < 0.5),
                                    min_samples=int(rng.integers(0, 4)), min_reads=int(rng.integers(0, 40)), min_sse=float(rng.random()))
                else:
                    m.merge(order, qgene="All", is_stranded=bool(rng.random() < 0.5))
                for k in range(3):
                    g = m.gaps(k)
                    m.set_recount(k, np.arange(len(g), dtype=np.int64), np.arange(len(g), dtype=np.int64))
                m.write(out, beta2_cryptic=bool(n & 1))
        except Exception as ex:                     # noqa: BLE001 -- any Python-level error is a clean rejection
            if isinstance(ex, (MemoryError, SystemError)):
                raise
        n += 1
    return n


def fuzz_bam(rng, seconds, tmp):
    w = synth.generate(synth.config_small(3000, seed=int(rng.integers(0, 1000)), stranded=True, paired=True))
    good = os.path.join(tmp, "g.bam")
    w.records.write_bam(good, w.chroms, w.chrom_len)
    seed = open(good, "rb").read()
    raw = b""                                         # the same records as ONE big uncompressed payload, re-packed after mutation
    off = 0
    while off < len(seed):
        bsize = int.from_bytes(seed[off + 16:off + 18], "little") + 1
        raw += zlib.decompress(seed[off + 18:off + bsize - 8], -15)
        off += bsize

    def bgzf(payload):
        out = b""
        for i in range(0, len(payload), 0xff00):
            chunk = payload[i:i + 0xff00]
            co = zlib.compressobj(1, zlib.DEFLATED, -15)
            d = co.compress(chunk) + co.flush()
            out += (b"\x1f\x8b\x08\x04\0\0\0\0\0\xff\x06\0BC\x02\0" + (len(d) + 25).to_bytes(2, "little") + d
                    + zlib.crc32(chunk).to_bytes(4, "little") + len(chunk).to_bytes(4, "little"))
        return out + bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")
    bad = os.path.join(tmp, "b.bam")
    n = 0
    for _ in budget(seconds):
        if rng.random() < 0.5:                        # container damage
            m = mutate(rng, seed)
        else:                                          # well-formed BGZF around damaged BAM records
            m = bgzf(mutate(rng, raw, rounds=int(rng.integers(1, 4))))
        with open(bad, "wb") as f:
            f.write(m)
        try:
            Records.from_bam(bad, w.chroms if rng.random() < 0.7 else w.chroms[:1], threads=int(rng.integers(0, 3)))
        except (IOError, ValueError, RuntimeError):
            pass
        n += 1
    return n


def main():
    seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 5.0
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
    with tempfile.TemporaryDirectory() as tmp:
        for name, fn in (("inflate", lambda: fuzz_inflate(rng, seconds)), ("bed", lambda: fuzz_bed(rng, seconds)),
                         ("genes", lambda: fuzz_genes(rng, seconds, tmp)), ("combine", lambda: fuzz_combine(rng, seconds, tmp)),
                         ("bam", lambda: fuzz_bam(rng, seconds, tmp))):
            t = time.time()
            n = fn()
            print("%-8s %7d inputs in %.1f s, no crash" % (name, n, time.time() - t), flush=True)


if __name__ == "__main__":
    main()
