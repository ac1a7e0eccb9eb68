"""TEST INFRASTRUCTURE ONLY -- ctypes wrapper of oracle/liboracle.so (spliser_oracle.c).
Takes the same numpy containers as the product API (Records, Junctions) but shares no code path
with libspliser_b200.so."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liboracle.so")

_i32p, _i64p, _u8p, _u16p, _u32p, _f64p = (C.POINTER(t) for t in (C.c_int32, C.c_int64, C.c_uint8, C.c_uint16, C.c_uint32, C.c_double))


class RecView(C.Structure):
    _fields_ = [("n_rec", C.c_int64), ("n_cigar", C.c_int64), ("pos", _i32p), ("flag", _u16p), ("cig_off", _u32p),
                ("cigar", _u32p), ("n_seg", C.c_int32), ("seg_chrom", _i32p), ("seg_off", _i64p)]


class OracleResult(C.Structure):
    _fields_ = [("n_sites", C.c_int64), ("chrom", _i32p), ("pos", _i32p), ("strand", _u8p), ("first_line", _i64p),
                ("alpha", _i64p), ("beta1", _i64p), ("beta2s", _i64p), ("beta2c", _i64p), ("beta2w", _f64p), ("sse", _f64p),
                ("pc_off", _i64p), ("pc_pos", _i32p), ("pc_cnt", _i64p), ("cp_off", _i64p), ("cp_pos", _i32p)]


_lib = None


def build(force=False):
    src = os.path.join(HERE, "spliser_oracle.c")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.run(["make", "-s", "-C", HERE, "-B"], check=True)
    return LIB


def load():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB)
        _lib.oracle_process.restype = C.c_int
        _lib.oracle_recount.restype = C.c_int
    return _lib


def _p(a, t):
    return a.ctypes.data_as(t)


def _view(rec):
    v = RecView()
    v.n_rec, v.n_cigar = len(rec.pos), len(rec.cigar)
    v.pos, v.flag, v.cig_off, v.cigar = _p(rec.pos, _i32p), _p(rec.flag, _u16p), _p(rec.cig_off, _u32p), _p(rec.cigar, _u32p)
    v.n_seg = len(rec.seg_chrom)
    v.seg_chrom, v.seg_off = _p(rec.seg_chrom, _i32p), _p(rec.seg_off, _i64p)
    return v


def process(rec, n_chrom, junc, flags, threads=0):
    """-> dict of numpy arrays (same fields as spliser_b200.SiteTable)."""
    lib = load()
    v = _view(rec)
    out = OracleResult()
    rc = lib.oracle_process(C.byref(v), C.c_int32(n_chrom), C.c_int64(len(junc)), _p(junc.chrom, _i32p), _p(junc.left, _i32p),
                            _p(junc.right, _i32p), _p(junc.score, _i64p), _p(junc.strand, _u8p), C.c_uint32(flags),
                            C.c_int(threads), C.byref(out))
    assert rc == 0
    n = out.n_sites

    def arr(p, cnt, dt):
        return np.ctypeslib.as_array(p, shape=(cnt,)).astype(dt, copy=True) if cnt else np.zeros(0, dt)
    pc_off = np.ctypeslib.as_array(out.pc_off, shape=(n + 1,)).astype(np.int64, copy=True)
    cp_off = np.ctypeslib.as_array(out.cp_off, shape=(n + 1,)).astype(np.int64, copy=True)
    res = dict(chrom=arr(out.chrom, n, np.int32), pos=arr(out.pos, n, np.int32), strand=arr(out.strand, n, np.uint8),
               first_line=arr(out.first_line, n, np.int64), alpha=arr(out.alpha, n, np.int64), beta1=arr(out.beta1, n, np.int64),
               beta2simple=arr(out.beta2s, n, np.int64), beta2cryptic=arr(out.beta2c, n, np.int64),
               beta2weighted=arr(out.beta2w, n, np.float64), sse=arr(out.sse, n, np.float64),
               partner_off=pc_off, partner_pos=arr(out.pc_pos, int(pc_off[-1]), np.int32),
               partner_cnt=arr(out.pc_cnt, int(pc_off[-1]), np.int64),
               comp_off=cp_off, comp_pos=arr(out.cp_pos, int(cp_off[-1]), np.int32))
    lib.oracle_result_free(C.byref(out))
    return res


def recount(rec, n_chrom, gaps, flags, threads=0):
    """gaps: list of (chrom_idx, pos, strand_str, partner_positions, competitor_positions) -> (beta1, beta2s)."""
    lib = load()
    n = len(gaps)
    s_chrom = np.array([g[0] for g in gaps], np.int32)
    s_pos = np.array([g[1] for g in gaps], np.int32)
    s_strand = np.array([(ord(g[2][0]) if g[2] else 0) for g in gaps], np.uint8)
    p_off, c_off = np.zeros(n + 1, np.int64), np.zeros(n + 1, np.int64)
    pp, cp = [], []
    for i, g in enumerate(gaps):
        pp.extend(g[3]); cp.extend(g[4])
        p_off[i + 1], c_off[i + 1] = len(pp), len(cp)
    p_pos, c_pos = np.array(pp, np.int32), np.array(cp, np.int32)
    b1, b2 = np.zeros(n, np.int64), np.zeros(n, np.int64)
    v = _view(rec)
    rc = lib.oracle_recount(C.byref(v), C.c_int32(n_chrom), C.c_int64(n), _p(s_chrom, _i32p), _p(s_pos, _i32p), _p(s_strand, _u8p),
                            _p(p_off, _i64p), _p(p_pos, _i32p), _p(c_off, _i64p), _p(c_pos, _i32p), C.c_uint32(flags),
                            C.c_int(threads), _p(b1, _i64p), _p(b2, _i64p))
    assert rc == 0
    return b1, b2


TABLE_FIELDS = ("chrom", "pos", "strand", "first_line", "alpha", "beta1", "beta2simple", "beta2cryptic", "beta2weighted",
                "sse", "partner_off", "partner_pos", "partner_cnt", "comp_off", "comp_pos")


def table_dict(t):
    """spliser_b200.SiteTable -> dict with the same keys as process() above."""
    return {k: getattr(t, k) for k in TABLE_FIELDS}


def diff_tables(a, b):
    """First difference between two table dicts (floats compared bit-for-bit), or None."""
    for k in TABLE_FIELDS:
        x, y = np.asarray(a[k]), np.asarray(b[k])
        if x.shape != y.shape:
            return "%s: shape %s != %s" % (k, x.shape, y.shape)
        if x.dtype.kind == "f":
            bad = np.nonzero(x.view(np.int64) != y.view(np.int64))[0]
        else:
            bad = np.nonzero(x != y)[0]
        if len(bad):
            i = int(bad[0])
            return "%s[%d]: %r != %r (%d differing entries)" % (k, i, x[i], y[i], len(bad))
    return None
