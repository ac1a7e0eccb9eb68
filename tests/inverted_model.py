"""Executable specification of the algorithm the CUDA kernels implement ("range-add minus
exceptions"), in plain Python, so that the *design* can be checked against the oracle on a CPU
box (tests/test_inverted_model.py).  It is not the product and not the oracle: it mirrors the
device data layout (site table sorted by position per chromosome, reverse partner index,
per-class coverage / span counters, per-site exception counters) one-to-one with
spliser_b200/csrc/kernels.cu; DESIGN.md "Algorithm" describes it in prose.
"""
from __future__ import annotations

import bisect

ADV_MAPPED = "M=X"


def read_plus(flag, rf):
    first = bool(flag & 64) or not (flag & 1)
    rev = bool(flag & 16)
    plus = first != rev            # fr: '+' iff first XOR rev (S:378-390)
    return (not plus) if rf else plus


def expand(pos1, ops):
    """-> (blocks [(a,b)], junctions [(l,r)], end_exclusive) in 1-based coordinates."""
    cur = pos1
    blocks, juncs = [], []
    for n, op in ops:
        if op in ADV_MAPPED:
            blocks.append((cur, cur + n))
            cur += n
        elif op == "N":
            juncs.append((cur - 1, cur + n - 1))
            cur += n
        elif op == "D":
            cur += n
    return blocks, juncs, cur


class Graph:
    """Device-side view of one chromosome's sites: parallel arrays in list order."""

    def __init__(self, pos, cls, ppos, cpos):
        self.pos = pos              # non-decreasing
        self.cls = cls              # 0 any (unstranded) / 1 '+' / 2 '-' / 3 never matches
        self.ppos = ppos            # per site: list of partner positions (P_t)
        self.cpos = cpos            # per site: sorted competitor positions (C_t)
        self.rp = {}                # reverse partner index: position -> [site idx t with position in P_t]
        for t, pl in enumerate(ppos):
            for p in pl:
                lst = self.rp.setdefault(p, [])
                if t not in lst:
                    lst.append(t)


def count_chrom(g: Graph, reads, stranded, rf, combine, want_dc=False):
    """reads: [(pos1, flag, ops)] of this chromosome.  Returns (beta1[], beta2s_bam[], dc{(t,p):n})."""
    S = len(g.pos)
    cov = [[0] * (S + 1) for _ in range(2)]     # difference arrays in site-index space
    span = [[0] * (S + 1) for _ in range(2)]
    covx = [0] * S
    spanx = [0] * S
    flank = [0] * S
    dc = {}
    for pos1, flag, ops in reads:
        blocks, juncs, end = expand(pos1, ops)
        k = 0
        if stranded:
            k = 0 if read_plus(flag, rf) else 1
        for a, b in blocks:                                   # K3: stabbing range add
            i0 = bisect.bisect_left(g.pos, a)
            i1 = bisect.bisect_right(g.pos, b - 2)
            if i0 < i1:
                cov[k][i0] += 1
                cov[k][i1] -= 1
        if not juncs:
            continue
        for l, r in juncs:                                    # K4: span range add
            i0 = bisect.bisect_right(g.pos, l)
            i1 = bisect.bisect_left(g.pos, r)
            if i0 < i1:
                span[k][i0] += 1
                span[k][i1] -= 1

        def pair(j, t):
            l, r = juncs[j]
            P, C = g.ppos[t], g.cpos[t]
            return (l in P and r in C) or (l in C and r in P)

        read_sites = set()
        for l, r in juncs:
            read_sites.add(l)
            read_sites.add(r)
        for j, (l, r) in enumerate(juncs):                    # K4: exceptions
            for e_is_r, e in ((False, l), (True, r)):
                for t in g.rp.get(e, ()):
                    if not pair(j, t):
                        continue
                    if e_is_r and l in g.ppos[t]:
                        continue                              # already seen at (j, l)
                    if any(pair(jj, t) for jj in range(j)):
                        continue                              # handled at an earlier junction
                    tp = g.pos[t]
                    if not (pos1 <= tp <= end - 1):           # POS <= t (S:435) and region overlap (S:422)
                        continue
                    ok = (not stranded) or (g.cls[t] == 1 and k == 0) or (g.cls[t] == 2 and k == 1)
                    covers = any(a <= tp and b >= tp + 2 for a, b in blocks)
                    partner_used = None
                    alpha = False
                    for (ll, rr) in juncs:
                        if ll == tp:
                            partner_used, alpha = rr, True
                        if rr == tp:
                            partner_used, alpha = ll, True
                    kstar = None
                    for kk, (ll, rr) in enumerate(juncs):
                        if ll < tp < rr:
                            kstar = kk
                    if alpha:                                 # branch 1 (S:519-527)
                        if want_dc:
                            for p in set(g.ppos[t]) & read_sites:
                                if p != partner_used:
                                    dc[(t, p)] = dc.get((t, p), 0) + 1
                    elif kstar is not None:
                        if kstar >= j:                        # comp already true at k*: flanking (S:503-505)
                            if ok:
                                spanx[t] += 1
                            if combine:
                                flank[t] += 1
                        # else: comp was still false at k*: mutually exclusive stands (S:507-512)
                    elif covers and ok:                       # branch 4 beta1-type (S:544-552)
                        covx[t] += 1
                        if want_dc:
                            for p in set(g.ppos[t]) & read_sites:
                                dc[(t, p)] = dc.get((t, p), 0) + 1
    beta1, beta2s = [0] * S, [0] * S
    run = [[0, 0], [0, 0]]
    for t in range(S):
        for k in range(2):
            run[0][k] += cov[k][t]
            run[1][k] += span[k][t]
        if not stranded:
            c, s = run[0][0], run[1][0]
        elif g.cls[t] == 1:
            c, s = run[0][0], run[1][0]
        elif g.cls[t] == 2:
            c, s = run[0][1], run[1][1]
        else:
            c, s = 0, 0
        beta1[t] = c - covx[t]
        beta2s[t] = covx[t] + s - spanx[t] + flank[t]
    return beta1, beta2s, dc
