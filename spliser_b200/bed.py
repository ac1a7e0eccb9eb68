"""BED12 junction file -> junction table (the text half of findAlphaCounts, SpliSER_v0_1_8.py:255-288).

Stays on the Python side of the ABI because it is text handling, not counting: the 12-column test
(S:259), the chromosome index in first-appearance order (S:265-268, appended to whatever the
annotation already registered, S:90-92), the -c filter (S:269) and the -g window filter (S:279-288).
"""
from __future__ import annotations

import numpy as np

from .api import Junctions


def parse_bed12(lines, chrom_index=None, qchrom="All", qgene_bounds=None, max_intron=0):
    """Returns (chrom_index, Junctions, strand_strings).

    chrom_index: list of names already registered by the annotation (mutated copy is returned).
    qgene_bounds: (leftPos, rightPos) of the query gene when -g is used, else None.
    strand_strings keeps the full column-6 text per kept row: the Gene column and the TSV print it
    verbatim, while the ABI only carries its first byte."""
    chroms = list(chrom_index) if chrom_index else []
    index = {c: i for i, c in enumerate(chroms)}
    jc, jl, jr, js, jst, sstr = [], [], [], [], [], []
    for line in lines:
        v = str(line).split("\t")
        if len(v) != 12:                       # header / malformed line (S:259)
            continue
        chrom = v[0]
        ci = index.get(chrom)
        if ci is None:
            ci = index[chrom] = len(chroms)
            chroms.append(chrom)
        if not (qchrom == chrom or qchrom == "All"):
            continue
        flank = v[10].split(",")
        left = int(v[1]) + int(flank[0])       # S:275
        right = int(v[2]) - int(flank[1])      # S:276
        score = int(v[4])                      # S:277
        if qgene_bounds is not None:           # S:279-288
            gl, gr = qgene_bounds
            lin = (left + max_intron >= gl) and (left <= gr)
            rin = (right - max_intron <= gr) and (right >= gl)
            if not (lin or rin):
                continue
        jc.append(ci); jl.append(left); jr.append(right); js.append(score)
        jst.append(ord(v[5][0]) if v[5] else 0)
        sstr.append(v[5])
    j = Junctions(np.array(jc, np.int32), np.array(jl, np.int32), np.array(jr, np.int32),
                  np.array(js, np.int64), np.array(jst, np.uint8))
    return chroms, j, sstr
