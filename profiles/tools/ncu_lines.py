"""Per-source-line view of an ncu capture, for kernels compiled with -lineinfo.

    ncu -i X.ncu-rep --page source --csv > sass.csv
    cuobjdump -xelf all libspliser_b200.so ; nvdisasm -g -c <file>.cubin > file.sass
    python profiles/tools/ncu_lines.py sass.csv file.sass <kernel substring> [top N]

ncu's CSV has one row per SASS instruction (address order) with executed-instruction counts and stall samples; nvdisasm
-g prints the same instructions with '//## File "...", line N' markers.  The two are aligned by instruction offset.
"""
import csv
import re
import sys
from collections import defaultdict


def main():
    sass_csv, disasm, kern = sys.argv[1], sys.argv[2], sys.argv[3]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    rows = list(csv.reader(open(sass_csv)))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hdr_i]
    col = {h: i for i, h in enumerate(hdr)}
    inst = rows[hdr_i + 1:]
    base = int(inst[0][0], 16)
    by_off = {}
    for r in inst:
        by_off[int(r[0], 16) - base] = r
    # walk the disassembly of the kernel
    line_of = {}
    cur = ("?", 0)
    inside = False
    for ln in open(disasm):
        if ln.startswith(".text."):
            inside = kern in ln
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/", ln)
        if m:
            line_of[int(m.group(1), 16)] = cur
    agg = defaultdict(lambda: [0, 0, 0, 0])          # inst executed, samples, long_sb, wait/short
    tot = [0, 0]
    for off, r in by_off.items():
        key = line_of.get(off, ("?", 0))
        ie = int(r[col["Instructions Executed"]] or 0)
        sm = int(r[col["# Samples"]] or 0)
        agg[key][0] += ie; agg[key][1] += sm
        agg[key][2] += int(r[col["stall_long_sb"]] or 0)
        agg[key][3] += int(r[col["stall_short_sb"]] or 0) + int(r[col["stall_wait"]] or 0)
        tot[0] += ie; tot[1] += sm
    print("total warp instructions %d, samples %d" % (tot[0], tot[1]))
    print("%-28s %14s %6s %9s %6s %9s %9s" % ("file:line", "inst", "%", "samples", "%", "long_sb", "short+wait"))
    for key, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print("%-28s %14d %6.2f %9d %6.2f %9d %9d" % ("%s:%d" % key, v[0], 100.0 * v[0] / max(1, tot[0]), v[1], 100.0 * v[1] / max(1, tot[1]), v[2], v[3]))


if __name__ == "__main__":
    main()
