#!/usr/bin/env python3
"""Join an ncu `--page source --csv` dump (per-SASS-instruction counters) with nvdisasm line info, so that
executed instructions and stall samples can be read per CUDA source line without the ncu GUI.

    ncu -i prof.ncu-rep --page source --csv > src.csv
    cuobjdump -xelf all libcaviar_b200.so && nvdisasm -g -c cv_api.sm_100a.cubin > dis.txt
    python tools/ncu_lines.py src.csv dis.txt 'features_kernelILi2ELb1' [source-file]
"""
import csv
import re
import sys
from collections import defaultdict


def main():
    src_csv, dis_txt, func = sys.argv[1:4]
    srcfile = sys.argv[4] if len(sys.argv) > 4 else None
    # 1. nvdisasm: ordered list of (line, inline-context) for every instruction of the function
    lines = []
    cur = None
    active = False
    for ln in open(dis_txt):
        if ln.startswith(".text."):
            active = func in ln
            continue
        if not active:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", ln):
            lines.append(cur)
    # 2. ncu rows in address order
    rows = list(csv.reader(open(src_csv)))
    hdr = rows[1]
    ci, cs, ca = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
    body = [r for r in rows[2:] if len(r) > ci]
    if len(body) != len(lines):
        print(f"warning: {len(body)} ncu rows vs {len(lines)} disassembled instructions", file=sys.stderr)
    per_line = defaultdict(lambda: [0, 0, defaultdict(int)])
    tot_i = tot_s = 0
    for r, loc in zip(body, lines):
        n, s = int(r[ci]), int(r[cs])
        tot_i += n; tot_s += s
        e = per_line[loc]
        e[0] += n; e[1] += s
        e[2][r[ca].split()[0] if r[ca].split() else "?"] += n
    text = {}
    if srcfile:
        for i, l in enumerate(open(srcfile), 1):
            text[i] = l.rstrip()
    print(f"total warp-instructions {tot_i}, samples {tot_s}")
    for loc, (n, s, ops) in sorted(per_line.items(), key=lambda kv: -kv[1][0])[:60]:
        top = ",".join(f"{k}:{v * 100 // max(n, 1)}" for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:4])
        same_file = bool(loc) and bool(srcfile) and loc[0] == srcfile.split("/")[-1]
        t = text.get(loc[1], "")[:90] if same_file else ("(inlined CUDA header intrinsic)" if loc else "")
        print(f"{100 * n / tot_i:5.1f}% inst {100 * s / max(tot_s, 1):5.1f}% stall  {loc}  [{top}]  {t.strip()}")


if __name__ == "__main__":
    main()
