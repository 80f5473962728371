#!/usr/bin/env python3
"""SASS mnemonic counts per kernel of the built library (cuobjdump -sass) + ptxas -v resource lines ->
profiles/rNN_sass_evidence.txt.   python tools/sass_evidence.py > profiles/r02_sass_evidence.txt"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "caviar_b200", "csrc", "libcaviar_b200.so")
COLS = [("UTCHMMA", r"\bUTC[A-Z]*MMA"), ("UTMALDG", r"\bUTMALDG"), ("LDTM", r"\bLDTM"), ("UTCBAR", r"\bUTCBAR"),
        ("UBLKCP", r"\bUBLKCP"), ("UBLKPF", r"\bUBLKPF"), ("SYNCS", r"\bSYNCS"), ("BAR", r"\bBAR\."),
        ("F*2 packed", r"\b(FFMA2|FADD2|FMUL2)\b"), ("fp64", r"\b(DFMA|DADD|DMUL)\b"), ("I2F", r"\bI2F"), ("F2F", r"\bF2F"),
        ("LDS", r"\bLDS"), ("STG", r"\bSTG")]
sass = subprocess.run(["cuobjdump", "-sass", LIB], stdout=subprocess.PIPE, text=True).stdout
print(f"# SASS evidence for caviar_b200/csrc/libcaviar_b200.so (cuobjdump -sass); arch flags: -gencode arch=compute_100a,code=sm_100a -lineinfo -O3")
print("# per kernel: total instructions | " + " | ".join(c for c, _ in COLS))
cur, body = None, []


def flush():
    if cur is None:
        return
    ins = [l for l in body if re.search(r"/\*[0-9a-f]{4}\*/", l)]
    print(f"{cur:70s} {len(ins):5d} | " + " | ".join(f"{sum(1 for l in ins if re.search(p, l)):4d}" for _, p in COLS))


for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        flush()
        cur, body = m.group(1), []
    else:
        body.append(line)
flush()
print("\n# ptxas -v (registers / spills / shared memory)")
sys.stdout.flush()
res = subprocess.run([sys.executable, os.path.join(ROOT, "caviar_b200", "build.py"), "--force", "-v"], stdout=subprocess.PIPE,
                     stderr=subprocess.STDOUT, text=True).stdout
for l in res.splitlines():
    if "Compiling entry" in l or "registers" in l or "spill" in l:
        print(l.strip())
