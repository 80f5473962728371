#!/usr/bin/env python3
"""CPU box: turn what tools/gpu_profile.sh brought back in gpurun_out/ into the committed summaries under profiles/:

    python tools/make_profiles.py [round-tag, default r01]

  profiles/<tag>_launches.csv                    compact launch list of the bench command (ncu, per-launch durations)
  profiles/<tag>_features_kernel_ncu_raw.txt     the raw --set full metrics we track for the fused kernel
  profiles/<tag>_features_kernel_hot_lines.txt   executed instructions / stall samples per CUDA source line
  profiles/<tag>_sass_evidence.txt               SASS mnemonic counts + ptxas resource usage of the built library
  profiles/traffic.json                          DRAM bytes per frame of the capture (bench.py's roofline.traffic)
"""
import csv
import json
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TAG = sys.argv[1] if len(sys.argv) > 1 else "r01"
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")
REP = os.path.join(OUT, "prof_features.ncu-rep")
FRAMES = 16384            # bench.py --frames of the capture (tools/gpu_profile.sh)


def sh(cmd, **kw):
    return subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, **kw).stdout


# 1. launch list
rows = [r for r in csv.reader(open(os.path.join(OUT, "launches.csv"))) if len(r) > 10 and r[0].isdigit()]
with open(os.path.join(PROF, f"{TAG}_launches.csv"), "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline\n")
    f.write("# (first 400 launches: set-up -- synthetic frames by torch RNG, box records, K0 staging -- then 3 warm-up + 2 timed steps;\n")
    f.write("#  kernel names cut at 90 characters; durations are cold-cache and serialised: shares, not absolutes)\n")
    f.write("id,kernel,block,grid,duration_ns\n")
    for r in rows:
        f.write(f'{r[0]},"{r[4][:90]}","{r[7]}","{r[8]}",{r[-1]}\n')
ours = {}
for r in rows:
    if "cv::" in r[4]:
        k = r[4].split("(")[0]
        ours.setdefault(k, [0, 0.0])
        ours[k][0] += 1
        ours[k][1] += float(r[-1]) / 1e6
print("our kernels in the launch list:", {k: (n, round(ms, 3)) for k, (n, ms) in ours.items()})

# 2. raw metrics + traffic
raw = sh([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), REP])
open(os.path.join(PROF, f"{TAG}_features_kernel_ncu_raw.txt"), "w").write(raw)
val = {}
for line in raw.splitlines():
    m = re.match(r"(\S+)\s+(\S+)\s+(\S+)$", line.strip())
    if m:
        val[m.group(1)] = (m.group(2), m.group(3))
scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
rd = float(val["dram__bytes_read.sum"][1]) * scale[val["dram__bytes_read.sum"][0]]
wr = float(val["dram__bytes_write.sum"][1]) * scale[val["dram__bytes_write.sum"][0]]
dur = val["gpu__time_duration.sum"]
json.dump({"cfg3": {"dram_bytes_per_frame": (rd + wr) / FRAMES, "dram_bytes_read": rd, "dram_bytes_write": wr,
                    "frames_in_capture": FRAMES, "kernel": "cv::features_kernel<2,true>",
                    "kernel_duration_under_ncu": f"{dur[1]} {dur[0]}",
                    "source": f"profiles/{TAG}_features_kernel_ncu_raw.txt (tools/gpu_profile.sh: ncu --set full "
                              f"--clock-control none, bench.py --frames {FRAMES})"}},
          open(os.path.join(PROF, "traffic.json"), "w"), indent=1)
print("traffic:", (rd + wr) / FRAMES, "B/frame; duration", dur)

# 3. per-line hot spots
with tempfile.TemporaryDirectory() as tmp:
    lib = os.path.join(ROOT, "caviar_b200", "csrc", "libcaviar_b200.so")
    subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, stdout=subprocess.DEVNULL)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    open(os.path.join(tmp, "dis.txt"), "w").write(sh(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)]))
    open(os.path.join(tmp, "src.csv"), "w").write(sh(["ncu", "-i", REP, "--page", "source", "--csv"]))
    hot = sh([sys.executable, os.path.join(ROOT, "tools", "ncu_lines.py"), os.path.join(tmp, "src.csv"),
              os.path.join(tmp, "dis.txt"), "features_kernelILi2ELb1", os.path.join(ROOT, "caviar_b200", "csrc", "cv_device.cuh")])
    open(os.path.join(PROF, f"{TAG}_features_kernel_hot_lines.txt"), "w").write(hot)

# 4. SASS evidence (rebuilds the library with -Xptxas -v)
open(os.path.join(PROF, f"{TAG}_sass_evidence.txt"), "w").write(sh([sys.executable, os.path.join(ROOT, "tools", "sass_evidence.py")]))
print("done")
