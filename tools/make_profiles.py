#!/usr/bin/env python3
"""CPU box: turn what tools/gpu_profile.sh brought back in gpurun_out/ into the committed summaries under profiles/:

    python tools/make_profiles.py [round-tag, default r02]

  profiles/<tag>_launches.csv                 compact launch list of the bench command (ncu, per-launch durations)
  profiles/<tag>_<kernel>_ncu_raw.txt         the raw --set full metrics we track, one file per captured kernel
  profiles/<tag>_features_kernel_hot_lines.txt   executed instructions / stall samples per CUDA source line (cfg3)
  profiles/<tag>_sass_evidence.txt            SASS mnemonic counts + ptxas resource usage of the built library
  profiles/traffic.json                       DRAM bytes per frame of every capture (bench.py's roofline.traffic) and the
                                              warp instructions per frame of the config-4 captures (issue roofline)
"""
import csv
import json
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TAG = sys.argv[1] if len(sys.argv) > 1 else "r02"
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")

# capture file -> (traffic.json key, frames in the capture, file stem, command[, launches in the capture])
CAPTURES = [
    ("prof_features", "cfg3", 16384, "features_kernel_cfg3", "tools/gpu_probe.py 3 16384: cv::features_kernel<2,true,false> on raw plan-order frames"),
    ("prof_stage", "cfg3_k0", 16384, "stage_frames_kernel_cfg3", "tools/gpu_probe.py 3 16384: cv::stage_frames_kernel (caller order -> plan order)"),
    ("prof_cfg4_full", "cfg4", 2048, "allpairs_kernel_cfg4_full", "tools/gpu_cfg4.py 2048: cv::allpairs_kernel<false,true> (fast tiles) + <false,false> (general tiles), dist + flags + scores + occupancy + moments", 2),
    ("prof_cfg4_lean", "cfg4_lean", 2048, "allpairs_kernel_cfg4_contact_map", "tools/gpu_cfg4.py 2048: cv::allpairs_kernel<true,true> + <true,false>, flags + scores + occupancy", 2),
    ("prof_cfg5", "cfg5", 4096, "features_kernel_cfg5", "tools/gpu_probe.py 5 4096: cv::features_kernel<1,true,false>"),
    ("prof_cov", "cov_n14000_k4096", 1, "cov_gemm_kernel", "tools/gpu_cov.py 14000 4096: cvcov::cov_gemm_kernel (tcgen05 kind::tf32)"),
]


def sh(cmd, **kw):
    return subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, **kw).stdout


# 1. launch list
launch_csv = os.path.join(OUT, "launches.csv")
if os.path.exists(launch_csv):
    rows = [r for r in csv.reader(open(launch_csv)) if len(r) > 10 and r[0].isdigit()]
    with open(os.path.join(PROF, f"{TAG}_launches.csv"), "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-all-configs\n")
        f.write("# (first 600 launches: set-up -- synthetic frames by torch RNG, box records, the slicing into plan order -- then 3 warm-up +\n")
        f.write("#  2 timed steps = cv::features_kernel + cv::finalize_kernel each, then the side measurements; kernel names cut at 90\n")
        f.write("#  characters; durations are cold-cache and serialised: shares, not absolutes)\n")
        f.write("id,kernel,block,grid,duration_ns\n")
        for r in rows:
            f.write(f'{r[0]},"{r[4][:90]}","{r[7]}","{r[8]}",{r[-1]}\n')
    ours = {}
    for r in rows:
        if "cv::" in r[4] or "cvcov::" in r[4]:
            k = r[4].split("(")[0]
            ours.setdefault(k, [0, 0.0])
            ours[k][0] += 1
            ours[k][1] += float(r[-1]) / 1e6
    print("our kernels in the launch list:", {k: (n, round(ms, 3)) for k, (n, ms) in ours.items()})

# 2. raw metrics + traffic
scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
traffic = {}
for cap in CAPTURES:
    stem, key, frames, name, what = cap[:5]
    n_rows = cap[5] if len(cap) > 5 else 1
    rep = os.path.join(OUT, stem + ".ncu-rep")
    if not os.path.exists(rep):
        print("missing", rep)
        continue
    text = f"# {what}\n# ncu --set full --import-source on --clock-control none, {n_rows} launch(es)\n"
    rd = wr = inst = dur_us = issue_w = dram_w = 0.0
    ok = True
    for row in range(n_rows):
        raw = sh([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep, str(row)])
        text += raw + ("\n" if row + 1 < n_rows else "")
        val = {}
        for line in raw.splitlines():
            m = re.match(r"(\S+)\s+(\S+)\s+(\S+)$", line.strip())
            if m:
                val[m.group(1)] = (m.group(2), m.group(3))
        try:
            rd += float(val["dram__bytes_read.sum"][1]) * scale[val["dram__bytes_read.sum"][0]]
            wr += float(val["dram__bytes_write.sum"][1]) * scale[val["dram__bytes_write.sum"][0]]
        except KeyError:
            print("no dram metrics in", rep)
            ok = False
            break
        unit, v = val["gpu__time_duration.sum"]
        us = float(v) * {"ms": 1e3, "us": 1.0, "ns": 1e-3, "s": 1e6}.get(unit, 1.0)
        dur_us += us
        issue_w += us * float(val.get("smsp__issue_active.avg.pct_of_peak_sustained_active", ("", "nan"))[1])
        dram_w += us * float(val.get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", ("", "nan"))[1])
        inst += float(val.get("smsp__inst_executed.sum", ("", "0"))[1])
    open(os.path.join(PROF, f"{TAG}_{name}_ncu_raw.txt"), "w").write(text)
    if not ok:
        continue
    entry = {"dram_bytes_per_frame": (rd + wr) / frames, "dram_bytes_read": rd, "dram_bytes_write": wr,
             "frames_in_capture": frames, "kernel_duration_under_ncu": f"{dur_us / 1e3:.6f} ms",
             "issue_active_pct": issue_w / dur_us, "dram_throughput_pct_of_peak": dram_w / dur_us,
             "source": f"profiles/{TAG}_{name}_ncu_raw.txt ({what})"}
    if inst:
        entry["warp_instructions_per_frame"] = inst / frames
    traffic[key] = entry
    print(key, {k: entry[k] for k in ("dram_bytes_per_frame", "kernel_duration_under_ncu", "issue_active_pct")})
if traffic:
    json.dump(traffic, open(os.path.join(PROF, "traffic.json"), "w"), indent=1)

# 3. per-line hot spots of the config-3 capture
rep = os.path.join(OUT, "prof_features.ncu-rep")
if os.path.exists(rep):
    with tempfile.TemporaryDirectory() as tmp:
        lib = os.path.join(ROOT, "caviar_b200", "csrc", "libcaviar_b200.so")
        subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, stdout=subprocess.DEVNULL)
        cubins = sorted(f for f in os.listdir(tmp) if f.endswith(".cubin"))
        for cubin in cubins:
            dis = sh(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)])
            if "features_kernelILi2ELb1ELb0" not in dis:
                continue
            open(os.path.join(tmp, "dis.txt"), "w").write(dis)
            open(os.path.join(tmp, "src.csv"), "w").write(sh(["ncu", "-i", rep, "--page", "source", "--csv"]))
            hot = sh([sys.executable, os.path.join(ROOT, "tools", "ncu_lines.py"), os.path.join(tmp, "src.csv"),
                      os.path.join(tmp, "dis.txt"), "features_kernelILi2ELb1ELb0", os.path.join(ROOT, "caviar_b200", "csrc", "cv_device.cuh")])
            open(os.path.join(PROF, f"{TAG}_features_kernel_hot_lines.txt"), "w").write(hot)

# 4. SASS evidence (rebuilds the library with -Xptxas -v)
open(os.path.join(PROF, f"{TAG}_sass_evidence.txt"), "w").write(sh([sys.executable, os.path.join(ROOT, "tools", "sass_evidence.py")]))
print("done")
