#!/usr/bin/env python3
"""Headline benchmark of the CAVIAR pair-feature hot path on B200 (BASELINE.json metric, config 3).

    python bench.py --gpus 1 --steps 10 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \\
           bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference --steps 3 --warmup 1      # the reference's CPU path on the host cores

One "step" = one pass of the fused hot path (distances, angles, flags, scores, occupancy / moment
aggregates) over the rank's whole shard of frames, resident in HBM in the staged layout, followed -- when
N > 1 -- by the single NCCL all-reduce of the per-feature aggregates.  Frames are sharded across ranks
(weak scaling: every rank owns --frames frames).  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "pair-distance evals/sec (frames x (pairs + triplets) / s)"
UNIT = "evals/s"
CUTS = dict(pair_cutoff=8.0, hbond_dist_cutoff=3.0, hbond_angle_cutoff=135.0)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=3, help="BASELINE.json config number (3 = single-GPU roofline case)")
    ap.add_argument("--frames", type=int, default=0, help="frames per rank (0 = the config's frame count)")
    ap.add_argument("--e2e-frames", type=int, default=4096, help="frames per end-to-end step (host buffers)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-frames", type=int, default=1000, help="frames of the CPU baseline sample")
    ap.add_argument("--direct-frames", type=int, default=4096, help="frames of the raw-frame (direct gather) side measurement; 0 = skip")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi sampled every 20 ms while the warm-up and timed steps run."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        if os.environ.get("CAVIAR_SMI_MS") == "0":
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", os.environ.get("CAVIAR_SMI_MS", "20"), "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        # "under load" = the samples at the highest power draw half (idle samples before the first kernel are dropped)
        pw = []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            try:
                pw.append(float(f[3]))
            except (ValueError, IndexError):
                pw.append(0.0)
        loaded = [c for c, w in zip(sm, pw) if pw and w >= 0.5 * max(pw)] or sm
        return {"sm_mhz": float(np.median(loaded)) if loaded else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "samples_under_load": len(loaded), "power_w_max": max(pw) if pw else None,
                "reasons": sorted(reasons)}


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(cfg_id, n_frames):
    """dram__bytes_read.sum + dram__bytes_write.sum of the fused kernel, per launch, from the committed ncu
    --set full capture (profiles/traffic.json holds bytes per frame of that capture; scaled to this launch)."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            rec = json.load(fh)
        return rec[f"cfg{cfg_id}"]["dram_bytes_per_frame"] * n_frames
    except Exception:
        return None


# ---------------------------------------------------------------------------------------------------------
# CPU arms (oracle port of the reference; the only place bench.py touches oracle/)
# ---------------------------------------------------------------------------------------------------------
def cpu_reference_pass(x, box, pairs, triplets, n_jobs):
    """One pass of the reference's CPU path over the sample: compute_distances_parallel (CAVIAR-1.0.py:110-122,
    no PBC -- the reference has none), the per-pair time mean (:342) and the `< cut` threshold (:186), in numpy
    with joblib threads exactly like the reference.  Triplet angles / imaging have no reference code (the
    Extractor hands them to cpptraj, a compiled program): they run in the C oracle on the same host threads.
    Returns seconds per part."""
    from oracle import c_oracle as co
    from oracle import caviar_oracle as orc
    t0 = time.perf_counter()
    d = orc.pair_distances(x, [tuple(p) for p in pairs.tolist()], n_jobs=n_jobs)
    t1 = time.perf_counter()
    mean = orc.pair_time_mean(d)
    flags = orc.contact_flags(d, CUTS["pair_cutoff"])
    occ = orc.occupancy(flags)
    t2 = time.perf_counter()
    ang = None
    if len(triplets):
        ang, dac = co.triplet_angles(x, triplets, box, n_threads=n_jobs, with_end_distances=True)
        hb = orc.hbond_flags(dac, ang, CUTS["hbond_dist_cutoff"], CUTS["hbond_angle_cutoff"])
        occ_t = orc.occupancy(hb)
    t3 = time.perf_counter()
    return {"pairs_s": t1 - t0, "aggregate_s": t2 - t1, "triplets_s": t3 - t2, "total_s": t3 - t0,
            "dist": d, "angle": ang, "mean": mean, "occ": occ}


def cpu_fused_c_pass(x, box, pairs, triplets, n_jobs):
    """The WHOLE path (minimum-image distances, angles, flags, scores, occupancy, moments) as one multi-threaded C
    pass per frame (oracle/caviar_oracle.c): the strongest CPU comparator we have.  Returns (seconds, result)."""
    from oracle import c_oracle as co
    t0 = time.perf_counter()
    r = co.fused_pass(x, pairs, triplets if len(triplets) else None, box, CUTS["pair_cutoff"], CUTS["hbond_dist_cutoff"],
                      CUTS["hbond_angle_cutoff"], n_threads=n_jobs)
    return time.perf_counter() - t0, r


def run_reference_arm(args):
    """--impl reference: the reference's CPU implementation of the path, timed on the host cores (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from caviar_b200 import synthetic as syn
    top = syn.make_topology(args.config)
    cfg = top.cfg
    pairs, triplets = syn.compact_indices(top)
    T = max(50, min(args.cpu_frames, 500))
    x = syn.parity_frames(top, T, only_referenced=True)
    box = syn.boxes_for(top, T)
    n_jobs = max((os.cpu_count() or 2) - 1, 1)
    times = []
    for i in range(args.warmup + args.steps):
        r = cpu_reference_pass(x, box, pairs, triplets, n_jobs)
        if i >= args.warmup:
            times.append(r["total_s"])
    ms = 1e3 * float(np.mean(times))
    evals = T * (len(pairs) + len(triplets))
    value = evals / (ms * 1e-3)
    sample = (f"{T} frames of {cfg.name} (compact (T,U,3) input as the reference feeds it, CAVIAR-1.0.py:335); "
              f"pairs via the reference algorithm without PBC (it has none: numpy + joblib threads), triplets with "
              f"27-image PBC via the C oracle on the same threads (cpptraj stand-in)")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(cfg, T, len(pairs), len(triplets), len(top.referenced)),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": n_jobs, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def workload_config(cfg, frames_per_rank, P, A, U, **extra):
    d = {"workload": f"config {cfg.name[3:]}: {cfg.description}", "n_atoms": cfg.n_atoms, "frames_per_gpu": frames_per_rank,
         "pairs": P, "triplets": A, "distinct_atoms_referenced": U, "box": cfg.box or "none",
         "cutoffs": CUTS, "outputs": "distances, angles, packed flags, frame scores, occupancy + fp64 sum/sumsq"}
    d.update(extra)
    return d


# ---------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------
def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch
    import torch.distributed as dist

    import caviar_b200 as cb
    from caviar_b200 import synthetic as syn

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: caviar_b200 has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa_node = None
    if world > 1:
        # one process per GPU: keep this rank's threads and page-locked buffers on the GPU's own NUMA node
        from caviar_b200.sharding import bind_to_gpu_numa_node
        numa_node = bind_to_gpu_numa_node(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL writes its version banner to stdout (C level) when the communicator comes up; rank 0's stdout must
        # carry exactly ONE JSON line, so file descriptor 1 points at stderr until the first collective is through.
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            warm = torch.zeros(1, device=dev)
            dist.all_reduce(warm)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)

    top = syn.make_topology(args.config)
    cfg = top.cfg
    T = args.frames or cfg.n_frames
    pairs, triplets = syn.compact_indices(top)       # compact numbering of the U referenced atoms
    U, P, A = len(top.referenced), len(pairs), len(triplets)

    t_setup = time.perf_counter()
    plan = cb.FeaturePlan(U, pairs, triplets, box=cfg.box, **CUTS)
    info = plan.info
    S = int(info["staged_floats_per_frame"]) if info["mode_name"] == "staged" else 3 * U
    # ---- resident inputs: the rank's whole shard in the layout the fused kernel streams ----
    coords = torch.empty((T, S) if info["mode_name"] == "staged" else (T, U, 3), dtype=torch.float32, device=dev)
    box_np = syn.boxes_for(top, T)
    rec = plan.prepare_box(torch.from_numpy(box_np).to(dev)) if box_np is not None else None
    chunk = max(1, min(T, (1 << 30) // (U * 12)))
    for t0 in range(0, T, chunk):
        t1 = min(T, t0 + chunk)
        raw = syn.device_frames(top, t1 - t0, device=dev, seed=1000 * rank + t0, only_referenced=True)
        if info["mode_name"] == "staged":
            plan.stage(raw, None if rec is None else rec[t0:t1], out=coords[t0:t1])
        else:
            coords[t0:t1].copy_(raw)
    del raw
    out = plan.alloc_outputs(T, device=dev)
    agg_f = torch.zeros((2, plan.F), dtype=torch.float64, device=dev)
    agg = cb.Aggregates(torch.zeros(plan.F, dtype=torch.int64, device=dev), agg_f[0], agg_f[1])
    ws = plan.new_workspace(dev)
    torch.cuda.synchronize()
    t_setup = time.perf_counter() - t_setup

    def step():
        agg.occupancy.zero_(); agg_f.zero_()
        plan.run(coords, rec, out=out, aggregates=agg, workspace=ws)
        if world > 1:                                   # the one exchange step of the path
            dist.all_reduce(agg.occupancy); dist.all_reduce(agg_f)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)           # samples through warm-up and the timed region (both are load)
    sampler.start()
    time.sleep(0.3)                              # let nvidia-smi come up before the kernels start
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    launches_per_step = plan.last_launches
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True),
           torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for k in range(args.steps):
        ev[k][0].record()
        agg.occupancy.zero_(); agg_f.zero_()
        plan.run(coords, rec, out=out, aggregates=agg, workspace=ws)
        ev[k][1].record()
        if world > 1:
            dist.all_reduce(agg.occupancy); dist.all_reduce(agg_f)
        ev[k][2].record()
    barrier()
    clocks = sampler.stop()
    total_ms = ev[0][0].elapsed_time(ev[-1][2])
    kern_ms = float(np.mean([a.elapsed_time(b) for a, b, _ in ev]))
    reduce_ms = float(np.mean([b.elapsed_time(c) for _, b, c in ev]))      # the one exchange step (0 at N = 1)
    tm = torch.tensor([total_ms, kern_ms, reduce_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    total_ms, kern_ms, reduce_ms = float(tm[0]), float(tm[1]), float(tm[2])
    ms_per_step = total_ms / args.steps
    evals_per_step = world * T * (P + A)
    value = evals_per_step / (ms_per_step * 1e-3)

    # ---- size-independent property at the full size: checksum of checksums (sum of frame scores == sum of
    #      occupancies == number of set flag bits), on the outputs of the last timed step ----
    score_sum = out.score.sum(dtype=torch.int64)
    bits = out.flags.view(torch.uint8)
    lut = torch.tensor([bin(i).count("1") for i in range(256)], dtype=torch.int64, device=dev)
    bit_sum = torch.zeros((), dtype=torch.int64, device=dev)
    for c0 in range(0, bits.shape[0], 8192):
        bit_sum += lut[bits[c0:c0 + 8192].long()].sum()
    sums = torch.stack([score_sum, bit_sum])
    if world > 1:
        dist.all_reduce(sums)                       # scores / flags are rank-local, the occupancy is already global
    occ_sum = int(agg.occupancy.sum())
    full_check = {"score_sum": int(sums[0]), "flag_bits": int(sums[1]), "occupancy_sum": occ_sum,
                  "equal": bool(int(sums[0]) == int(sums[1]) == occ_sum)}

    # ---- roofline of the fused kernel (algorithmic bytes per launch / its average duration) ----
    peak, peak_src = measured_peak()
    alg_bytes = int(info["algorithmic_bytes_per_frame"]) * T
    achieved = alg_bytes / (kern_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": ncu_traffic(args.config, T), "kernel": "cv::features_kernel<BOX=2,PIPE,FRAC> (+finalize)",
                "kernel_ms": kern_ms, "algorithmic_bytes_per_launch": alg_bytes, "peak_source": peak_src,
                "frac_of_nominal_8000": achieved / 8000.0,      # north_star's target is quoted on the ~8 TB/s nominal peak
                "streamed_bytes_per_launch": T * (S * 4 + 4 * info["boxrec_floats_per_frame"] * info["n_tiles"]
                                                  + 4 * P + 4 * A + 4 * plan.W + 4)}

    # ---- raw (T, N, 3) frames resident in HBM: compaction (K0) + fused kernel, and the direct-gather kernel ----
    direct = None
    if args.direct_frames > 0 and world == 1:
        del coords
        Td = min(args.direct_frames, T)
        raw = syn.device_frames(top, Td, device=dev, seed=77, only_referenced=False)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5

        def timed(fn):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            e0.record()
            for _ in range(reps):
                fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / reps

        with cb.FeaturePlan(cfg.n_atoms, top.pairs, top.triplets, box=cfg.box, **CUTS) as fplan:
            fws = fplan.new_workspace(dev)
            staged_buf = torch.empty((Td, int(fplan.info["staged_floats_per_frame"])), dtype=torch.float32, device=dev)
            ms_stage = timed(lambda: fplan.stage(raw, None if rec is None else rec[:Td], out=staged_buf))
            ms_both = timed(lambda: fplan.run(fplan.stage(raw, None if rec is None else rec[:Td], out=staged_buf), rec[:Td], out=out, aggregates=agg,
                                              workspace=fws, n_frames=Td))
            del staged_buf
        with cb.FeaturePlan(cfg.n_atoms, top.pairs, top.triplets, box=cfg.box, force="direct", **CUTS) as dplan:
            dws = dplan.new_workspace(dev)
            ms_direct = timed(lambda: dplan.run(raw, rec[:Td], out=out, aggregates=agg, workspace=dws, n_frames=Td))
        alg = int(info["algorithmic_bytes_per_frame"]) * Td
        direct = {"input": f"raw ({Td}, {cfg.n_atoms}, 3) frames resident in HBM (12*N = {12 * cfg.n_atoms} B per frame)",
                  "compaction_plus_fused": {"value": Td * (P + A) / (ms_both * 1e-3), "unit": UNIT, "ms": ms_both,
                                            "compaction_ms": ms_stage, "roofline_frac": alg / (ms_both * 1e-3) / 1e9 / peak},
                  "direct_gather": {"value": Td * (P + A) / (ms_direct * 1e-3), "unit": UNIT, "ms": ms_direct,
                                    "roofline_frac": alg / (ms_direct * 1e-3) / 1e9 / peak}}
        del raw

    # ---- end to end through the reference-facing host call (pinned host buffers, H2D + D2H inside) ----
    Te = min(args.e2e_frames, T)
    xh = torch.empty((Te, U, 3), dtype=torch.float32).pin_memory()
    xh.copy_(syn.device_frames(top, Te, device=dev, seed=4242 + rank, only_referenced=True).cpu())
    host_out = {"dist": torch.empty((Te, P), dtype=torch.float32).pin_memory() if P else None,
                "angle": torch.empty((Te, A), dtype=torch.float32).pin_memory() if A else None,
                "flags": torch.empty((Te, plan.W), dtype=torch.int32).pin_memory(),
                "score": torch.empty((Te,), dtype=torch.int32).pin_memory()}
    host_out = {k: v for k, v in host_out.items() if v is not None}
    box_e = box_np[:Te] if box_np is not None else None
    e2e_launches = 0

    def e2e_step():
        res = plan.features_host(xh.numpy(), box_e, out=dict(host_out))
        if world > 1:
            o = torch.from_numpy(res["occupancy"].astype(np.int64)).to(dev)
            s = torch.from_numpy(np.stack([res["sum"], res["sumsq"]])).to(dev)
            dist.all_reduce(o); dist.all_reduce(s)
            torch.cuda.synchronize()
        return res

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        e2e_step()
        e2e_launches += plan.last_launches
    barrier()
    e2e_s = (time.perf_counter() - t0) / args.e2e_steps
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te[0])
    box_b = {None: 0, "ortho": 12, "triclinic": 36}[cfg.box]
    e2e = {"value": world * Te * (P + A) / e2e_s, "unit": UNIT,
           "h2d_bytes_per_step": Te * (U * 12 + box_b),
           "d2h_bytes_per_step": Te * (4 * P + 4 * A + 4 * plan.W + 4) + 24 * plan.F,
           "frames_per_step": Te, "ms_per_step": 1e3 * e2e_s,
           "call": "FeaturePlan.features_host -> caviar_features_host (pinned host buffers)"}

    # ---- CPU baseline beside it (rank 0, N = 1): bounded sample of the same workload ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        Tc = min(args.cpu_frames, Te)
        n_jobs = max((os.cpu_count() or 2) - 1, 1)
        xs = xh.numpy()[:Tc]
        r = cpu_reference_pass(xs, box_e[:Tc] if box_e is not None else None, pairs, triplets, n_jobs)
        # the CPU pass doubles as a spot check of the GPU result on the same frames
        res = plan.features_host(xs, box_e[:Tc] if box_e is not None else None, want=("dist", "angle"))
        check = {}
        if A:
            check["max_angle_err_deg"] = float(np.abs(res["angle"] - r["angle"]).max())
        if P and cfg.box is None:
            check["dist_bit_exact"] = bool(res["dist"].tobytes() == r["dist"].tobytes())
        elif P:
            same = np.abs(res["dist"] - r["dist"]) < 1e-4      # pairs whose minimum image is the unshifted one
            check["pairs_unshifted_frac_within_1e-4A"] = float(same.mean())
        c_s, rc = cpu_fused_c_pass(xs, box_e[:Tc] if box_e is not None else None, pairs, triplets, n_jobs)
        if P and cfg.box is not None:
            check["max_dist_err_A_vs_c_oracle"] = float(np.abs(res["dist"].astype(np.float64) - rc["dist"]).max())
        cpu = {"value": Tc * (P + A) / r["total_s"], "unit": UNIT, "cores": n_jobs, "kind": "port",
               "sample": f"{Tc} frames of {cfg.name}: reference compute_distances_parallel algorithm (numpy, no PBC, joblib "
                         f"threads) {r['pairs_s']:.2f}s + mean/threshold {r['aggregate_s']:.2f}s + triplet angles with "
                         f"27-image PBC in the C oracle (cpptraj stand-in) {r['triplets_s']:.2f}s",
               "pairs_only_value": Tc * P / max(r["pairs_s"] + r["aggregate_s"], 1e-9),
               "c_fused_value": Tc * (P + A) / max(c_s, 1e-9),
               "c_fused_note": f"whole path incl. PBC as one multi-threaded C pass (oracle/caviar_oracle.c), {rc['threads']} threads, {c_s:.2f}s",
               "spot_check": check}

    # ---- the reference's own shape through its own operator (B1), pageable numpy in / numpy out ----
    native = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import caviar_oracle as orc          # cpu_baseline leg: the checker / CPU arm, never the product
        Tn, nres = 7000, 300                              # T=7000, 300 residues, minsep 5
        rng = np.random.default_rng(0)
        Xn = (np.cumsum(rng.normal(0, 0.38, size=(nres, 3)), axis=0)[None]
              + rng.normal(0, 0.05, size=(Tn, nres, 3))).astype(np.float32)
        pl = [(i, j) for i in range(nres) for j in range(i + 5, nres)]
        cb.compute_distances_parallel(Xn[:64], pl)        # warm-up
        ts = []
        for _ in range(3):
            t0 = time.perf_counter(); Dn = cb.compute_distances_parallel(Xn, pl); ts.append(time.perf_counter() - t0)
        n_jobs = max((os.cpu_count() or 2) - 1, 1)
        t0 = time.perf_counter(); Dr = orc.pair_distances(Xn, pl, n_jobs=n_jobs); cpu_s = time.perf_counter() - t0
        native = {"call": "compute_distances_parallel(X (7000,300,3), all pairs minsep 5 -> P=43660)  [CAVIAR-1.0.py:110-122]",
                  "ours_s": min(ts), "ours_evals_per_s": Tn * len(pl) / min(ts), "cpu_s": cpu_s, "cpu_threads": n_jobs,
                  "cpu_evals_per_s": Tn * len(pl) / cpu_s, "bit_exact": bool(Dn.tobytes() == Dr.tobytes())}
        del Dn, Dr

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": workload_config(cfg, T, P, A, U, plan=info, l2="inputs (%.1f GB per rank) >> 126 MB L2, no flush needed"
                                          % (coords_bytes(T, S) / 1e9), setup_s=round(t_setup, 1),
                                          parallelism=f"frame-sharded x{world}, one NCCL all-reduce of aggregates per step",
                                          rank0_numa_node=numa_node),
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "clocks": clocks,
                "gpu_launches": launches_per_step * args.steps,
                "exchange": {"what": "NCCL sum all-reduce of occupancy i64[F] + {sum, sumsq} f64[2F] (24*F bytes)",
                             "bytes": 24 * plan.F, "ms_per_step": reduce_ms if world > 1 else 0.0},
                "raw_frame_input": direct,
                "reference_native_shape": native, "full_size_check": full_check}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def coords_bytes(T, S):
    return 4.0 * T * S


if __name__ == "__main__":
    main()
