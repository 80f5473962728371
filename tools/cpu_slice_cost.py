#!/usr/bin/env python3
"""What handing frames over in PLAN order costs whoever slices the selected atoms out of the trajectory (CAVIAR-1.0.py:335):
the same host-side gather with the caller's sorted list and with plan.atom_order, config 3, 256 raw frames of 100 000 atoms.
Runs anywhere (host only)."""
import sys, time
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import caviar_b200 as cb
from caviar_b200 import synthetic as syn, _lib
top = syn.make_topology(3)
N = top.cfg.n_atoms; T = 256
rng = np.random.default_rng(0)
frames = rng.random((T, N, 3), dtype=np.float32)            # 307 MB of raw frames (the trajectory)
sel = np.asarray(top.referenced, dtype=np.int32)            # the caller's (sorted) selection
plan = cb.FeaturePlan(N, top.pairs, top.triplets, box="triclinic", describe_only_sms=148)
order = np.asarray(plan.atom_order, dtype=np.int32)         # the plan's list (per-tile blocks, repeats included)
print("selection:", len(sel), "atoms sorted; plan order:", len(order), "slots")
def slice_native(idx):
    out = np.empty((T, len(idx), 3), dtype=np.float32)
    _lib.check(_lib.lib.caviar_compact_frames_host(frames.ctypes.data, T, frames.strides[0], N, idx.ctypes.data, len(idx), 0, out.ctypes.data))
    return out
for name, idx in (("caller (sorted) order", sel), ("plan order", order)):
    best_n = best_p = 1e9
    for _ in range(5):
        t0 = time.perf_counter(); a = slice_native(idx); best_n = min(best_n, time.perf_counter() - t0)
    for _ in range(3):
        t0 = time.perf_counter(); b = frames[:, idx, :]; best_p = min(best_p, time.perf_counter() - t0)
    assert np.array_equal(a, b)
    print(f"{name:22s}: caviar_compact_frames_host {best_n*1e3:7.1f} ms ({T/best_n:7.0f} frames/s, {a.nbytes/best_n/1e9:5.2f} GB/s out)   numpy fancy index {best_p*1e3:7.1f} ms")
