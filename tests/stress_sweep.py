#!/usr/bin/env python3
"""GPU box: the randomised shape sweep of tests/test_gpu_fuzz.py over N more seeds (python tests/stress_sweep.py 1500).
Round 1: 1 900 shapes, one finding -- an arm 3.8e-6 A from a Voronoi face of the lattice (two images of the same length
at float32 resolution, angle undefined): now set aside by the parity check through oracle.triplet_arm_gaps."""
import sys, os, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))   # (under tests/: the sweep checks against the oracle)
import numpy as np
import caviar_b200 as cb
import test_gpu_fuzz as tf
fails = 0
for seed in range(24, 24 + int(sys.argv[1])):
    try:
        tf.test_random_shapes_match_oracle(cb, seed)
    except Exception as e:
        fails += 1
        print("SEED", seed, "FAILED:", repr(e)[:300]); traceback.print_exc(limit=3)
print("stress done, failures:", fails)
