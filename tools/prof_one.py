"""One C3 half-ensemble evaluation (128 walkers), for ncu captures."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from gptransits_b200 import Engine

class A:
    walkers = 0
    gpus = 1
work = bench.build_workload("c3", 0, A)
t = work["t"]
np.random.seed(42)
gp, tr, spec = bench.model_of(work, t)
eng = Engine(0)
ys = bench.inject_transit(eng, work, spec)
eng.upload_star(0, t, ys[0], work["yerr"])
W = int(sys.argv[1]) if len(sys.argv) > 1 else 128
theta = np.hstack([gp.sample_prior(W), tr.sample_prior(W)])
eng.set_option("halo", int(sys.argv[2]) if len(sys.argv) > 2 else 448)
for _ in range(3):
    lp = eng.log_prob(theta)
print("ok", lp[:2])
