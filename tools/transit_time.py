"""Time the transit path alone on the C3 half-ensemble (128 walkers)."""
import os, sys, json
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
W = 128
theta = np.hstack([gp.sample_prior(W), tr.sample_prior(W)])
eng.set_option("halo", 448)
eng.log_prob(theta)
eng.set_option("reset_timers", 1)
eng.set_option("time_kernels", 1)
lp, ms = eng.log_prob_resident_timed(theta, 5, flush_l2=True)
k = {n: (eng.stat("ms_" + n) / max(1, eng.stat("n_" + n))) for n in ("transit", "trscan", "chunk")}
print(os.environ.get("GPT_LIB_PATH", "default"), json.dumps(k), "items", eng.stat("transit_items"), "lp0", lp[0])
