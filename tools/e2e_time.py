"""Breakdown of the end-to-end call (gpt_star_upload + gpt_sample with host buffers) at C3."""
import os, sys, time
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
W = work["W"]
init = bench.draw_prior(gp, tr, W).reshape(1, W, -1)
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
eng.sample(init, 2, seed=1, nstars=1)
for rep in range(4):
    t0 = time.perf_counter()
    eng.upload_star(0, t, ys[0], work["yerr"])
    t1 = time.perf_counter()
    eng.sample(init, steps, seed=1, nstars=1)
    t2 = time.perf_counter()
    print(f"rep {rep}: upload {1e3*(t1-t0):.2f} ms, sample({steps}) {1e3*(t2-t1):.2f} ms, "
          f"{W*steps/(t2-t0):.0f} evals/s; halo {eng.stat('halo')}, pass2 {eng.stat('pass2_chunks')}")
