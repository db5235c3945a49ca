"""Where the end-to-end time of the C4 batch goes: star uploads, sampler run, chain read (GPU box)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from gptransits_b200 import Engine  # noqa: E402


class A:
    walkers = 0
    gpus = 1


wl = sys.argv[1] if len(sys.argv) > 1 else "c4"
steps = 20
work = bench.build_workload(wl, 0, A)
t = work["t"]
np.random.seed(42)
gp, tr, spec = bench.model_of(work, t)
eng = Engine(0)
ys = bench.inject_transit(eng, work, spec)
ns, W = len(ys), work["W"]
ndim = spec["prior_kind"].size
init = bench.draw_prior(gp, tr, W * ns).reshape(ns, W, ndim)
for rep in range(3):
    t0 = time.perf_counter()
    for s, y in enumerate(ys):
        eng.upload_star(s, t, y, work["yerr"])
    t1 = time.perf_counter()
    eng.sampler_begin(init, seed=42, nstars=ns)
    t2 = time.perf_counter()
    eng.sampler_run(steps, keep=True)
    t3 = time.perf_counter()
    chain, logp = eng.sampler_read(0, steps)
    t4 = time.perf_counter()
    eng.sampler_end()
    t5 = time.perf_counter()
    c2, l2, _ = eng.sample(init, steps, seed=42, nstars=ns)
    t6 = time.perf_counter()
    print(f"rep {rep}: upload {1e3*(t1-t0):.1f} ms, begin {1e3*(t2-t1):.1f}, run {1e3*(t3-t2):.1f}, read {1e3*(t4-t3):.1f} "
          f"({(chain.nbytes+logp.nbytes)/1e6:.0f} MB), end {1e3*(t5-t4):.1f}; gpt_sample call {1e3*(t6-t5):.1f} ms", flush=True)
eng.close()
