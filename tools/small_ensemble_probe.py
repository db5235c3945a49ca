"""Kernel times of one posterior evaluation of the C3 star for small walker counts (what one rank of the
walker-sharded sampler evaluates per half-step at 2/4/8 GPUs: 64/32/16 walkers)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from gptransits_b200 import Engine  # noqa: E402


class A:
    walkers = 0
    gpus = 1


work = bench.build_workload("c3", 0, A)
t = work["t"]
np.random.seed(42)
gp, tr, spec = bench.model_of(work, t)
eng = Engine(0)
for kv in sys.argv[1:]:
    k, v = kv.split("=")
    eng.set_option(k, float(v))
ys = bench.inject_transit(eng, work, spec)
eng.upload_star(0, t, ys[0], work["yerr"])
theta = np.hstack([gp.sample_prior(128), tr.sample_prior(128)])
for W in (128, 64, 32, 16):
    th = np.ascontiguousarray(theta[:W])
    for _ in range(6):
        eng.log_prob(th)            # lets the halo adapt
    eng.set_option("reset_timers", 1)
    eng.set_option("time_kernels", 1)
    lp, ms = eng.log_prob_resident_timed(th, 8, flush_l2=False)
    k = {n: round(1e3 * eng.stat("ms_" + n) / max(1, eng.stat("n_" + n)), 1) for n in ("prep", "transit", "chunk", "final", "pass2")}
    eng.set_option("time_kernels", 0)
    lp, ms2 = eng.log_prob_resident_timed(th, 8, flush_l2=False)
    print(json.dumps({"W": W, "us": k, "ms_timed_kernels": float(np.median(ms)), "ms_plain": float(np.median(ms2)),
                      "chunks": eng.stat("chunks"), "L": eng.stat("chunk_len"), "halo": eng.stat("halo")}), flush=True)
eng.close()
