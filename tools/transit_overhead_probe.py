"""What the transit handling costs inside gp_chunk_kernel: the C3 star at the sampler geometry (128 walkers,
148 chunks, halo 320) with and without the transit block of the model, and with a window that is never open."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from gptransits_b200 import Engine, GPModel, build_spec  # noqa: E402


class A:
    walkers = 0
    gpus = 1


W = 128
work = bench.build_workload("c3", 0, A)
t = work["t"]
np.random.seed(42)
gp, tr, spec = bench.model_of(work, t)
eng = Engine(0)
ys = bench.inject_transit(eng, work, spec)
eng.upload_star(0, t, ys[0], work["yerr"])
theta = np.hstack([gp.sample_prior(W), tr.sample_prior(W)])


def run(label, th):
    eng.set_option("chunks", 148)
    eng.set_option("halo", 320)
    eng.log_prob(th)
    eng.log_prob(th)
    eng.set_option("reset_timers", 1)
    eng.set_option("time_kernels", 1)
    lp, ms = eng.log_prob_resident_timed(th, 8, flush_l2=True)
    k = {n: round(1e3 * eng.stat("ms_" + n) / max(1, eng.stat("n_" + n)), 1) for n in ("prep", "transit", "chunk", "final", "pass2")}
    eng.set_option("time_kernels", 0)
    print(json.dumps({"case": label, "us": k, "ms_total": float(np.median(ms)), "items": eng.stat("transit_items")}), flush=True)


run("gp + transit", theta)
# the same walkers with a planet too small and too far to open a window wider than a sample or two
th2 = theta.copy()
th2[:, 8 + 3] = 1e4   # aR huge -> transit duration ~ 0
run("gp + transit, window (almost) never open", th2)
spec_gp = build_spec(gp, None)
eng.set_model(spec_gp)
run("gp only", theta[:, :8].copy())
eng.close()
