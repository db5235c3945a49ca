"""A long C3 chain through the public call (gpt_star_upload + gpt_sample): sustained throughput, acceptance,
how often the chunk verification flagged anything, and the halo the adaptation settles on."""
import json, os, sys, time
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
for kv in sys.argv[2:]:
    k, v = kv.split("=")
    eng.set_option(k, float(v))
ys = bench.inject_transit(eng, work, spec)
W = work["W"]
init = bench.draw_prior(gp, tr, W).reshape(1, W, -1)
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
eng.upload_star(0, t, ys[0], work["yerr"])
eng.sample(init, 3, seed=1, nstars=1)
p0, l0 = eng.stat("pass2_chunks"), eng.stat("launches")
t0 = time.perf_counter()
eng.upload_star(0, t, ys[0], work["yerr"])
chain, logp, nacc = eng.sample(init, steps, seed=1, nstars=1)
dt = time.perf_counter() - t0
print(json.dumps({"workload": work["label"], "steps": steps, "wall_s": dt, "evals_per_s": W * steps / dt,
                  "ms_per_step": 1e3 * dt / steps, "acceptance": float(nacc.mean() / steps),
                  "flagged_chunks": eng.stat("pass2_chunks") - p0, "spec_replays": eng.stat("spec_replays"),
                  "launches_per_step": (eng.stat("launches") - l0) / steps, "halo_end": eng.stat("halo"),
                  "finite_logp_fraction": float(np.isfinite(logp).mean()), "options": sys.argv[2:]}))
