"""True chunk-coupling error vs the verification criterion's ratio, as a function of the halo."""
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
eng.set_option("chunks", 1)
_, exact = eng.log_prob(theta, return_parts=True)
eng.set_option("chunks", 148)
for H in (96, 128, 192, 256, 320, 384, 448):
    eng.set_option("halo", H)
    eng.set_option("verify", 0)
    _, p = eng.log_prob(theta, return_parts=True)
    eq = np.max(np.abs(p[:, 1] - exact[:, 1]) / np.abs(exact[:, 1]))
    el = np.max(np.abs(p[:, 2] - exact[:, 2]) / np.abs(exact[:, 2]))
    eng.set_option("verify", 1)
    b0 = eng.stat("pass2_chunks")
    eng.log_prob(theta)
    print(json.dumps({"halo": H, "true_rel_err_quad": eq, "true_rel_err_logdet": el, "criterion_ratio": eng.stat("verify_ratio"),
                      "flagged_chunks": eng.stat("pass2_chunks") - b0}))
