"""Feasibility probe: does the transit flux kernel of a second stream fill the FP64 issue slots the
single-warp-per-scheduler chunk kernel leaves idle?  Engine A runs C3 sampler steps; engine B (own
stream, same GPU) evaluates a transit-only posterior in a loop from another thread."""
import os, sys, threading, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from gptransits_b200 import Engine, build_spec

class A:
    walkers = 0
    gpus = 1
work = bench.build_workload("c3", 0, A)
t = work["t"]
np.random.seed(42)
gp, tr, spec = bench.model_of(work, t)
engA = Engine(0)
ys = bench.inject_transit(engA, work, spec)
engA.upload_star(0, t, ys[0], work["yerr"])
W = work["W"]
init = bench.draw_prior(gp, tr, W).reshape(1, W, -1)
engA.sampler_begin(init, seed=1, nstars=1)
engA.sampler_run(40, keep=False)

engB = Engine(0)
specB = build_spec(None, tr)
engB.set_model(specB)
engB.upload_star(0, t, ys[0], work["yerr"])
WB = int(sys.argv[1]) if len(sys.argv) > 1 else 256
thB = tr.sample_prior(WB)
engB.log_prob(thB)

def run_a(n=60):
    ms = engA.sampler_run_timed(n, keep=False, flush_l2=False)
    return float(np.mean(ms))

print("A alone: %.3f ms/step" % run_a())
t0 = time.perf_counter(); nb = 0
for _ in range(50):
    engB.log_prob(thB); nb += 1
tb = (time.perf_counter() - t0) / nb
print("B alone: %.3f ms per %d-walker transit evaluation" % (tb * 1e3, WB))
stop = False; count = [0]
def loop_b():
    while not stop:
        engB.log_prob(thB); count[0] += 1
th = threading.Thread(target=loop_b); th.start()
time.sleep(0.05)
c0 = count[0]; t0 = time.perf_counter()
a = run_a(120)
dt = time.perf_counter() - t0; c1 = count[0]
stop = True; th.join()
print("A with B: %.3f ms/step; B: %.3f ms per evaluation (%d in %.3f s)" % (a, dt / max(c1 - c0, 1) * 1e3, c1 - c0, dt))
