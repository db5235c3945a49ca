"""Per-GPU work of the C4 batch at 1/2/4/8 GPUs, emulated on one GPU: 1024/k stars x 64 walkers, kernel times."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from gptransits_b200 import Engine  # noqa: E402

eng = Engine(0)
for k in (1, 2, 4, 8):
    class A:
        walkers = 0
        gpus = k
    work = bench.build_workload("c4", 0, A)
    t = work["t"]
    np.random.seed(42)
    gp, tr, spec = bench.model_of(work, t)
    eng.clear_stars()
    ys = bench.inject_transit(eng, work, spec)
    for s, y in enumerate(ys):
        eng.upload_star(s, t, y, work["yerr"])
    ns, W = len(ys), work["W"]
    ndim = spec["prior_kind"].size
    init = bench.draw_prior(gp, tr, W * ns).reshape(ns, W, ndim)
    for opt in ([], [("chunks", 1)], [("chunks", 2)], [("chunks", 4)]):
        eng.set_option("chunks", 0)
        for kk, v in opt:
            eng.set_option(kk, v)
        eng.sampler_begin(init, seed=42, nstars=ns, star_ids=work["star_ids"])
        eng.sampler_run(3, keep=False)
        ms = eng.sampler_run_timed(5, keep=False, flush_l2=True)
        eng.set_option("reset_timers", 1)
        eng.set_option("time_kernels", 1)
        eng.sampler_run_timed(5, keep=False, flush_l2=True)
        kern = {n: round(1e3 * eng.stat("ms_" + n) / max(1, eng.stat("n_" + n)), 1) for n in ("prep", "transit", "trscan", "chunk", "final", "pass2")}
        eng.set_option("time_kernels", 0)
        eng.sampler_end()
        print(json.dumps({"gpus": k, "stars": ns, "opt": opt, "ms_per_step": round(float(ms.mean()), 3),
                          "evals_per_s_x_gpus": round(k * ns * W / (float(ms.mean()) * 1e-3)), "chunks": eng.stat("chunks"),
                          "L": eng.stat("chunk_len"), "halo": eng.stat("halo"), "us": kern}), flush=True)
eng.close()
