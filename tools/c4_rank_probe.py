"""What one rank of the C4 batch runs at 8 GPUs (128 stars x 64 walkers), for an ncu launch list."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from gptransits_b200 import Engine  # noqa: E402


class A:
    walkers = 0
    gpus = int(sys.argv[1]) if len(sys.argv) > 1 else 8


eng = Engine(0)
for kv in sys.argv[2:]:
    k, v = kv.split("=")
    eng.set_option(k, float(v))
work = bench.build_workload("c4", 0, A)
t = work["t"]
np.random.seed(42)
gp, tr, spec = bench.model_of(work, t)
ys = bench.inject_transit(eng, work, spec)
eng.upload_stars(0, t, np.stack(ys), work["yerr"])
ns, W = len(ys), work["W"]
init = bench.draw_prior(gp, tr, W * ns).reshape(ns, W, -1)
eng.sampler_begin(init, seed=42, nstars=ns, star_ids=work["star_ids"])
eng.sampler_run(4, keep=False)
ms = eng.sampler_run_timed(6, keep=False, flush_l2=True)
print("ms/step", float(ms.mean()), "chunks", eng.stat("chunks"), eng.stat("chunk_len"), eng.stat("halo"))
eng.sampler_end()
eng.close()
