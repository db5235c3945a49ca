"""Sweep chunk geometry / halo options on the C3 workload (device-resident log_prob batches)."""
import json
import sys
import os
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from gptransits_b200 import Engine  # noqa: E402


def main():
    class A:
        walkers = 0
        gpus = 1
    W = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    work = bench.build_workload("c3", 0, A)
    t = work["t"]
    np.random.seed(42)
    gp, tr, spec = bench.model_of(work, t)
    eng = Engine(0)
    ys = bench.inject_transit(eng, work, spec)
    eng.upload_star(0, t, ys[0], work["yerr"])
    theta = np.hstack([gp.sample_prior(W), tr.sample_prior(W)])
    peak = eng.fp64_peak_tflops()
    print("fp64 peak", peak)
    ref = None
    for chunks in (148, 222, 296):
        for halo in (320,):
            eng.set_option("chunks", chunks)
            eng.set_option("halo", halo)
            eng.log_prob(theta)
            eng.log_prob(theta)
            p2 = eng.stat("pass2_chunks")
            eng.set_option("reset_timers", 1)
            eng.set_option("time_kernels", 1)
            lp, ms = eng.log_prob_resident_timed(theta, 5, flush_l2=True)
            k = {n: (eng.stat("ms_" + n) / max(1, eng.stat("n_" + n))) for n in ("prep", "transit", "trscan", "chunk", "final", "pass2")}
            eng.set_option("time_kernels", 0)
            if ref is None:
                ref = lp
            err = np.max(np.abs(lp - ref) / np.abs(ref))
            fl = 254.0 * t.size * W
            print(json.dumps({"chunks_opt": chunks, "halo_opt": halo, "halo": eng.stat("halo"), "chunks": eng.stat("chunks"),
                              "L": eng.stat("chunk_len"), "ms_total": float(np.median(ms)), "kern_ms": k,
                              "chunk_frac": fl / (k["chunk"] * 1e-3) / 1e12 / peak,
                              "pass2_new": eng.stat("pass2_chunks") - p2, "ratio": eng.stat("verify_ratio"),
                              "relerr_vs_first": err, "items": eng.stat("transit_items")}))
    eng.close()


if __name__ == "__main__":
    main()
