#!/usr/bin/env python
"""bench.py -- log-likelihood evaluations/s of the gptransits posterior on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c3|c4|c5|c2]

A "step" is ONE ENSEMBLE STEP of the device-resident stretch-move sampler on the workload star(s):
every walker is proposed and its full posterior (priors + supersampled Mandel-Agol transit +
celerite log-likelihood over all N samples) evaluated once, in two half-ensemble launches, with
accept/reject and the chain append -- i.e. W evaluations per star per step.

  value   evaluations/s with everything resident in HBM, timed per step with CUDA events on the
          engine's stream, L2 flushed between steps (256 MiB overwrite outside the timed intervals).
  e2e     the same metric through the C ABI with HOST buffers: gpt_star_upload + gpt_sample(init,
          K steps) -> host chain/log-prob arrays, wall clock around the calls.
  roofline  the dominant kernel (gp_chunk_kernel) against the FP64 FMA peak measured on this GPU by
          a register-resident DFMA kernel (MEASURED_PEAKS.json has no FP64 entry); FLOP counts are
          SURVEY.md section 8(d)'s canonical per-sample figures.
  cpu_baseline  the C oracle (oracle/gpt_oracle.c, a port of celerite+batman arithmetic) on the host
          cores, bounded sample.

Multi-GPU (torchrun): the path shards by star with no data-path collective; each rank samples its
own synthetic star of the same shape (weak scaling); time = max over ranks.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

F_GP = {2: 4 * 16 + 17.5 * 4 + 5, 4: 139.0, 6: 254.0, 8: 4 * 64 + 17.5 * 8 + 5}   # 4J^2+17.5J+5
F_TR = 90.0                                                                      # 6 supersamples x 15


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=["c1", "c2", "c3", "c4", "c5"])
    ap.add_argument("--walkers", type=int, default=0)
    ap.add_argument("--no-flush", action="store_true")
    ap.add_argument("--cpu-evals", type=int, default=0, help="size of the bounded CPU sample")
    ap.add_argument("--seed", type=int, default=42)
    ap.add_argument("--opt", action="append", default=[], metavar="KEY=VALUE",
                    help="engine option (gpt_set_option), e.g. --opt pdl=0; recorded in config")
    ap.add_argument("--no-splits", action="store_true", help="skip the c4_strong / walker_sharded sub-records")
    ap.add_argument("--batch-walkers", type=int, default=2048,
                    help="size of the extra resident log_prob batch (0 = skip)")
    return ap.parse_args()


# --------------------------------------------------------------------------- workloads
def build_workload(name, rank, args):
    """-> dict(label, t, y_list, yerr, config, W, N, nstars, J)   (transit still to be injected)"""
    from gptransits_b200 import synth
    cfg = synth.c3_config()
    if name == "c3":
        t, y, s = synth.make_star(65536, seed=2452144 + rank)
        return dict(label="C3 synthetic Kepler long-cadence N=65536, 2 granulation + oscillation bump + white noise "
                          "+ transit, 256 walkers", t=t, ys=[y], yerr=s, config=cfg, W=args.walkers or 256, J=6)
    if name == "c5":
        t, y, s = synth.make_star(1048576, seed=2452144 + rank, cadence="tess")
        return dict(label="C5 synthetic TESS 2-min N=1048576, 32 walkers", t=t, ys=[y], yerr=s, config=cfg,
                    W=args.walkers or 32, J=6)
    if name == "c4":
        world = max(1, args.gpus)
        per = 1024 // world
        t, y, s, th = synth.make_stars(1024, 4096)
        lo = rank * per
        return dict(label=f"C4 1024 synthetic stars x N=4096 x 64 walkers, {per} stars per GPU", t=t,
                    ys=list(y[lo:lo + per]), yerr=s, config=cfg, W=args.walkers or 64, J=6, star_ids=np.arange(lo, lo + per))
    if name == "c1":
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from conftest import KIC_CONFIG
        with np.load(os.path.join(ROOT, "tests", "golden", "example_lightcurves.npz")) as z:
            t, f, e = z["kic_time"], z["kic_flux"], z["kic_flux_err"]
        return dict(label="C1 examples/kic-012008916 (N=1292), GP only, 24 walkers", t=t, ys=[f * 1e6 - 1e6],
                    yerr=e * 1e6, config=KIC_CONFIG, W=args.walkers or 24, J=4, no_inject=True)
    if name == "c2":
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from conftest import SIM_CONFIG
        with np.load(os.path.join(ROOT, "tests", "golden", "example_lightcurves.npz")) as z:
            t, f = z["sim_time"], z["sim_flux"]
        return dict(label="C2 examples/sim-2452144 (N=1320), GP + transit, 64 walkers", t=t, ys=[f * 1e6 - 1e6],
                    yerr=None, config=SIM_CONFIG, W=args.walkers or 64, J=4, no_inject=True)
    raise ValueError(name)


def model_of(work, t):
    from gptransits_b200 import BatmanModel, GPModel, build_spec
    gp = GPModel(work["config"]["gp"])
    tr = None
    if work["config"].get("transit"):
        tr = BatmanModel(work["config"]["transit"][0])
        tr.init_model(t, t[1] - t[0], 6)
    return gp, tr, build_spec(gp, tr)


def draw_prior(gp, tr, n):
    """n walkers from the priors, GP block then transit block (model.py:257-266)."""
    blocks = [gp.sample_prior(n)]
    if tr is not None:
        blocks.append(tr.sample_prior(n))
    return np.hstack(blocks)


def inject_transit(eng, work, spec):
    """y = 1e6 + gp + (flux - 1) * 1e6 with the truth transit computed by the engine's own flux kernel
    (baseline kept, as the reference scales the file flux: model.py:164)."""
    from gptransits_b200 import synth
    eng.set_model(spec)
    t = work["t"]
    if work.get("no_inject"):
        return [y + 1e6 for y in work["ys"]]
    eng.upload_star(0, t, work["ys"][0], work["yerr"])
    pars = np.concatenate([synth.TRUTH_TR, [synth.TR_FIXED["e"], synth.TR_FIXED["w"], synth.TR_FIXED["u1"],
                                             synth.TR_FIXED["u2"]]])
    flux = eng.transit_flux(pars[None, :])[0]
    return [1e6 + y + (flux - 1) * 1e6 for y in work["ys"]]


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            p = [x.strip() for x in r.split(",")]
            try:
                sm.append(float(p[0]))
                smax.append(float(p[1]))
                power.append(float(p[2]))
                for nm, v in zip(names, p[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------- reference arm / cpu baseline
def cpu_evals_per_s(spec, t, y, yerr, theta, nthreads):
    import oracle
    t0 = time.perf_counter()
    oracle.log_prob(spec, t, y, yerr, theta, nthreads=nthreads)
    dt = time.perf_counter() - t0
    return theta.shape[0] / dt, dt


def run_reference(args):
    """The reference's CPU implementation of the path on the host cores.  Its own packages (celerite,
    batman, emcee) are not installable here, so this is the oracle port with all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle
    from gptransits_b200 import synth
    oracle.build()
    work = build_workload(args.workload, 0, args)
    t = work["t"]
    np.random.seed(args.seed)
    gp, tr, spec = model_of(work, t)
    # the oracle's own light curve supplies the injected truth transit (no GPU in this arm)
    if work.get("no_inject"):
        y = work["ys"][0] + 1e6
    else:
        pars = np.concatenate([synth.TRUTH_TR, [0.0, 90.0, 0.4996, 0.0847]])
        flux = oracle.light_curve(t, pars, 6, spec["exp_time"])
        y = 1e6 + work["ys"][0] + (flux - 1) * 1e6
    cores = os.cpu_count() or 1
    W = work["W"]
    # one step = the evaluations one ensemble step of our arm makes on one GPU (W per star); the stars of a batch
    # share the model, so star 0 stands for all of them
    nsample = args.cpu_evals or W * len(work["ys"])
    theta = draw_prior(gp, tr, min(nsample, 4096))
    theta = theta[np.resize(np.arange(theta.shape[0]), nsample)]
    for _ in range(args.warmup):
        cpu_evals_per_s(spec, t, y, work["yerr"], theta[:cores], cores)
    times = []
    for _ in range(args.steps):
        _, dt = cpu_evals_per_s(spec, t, y, work["yerr"], theta, cores)
        times.append(dt)
    total = float(np.sum(times))
    value = nsample * args.steps / total
    sample = (f"{nsample} posterior evaluations per step (one ensemble step of the {args.workload} workload on one GPU; "
              f"star 0 stands for every star), {cores} OpenMP threads, oracle port of celerite+batman arithmetic")
    out = {"impl": "reference", "metric": "log-likelihood evals/s (walkers x steps)", "value": value,
           "unit": "evals/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f64", "data": "synthetic", "config": {"workload": work["label"], "evals_per_step": nsample},
           "cpu_baseline": {"value": value, "unit": "evals/s", "cores": cores, "kind": "port", "sample": sample},
           "e2e": {"value": value, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out))


def hbm_view(algorithmic_bytes, kernel_ms):
    """The same kernel against the measured HBM copy bandwidth (MEASURED_PEAKS.json, else the
    profiling recipe's fallback): shows that the path is not memory bound."""
    peak, src = 6650.0, "fallback"
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
        src = "measured"
    except Exception:
        pass
    if not kernel_ms:
        return None
    gbs = algorithmic_bytes / (kernel_ms * 1e-3) / 1e9
    return {"achieved_gbs": gbs, "peak_gbs": peak, "peak_source": src, "frac": gbs / peak}


# --------------------------------------------------------------------------- parity at the bench geometry
def parity_check(eng, spec, t, ys, yerr, x, lpx, W, nstars, per_star=16, max_stars=8):
    """The final ensemble of the timed run against the CPU oracle, outside the timed region: the first half of
    every star's walkers is evaluated in one launch of the sampler's own shape (W/2 walkers per star, all stars:
    the same chunk geometry and halo) and `per_star` of them on up to `max_stars` stars are compared term by
    term; the log-probabilities the sampler itself stored for those walkers are compared too."""
    import oracle
    H = W // 2
    ndim = x.shape[-1]
    th = np.ascontiguousarray(x[:, :H, :])
    _, parts = eng.log_prob(th, nstars=nstars, return_parts=True)
    geom = {k: eng.stat(k) for k in ("chunks", "chunk_len", "halo")}
    stars = np.unique(np.linspace(0, nstars - 1, min(nstars, max_stars)).astype(int))
    sel = np.unique(np.linspace(0, H - 1, min(H, per_star)).astype(int))
    worst = {"quad": 0.0, "logdet": 0.0, "lnprior": 0.0, "lp_sampler": 0.0}
    n = 0
    for s in stars:
        want_lp, want = oracle.log_prob(spec, t, ys[s], yerr, th[s, sel], return_parts=True, nthreads=os.cpu_count() or 1)
        got = parts[s, sel]
        fin = np.isfinite(want[:, 3])
        if not np.array_equal(np.isfinite(got[:, 3]), fin):
            worst["quad"] = float("inf")
        for name, col in (("lnprior", 0), ("quad", 1), ("logdet", 2)):
            if fin.any():
                d = np.abs(got[fin, col] - want[fin, col]) / np.maximum(np.abs(want[fin, col]), 1e-300)
                worst[name] = max(worst[name], float(d.max()))
        f2 = np.isfinite(want_lp)
        if f2.any():
            d = np.abs(lpx[s, sel][f2] - want_lp[f2]) / np.maximum(np.abs(want_lp[f2]), 1e-300)
            worst["lp_sampler"] = max(worst["lp_sampler"], float(d.max()))
        n += int(sel.size)
    ok = all(v < 1e-9 for v in worst.values())
    return {"n": n, "stars": int(stars.size), "max_rel_quad": worst["quad"], "max_rel_logdet": worst["logdet"],
            "max_rel_lnprior": worst["lnprior"], "max_rel_lp_sampler": worst["lp_sampler"], "tolerance": 1e-9,
            "chunks": geom["chunks"], "chunk_len": geom["chunk_len"], "halo": geom["halo"], "ok": bool(ok)}


def chain_hash(a):
    import hashlib
    return hashlib.sha256(np.ascontiguousarray(np.round(a, 6)).tobytes()).hexdigest()[:16]


def north_star_splits(eng, args, rank, world, dist, barrier, max_over_ranks):
    """The two multi-GPU splits BASELINE.json names, measured in the same run at every N (N = 1 included, so the
    scaling of each is value_N / value_1):
      c4_strong       1024 stars x 64 walkers x N = 4096 sharded by star, no collective: evals/s of the whole job
      walker_sharded  ONE C3 star, 256 walkers, the active half-ensemble sharded by walker; the proposal
                      log-probabilities are all-gathered over NVLink (NCCL on the engine's stream): ms per step,
                      and a hash of the chain that must not depend on N"""
    from gptransits_b200 import synth
    from gptransits_b200.parallel import run_walker_sharded
    out = {}
    steps, warm = 5, 3
    # ---- C4 strong
    class A4:
        walkers = 0
        gpus = world
    work = build_workload("c4", rank, A4)
    t = work["t"]
    np.random.seed(args.seed)
    gp, tr, spec = model_of(work, t)
    eng.clear_stars()
    ys = inject_transit(eng, work, spec)
    for s_, y in enumerate(ys):
        eng.upload_star(s_, t, y, work["yerr"])
    ns, W = len(ys), work["W"]
    ndim = spec["prior_kind"].size
    np.random.seed(args.seed + 17 + rank)
    init = draw_prior(gp, tr, W * ns).reshape(ns, W, ndim)
    eng.sampler_begin(init, seed=args.seed, nstars=ns, star_ids=work["star_ids"])
    eng.sampler_run(warm, keep=False)
    barrier()
    ms = eng.sampler_run_timed(steps, keep=False, flush_l2=True)
    eng.sampler_end()
    tot = max_over_ranks(float(ms.sum()))
    out["c4_strong"] = {"value": 1024 * W * steps / (tot * 1e-3), "unit": "evals/s", "ms_per_step": tot / steps,
                        "stars_per_gpu": ns, "steps": steps, "scaling": "strong", "collective": None}
    # ---- one C3 star, walkers sharded
    class A3:
        walkers = 0
        gpus = 1
    work = build_workload("c3", 0, A3)     # the same star on every rank
    t = work["t"]
    np.random.seed(args.seed)
    gp, tr, spec = model_of(work, t)
    eng.clear_stars()
    ys = inject_transit(eng, work, spec)
    eng.upload_star(0, t, ys[0], work["yerr"])
    W = work["W"]
    np.random.seed(args.seed + 5)
    init = draw_prior(gp, tr, W)
    run_walker_sharded(eng, dist, init, warm, seed=args.seed)          # warm-up (halo, communicator)
    barrier()
    chain, logp, ms = run_walker_sharded(eng, dist, init, steps, seed=args.seed, timed=True)
    tot = max_over_ranks(float(ms.sum()))
    h = chain_hash(chain)
    same = True
    if dist is not None:
        hs = [None] * world
        dist.all_gather_object(hs, h)
        same = len(set(hs)) == 1
    out["walker_sharded"] = {"ms_per_step": tot / steps, "value": W * steps / (tot * 1e-3), "unit": "evals/s",
                             "walkers": W, "steps": steps, "chain_hash": h, "hash_equal_across_ranks": bool(same),
                             "collective": "ncclAllGather of W/2 doubles, twice per step, on the engine's stream"}
    return out


# --------------------------------------------------------------------------- our arm
def run_ours(args):
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        # NCCL may print its version banner on stdout at the first collective: keep stdout for the
        # single JSON line by pointing fd 1 at stderr while the communicator is brought up
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    from gptransits_b200 import Engine

    work = build_workload(args.workload, rank, args)
    t = work["t"]
    N, W, J = t.size, work["W"], work["J"]
    np.random.seed(args.seed + rank)
    gp, tr, spec = model_of(work, t)
    eng = Engine(local_rank)
    for kv in args.opt:
        k, v = kv.split("=")
        eng.set_option(k, float(v))
    ys = inject_transit(eng, work, spec)
    nstars = len(ys)
    for s, y in enumerate(ys):
        eng.upload_star(s, t, y, work["yerr"])
    ndim = spec["prior_kind"].size
    init = draw_prior(gp, tr, W * nstars).reshape(nstars, W, ndim)
    star_ids = work.get("star_ids")
    flush = not args.no_flush
    peak_tf = eng.fp64_peak_tflops()
    peak_nom = eng.stat("fp64_nominal_tflops")   # SMs x 64 lanes x 2 FLOP x max SM clock

    def barrier():
        if dist is not None:
            dist.barrier()

    def max_over_ranks(x):
        if dist is None:
            return x
        import torch
        v = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(v, op=dist.ReduceOp.MAX)
        return float(v.item())

    # ---- warm-up + timed region (device resident)
    eng.sampler_begin(init, seed=args.seed, nstars=nstars, star_ids=star_ids)
    eng.sampler_run(max(3, args.warmup), keep=True)
    barrier()
    clocks = ClockSampler(local_rank)
    clocks.start()
    l0 = eng.stat("launches")
    ms = eng.sampler_run_timed(args.steps, keep=True, flush_l2=flush)
    launches = int(eng.stat("launches") - l0)
    barrier()
    geom = {k: eng.stat(k) for k in ("chunks", "chunk_len", "halo")}
    total_ms = max_over_ranks(float(ms.sum()))
    evals_per_step = W * nstars * world
    value = evals_per_step * args.steps / (total_ms * 1e-3)

    # ---- second pass with per-kernel events: duration of the dominant kernel for the roofline
    eng.set_option("reset_timers", 1)
    eng.set_option("time_kernels", 1)
    ms2 = eng.sampler_run_timed(args.steps, keep=True, flush_l2=flush)
    kern = {k: (eng.stat("ms_" + k), eng.stat("n_" + k)) for k in ("prep", "transit", "trscan", "chunk", "final", "pass2")}
    eng.set_option("time_kernels", 0)
    clk = clocks.stop()
    xfin, lpfin, nacc = eng.sampler_state()
    eng.sampler_end()
    chunk_ms, chunk_n = kern["chunk"]
    walkers_per_launch = (W // 2) * nstars
    flops_per_launch = F_GP[J] * N * walkers_per_launch
    achieved_tf = flops_per_launch / (chunk_ms / max(chunk_n, 1) * 1e-3) / 1e12 if chunk_n else None
    f_tr = F_TR if tr is not None else 0.0
    step_tf = (F_GP[J] + f_tr) * N * W * nstars / (float(ms.mean()) * 1e-3) / 1e12
    # HBM view of the same kernel: algorithmic bytes = star data once per walker group + carries
    bytes_per_launch = 32.0 * N * nstars + walkers_per_launch * (8 * ndim + 32)
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get(args.workload)
        except Exception:
            traffic = None

    # ---- end to end through the C ABI with host buffers (one untimed call first: allocations)
    eng.sample(init, 2, seed=args.seed, nstars=nstars)
    barrier()
    ymat = np.ascontiguousarray(np.stack(ys)) if nstars > 1 else None   # the batch as the caller holds it: [nstars, N]
    e2e_reps = []
    for _ in range(3):      # median of three identical calls (host jitter); every call uploads, samples and reads back
        barrier()
        t0 = time.perf_counter()
        if nstars > 1:
            eng.upload_stars(0, t, ymat, work["yerr"])      # one time grid for the whole batch (gpt_stars_upload)
        else:
            eng.upload_star(0, t, ys[0], work["yerr"])
        chain, logp, _ = eng.sample(init, args.steps, seed=args.seed, nstars=nstars)
        e2e_reps.append(max_over_ranks(time.perf_counter() - t0))
    e2e_s = float(np.median(e2e_reps))
    # the same call with the time grid analysed from scratch (gpt_star_upload keeps the analysis of the grid when the
    # time stamps of a re-upload are bit-identical, as they are in the three calls above)
    eng.clear_stars()
    barrier()
    t0 = time.perf_counter()
    if nstars > 1:
        eng.upload_stars(0, t, ymat, work["yerr"])
    else:
        eng.upload_star(0, t, ys[0], work["yerr"])
    chain, logp, _ = eng.sample(init, args.steps, seed=args.seed, nstars=nstars)
    e2e_cold_s = max_over_ranks(time.perf_counter() - t0)
    e2e_value = evals_per_step * args.steps / e2e_s
    # bytes that cross the bus: the first star's assembled blob (samples + times), the fluxes of the others, the ensemble
    h2d = (40.0 * N + 8.0 * N * (nstars - 1) + init.nbytes) / args.steps
    d2h = (chain.nbytes + logp.nbytes) / args.steps
    # posterior seam (what emcee would call): one vectorised log_prob on host theta
    th = init.reshape(-1, ndim)
    eng.log_prob(th, nstars=nstars)
    t0 = time.perf_counter()
    reps = 5
    for _ in range(reps):
        eng.log_prob(th, nstars=nstars)
    lp_rate = th.shape[0] * reps / (time.perf_counter() - t0)
    stats = {k: eng.stat(k) for k in ("pass2_chunks", "verify_ratio")}
    stats.update(geom)

    # ---- large resident batch of the same star: long chunks, little halo overhead (kernel efficiency view)
    batched = None
    if world == 1 and args.batch_walkers > 0:
        WB = args.batch_walkers
        np.random.seed(args.seed + 1000)
        thb = draw_prior(gp, tr, WB)
        eng.log_prob(thb[:WB])
        eng.log_prob(thb[:WB])
        eng.set_option("reset_timers", 1)
        eng.set_option("time_kernels", 1)
        _, msb = eng.log_prob_resident_timed(thb, 5, flush_l2=flush)
        kb = {k: eng.stat("ms_" + k) / max(1.0, eng.stat("n_" + k)) for k in ("transit", "chunk", "final", "pass2")}
        eng.set_option("time_kernels", 0)
        fb = F_GP[J] * N * WB
        batched = {"walkers": WB, "evals_per_s": WB / (float(np.median(msb)) * 1e-3), "ms": float(np.median(msb)),
                   "kernel_ms": kb, "chunk_tflops": fb / (kb["chunk"] * 1e-3) / 1e12,
                   "chunk_frac": fb / (kb["chunk"] * 1e-3) / 1e12 / peak_tf,
                   "whole_eval_frac": (F_GP[J] + f_tr) * N * WB / (float(np.median(msb)) * 1e-3) / 1e12 / peak_tf,
                   "chunks": eng.stat("chunks"), "chunk_len": eng.stat("chunk_len"), "halo": eng.stat("halo")}

    # ---- parity of the final ensemble at the bench geometry (outside every timed region)
    parity = None
    if rank == 0:
        import oracle
        oracle.build()
        parity = parity_check(eng, spec, t, ys, work["yerr"], xfin, lpfin, W, nstars)

    splits = None
    if not args.no_splits and args.workload == "c3":
        splits = north_star_splits(eng, args, rank, world, dist, barrier, max_over_ranks)

    cpu = None
    if rank == 0 and world == 1:
        import oracle
        oracle.build()
        cores = os.cpu_count() or 1
        # bounded sample: a pilot sizes it to about 8 s of wall time on all cores (theta rows are reused)
        pilot = th[:max(cores, 8)]
        vp, _ = cpu_evals_per_s(spec, t, ys[0], work["yerr"], pilot, cores)
        nsample = args.cpu_evals or int(min(max(8.0 * vp, 2 * cores), 200000))
        rows = np.resize(np.arange(th.shape[0]), nsample)
        v, dt = cpu_evals_per_s(spec, t, ys[0], work["yerr"], th[rows], cores)
        n1 = max(2, min(nsample // cores, int(2.0 * vp / cores) + 2))
        v1, dt1 = cpu_evals_per_s(spec, t, ys[0], work["yerr"], th[rows[:n1]], 1)
        cpu = {"value": v, "unit": "evals/s", "cores": cores, "kind": "port",
               "sample": f"{nsample} posterior evaluations of star 0 of this workload in {dt:.2f} s on {cores} threads; "
                         f"single thread: {v1:.2f} evals/s",
               "single_thread": v1}

    if rank == 0:
        out = {
            "metric": "log-likelihood evals/s (walkers x steps)", "value": value, "unit": "evals/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "strong" if args.workload == "c4" else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": work["label"], "N": N, "walkers": W, "stars_per_gpu": nstars, "ndim": ndim, "J": J,
                       "evals_per_step": evals_per_step, "l2": "flushed between steps" if flush else "not flushed",
                       "sharding": "by star, no collective" if world > 1 else "single GPU",
                       "chunks": stats["chunks"], "chunk_len": stats["chunk_len"], "halo": stats["halo"],
                       **({"engine_options": args.opt} if args.opt else {})},
            "sampler_steps_per_s": args.steps / (total_ms * 1e-3),
            "acceptance_fraction": float(nacc.mean() / (max(3, args.warmup) + 2 * args.steps)),
            "e2e": {"value": e2e_value, "unit": "evals/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "call": "gpt_star(s)_upload + gpt_sample (host init -> host chain/log_prob); median of 3 calls",
                    "calls_s": e2e_reps,
                    "value_with_grid_analysis": evals_per_step * args.steps / e2e_cold_s,
                    "note": "the three calls re-upload the star on unchanged time stamps, so gpt_star_upload keeps its analysis "
                            "of the time grid (the reference's one-time celerite GP.compute); value_with_grid_analysis is the "
                            "same call after gpt_star_clear",
                    "log_prob_seam_evals_per_s": lp_rate},
            "gpu_launches": launches,
            "roofline": {"bound": "fp64", "kernel": "gp_chunk_kernel", "achieved": achieved_tf, "peak": peak_tf,
                         "unit": "TFLOP/s", "frac": (achieved_tf / peak_tf) if achieved_tf else None,
                         "peak_source": "measured on this GPU by gpt_measure_fp64_peak (DFMA kernel); "
                                        "MEASURED_PEAKS.json has no FP64 entry",
                         "peak_nominal": peak_nom, "frac_of_nominal": (achieved_tf / peak_nom) if achieved_tf and peak_nom else None,
                         "flops_per_launch": flops_per_launch, "kernel_ms": chunk_ms / max(chunk_n, 1),
                         "kernel_launches_timed": chunk_n, "step_achieved": step_tf, "step_frac": step_tf / peak_tf,
                         "hbm_algorithmic_bytes_per_launch": bytes_per_launch, "traffic": traffic,
                         "hbm_view": hbm_view(bytes_per_launch, chunk_ms / max(chunk_n, 1)),
                         "kernel_share_of_step": {k: v[0] / float(ms2.sum()) for k, v in kern.items()},
                         "kernel_us_per_launch": {k: 1e3 * v[0] / max(v[1], 1) for k, v in kern.items()},
                         "pass2_chunks": stats["pass2_chunks"], "verify_ratio": stats["verify_ratio"]},
            "clocks": clk,
        }
        out["parity_check"] = parity
        if splits:
            out["north_star_splits"] = splits
        if batched:
            out["batched_log_prob"] = batched
        if cpu:
            out["cpu_baseline"] = cpu
        print(json.dumps(out))
    eng.close()
    if dist is not None:
        dist.destroy_process_group()
    if rank == 0 and parity is not None and not parity["ok"]:
        sys.exit("parity check against the oracle failed: " + json.dumps(parity))


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
