"""Two-GPU checks (skipped on a single-GPU box): a single star with the half-ensemble sharded by walker
across ranks (one NCCL all-gather of proposal log-probabilities per half-step), and stars sharded
across ranks, both against the single-GPU device sampler."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
pytestmark = pytest.mark.gpu


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _setup():
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    from conftest import KIC_CONFIG
    from gptransits_b200 import GPModel, build_spec
    with np.load(ROOT / "tests" / "golden" / "example_lightcurves.npz") as z:
        t, f, e = z["kic_time"], z["kic_flux"], z["kic_flux_err"]
    gp = GPModel(KIC_CONFIG["gp"])
    spec = build_spec(gp, None)
    np.random.seed(3)
    init = gp.sample_prior(24)
    return t, (f - 1) * 1e6, e * 1e6, spec, init


def _worker(rank, world, port, outdir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    t, y, yerr, spec, init = _setup()
    from gptransits_b200 import Engine
    from gptransits_b200.parallel import HalfEnsembleSampler, run_star_sharded, run_walker_sharded
    eng = Engine(rank)
    eng.set_model(spec)
    eng.upload_star(0, t, y, yerr)
    chain, logp = HalfEnsembleSampler(lambda th: eng.log_prob(th), init, seed=9, dist=dist).run(6)
    np.save(Path(outdir) / f"walker_{rank}.npy", chain)
    # the device-resident form of the same protocol (C ABI half-step calls + NCCL all-gather in place)
    # (the library's own NCCL communicator: proposals, slice of posteriors, all-gather, accept on one stream)
    dchain, dlogp = run_walker_sharded(eng, dist, init, 6, seed=9)
    np.save(Path(outdir) / f"dwalker_{rank}.npy", dchain)
    # ... and the half-step calls around torch.distributed's all-gather
    hchain, hlogp = run_walker_sharded(eng, dist, init, 6, seed=9, host_collective=True)
    np.save(Path(outdir) / f"hwalker_{rank}.npy", hchain)
    # stars sharded by rank: 3 stars, global ids keep the RNG streams independent of the sharding
    ys = [y + 5.0 * s for s in range(3)]
    inits = np.stack([init] * 3)
    lo, hi, c, l = run_star_sharded(eng, t, ys, yerr, inits, 5, 9, rank, world)
    if c is not None:
        np.save(Path(outdir) / f"stars_{lo}_{hi}.npy", c)
    dist.barrier()
    dist.destroy_process_group()
    eng.close()


@pytest.mark.skipif(_ngpu() < 2, reason="needs two GPUs")
def test_two_gpu_sharding(tmp_path):
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    t, y, yerr, spec, init = _setup()
    from gptransits_b200 import Engine
    eng = Engine(0)
    eng.set_model(spec)
    eng.upload_star(0, t, y, yerr)
    want, _, _ = eng.sample(init, 6, seed=9)
    c0, c1 = np.load(tmp_path / "walker_0.npy"), np.load(tmp_path / "walker_1.npy")
    assert np.array_equal(c0, c1)
    assert np.max(np.abs(c0 - want) / np.maximum(np.abs(want), 1e-300)) < 1e-9
    d0, d1 = np.load(tmp_path / "dwalker_0.npy"), np.load(tmp_path / "dwalker_1.npy")
    assert np.array_equal(d0, d1)
    assert np.max(np.abs(d0 - want) / np.maximum(np.abs(want), 1e-300)) < 1e-9
    h0, h1 = np.load(tmp_path / "hwalker_0.npy"), np.load(tmp_path / "hwalker_1.npy")
    assert np.array_equal(h0, h1)
    assert np.max(np.abs(h0 - want) / np.maximum(np.abs(want), 1e-300)) < 1e-9
    got = {}
    for p in tmp_path.glob("stars_*.npy"):
        lo, hi = map(int, p.stem.split("_")[1:])
        arr = np.load(p)
        for k in range(hi - lo):
            got[lo + k] = arr[k]
    assert sorted(got) == [0, 1, 2]
    for s in range(3):
        eng.clear_stars()
        eng.upload_star(0, t, y + 5.0 * s, yerr)
        eng.sampler_begin(init, seed=9, star_ids=[s])
        eng.sampler_run(5, keep=True)
        c, _ = eng.sampler_read(0, 5)
        eng.sampler_end()
        assert np.array_equal(c[0], got[s])
    eng.close()


def _worker_one_gpu(rank, world, port, outdir):
    """Two ranks on ONE GPU: the half-step C ABI (gpt_sampler_half_eval / half_accept) with the slices of the
    proposal log-probabilities exchanged through the host (gloo) -- the walker-sharded protocol on a box where
    NCCL cannot run (it refuses two ranks on one device)."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    t, y, yerr, spec, init = _setup()
    from gptransits_b200 import Engine
    from gptransits_b200.parallel import run_walker_sharded
    eng = Engine(0)
    eng.set_model(spec)
    eng.upload_star(0, t, y, yerr)
    chain, logp = run_walker_sharded(eng, dist, init, 6, seed=9)
    np.save(Path(outdir) / f"gwalker_{rank}.npy", chain)
    dist.barrier()
    dist.destroy_process_group()
    eng.close()


@pytest.mark.skipif(_ngpu() < 1, reason="needs a GPU")
def test_walker_sharding_two_ranks_on_one_gpu(tmp_path):
    import torch.multiprocessing as mp
    mp.spawn(_worker_one_gpu, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    t, y, yerr, spec, init = _setup()
    from gptransits_b200 import Engine
    eng = Engine(0)
    eng.set_model(spec)
    eng.upload_star(0, t, y, yerr)
    want, _, _ = eng.sample(init, 6, seed=9)
    eng.close()
    g0, g1 = np.load(tmp_path / "gwalker_0.npy"), np.load(tmp_path / "gwalker_1.npy")
    assert np.array_equal(g0, g1)
    assert np.max(np.abs(g0 - want) / np.maximum(np.abs(want), 1e-300)) < 1e-9
