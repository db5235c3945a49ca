"""world_size-2 gloo runs of the multi-GPU host logic on CPU (the per-walker evaluation is the
oracle here; on GPUs it is the CUDA engine)."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, outdir):
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    import oracle
    from conftest import KIC_CONFIG
    from gptransits_b200 import GPModel, build_spec
    from gptransits_b200.parallel import HalfEnsembleSampler, gather_chains, star_block
    dist.init_process_group("gloo", rank=rank, world_size=world)
    with np.load(ROOT / "tests" / "golden" / "example_lightcurves.npz") as z:
        t, f, e = z["kic_time"][:160], z["kic_flux"][:160], z["kic_flux_err"][:160]
    gp = GPModel(KIC_CONFIG["gp"])
    spec = build_spec(gp, None)
    np.random.seed(3)
    init = gp.sample_prior(12)
    y, yerr = (f - 1) * 1e6, e * 1e6
    calls = []

    def fn(theta):
        calls.append(theta.shape[0])
        return oracle.log_prob(spec, t, y, yerr, theta)

    # (1) single star, half-ensemble sharded by walker with one all-gather per half-step
    chain, logp = HalfEnsembleSampler(fn, init, seed=9, dist=dist).run(5)
    np.save(Path(outdir) / f"chain_{rank}.npy", chain)
    np.save(Path(outdir) / f"calls_{rank}.npy", np.array(calls))

    # (2) star sharding: 5 stars over 2 ranks, no collective until the final gather
    nstars = 5
    lo, hi = star_block(nstars, rank, world)
    local_chain, local_logp = [], []
    for s in range(lo, hi):
        ys = y + 10.0 * s
        c, l, _ = oracle.sample(spec, t, ys, yerr, init, 3, seed=9, star=s)
        local_chain.append(c)
        local_logp.append(l)
    gathered = gather_chains(dist, (lo, hi, local_chain, local_logp), nstars, rank, world)
    if rank == 0:
        np.save(Path(outdir) / "stars.npy", np.stack([g[0] for g in gathered]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_sharding(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    sys.path.insert(0, str(ROOT))
    import oracle
    from conftest import KIC_CONFIG
    from gptransits_b200 import GPModel, build_spec
    with np.load(ROOT / "tests" / "golden" / "example_lightcurves.npz") as z:
        t, f, e = z["kic_time"][:160], z["kic_flux"][:160], z["kic_flux_err"][:160]
    gp = GPModel(KIC_CONFIG["gp"])
    spec = build_spec(gp, None)
    np.random.seed(3)
    init = gp.sample_prior(12)
    y, yerr = (f - 1) * 1e6, e * 1e6
    want, _, _ = oracle.sample(spec, t, y, yerr, init, 5, seed=9)
    c0, c1 = np.load(tmp_path / "chain_0.npy"), np.load(tmp_path / "chain_1.npy")
    assert np.array_equal(c0, c1)          # replicated state never diverges
    assert np.array_equal(c0, want)        # and equals the unsharded run: independent of world size
    calls = np.load(tmp_path / "calls_0.npy")
    assert calls[0] == 6 and set(calls[1:]) == {3}   # each rank evaluates half of each half-ensemble
    stars = np.load(tmp_path / "stars.npy")
    for s in range(5):
        c, _, _ = oracle.sample(spec, t, y + 10.0 * s, yerr, init, 3, seed=9, star=s)
        assert np.array_equal(stars[s], c)
