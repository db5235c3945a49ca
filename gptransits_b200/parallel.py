"""Multi-GPU sharding of the posterior/sampler path: one process per GPU, stars are independent.

The path shards by star: rank r owns a contiguous block of stars, samples
them with no data-path collective, and the chains are gathered once at the end.  The Philox
stream is keyed by the GLOBAL star id, so results are identical for any number of GPUs.

For a single star the ensemble can instead be split by walker: each rank evaluates its slice of
the active half-ensemble's proposals and the slices are all-gathered (the only collective, twice
per step).  `HalfEnsembleSampler` implements that protocol over torch.distributed (NCCL on GPUs,
gloo in the CPU tests); the per-walker evaluation is a callable, so the same host logic drives
the CUDA engine in production and the oracle in tests.
"""
import numpy as np


def star_block(nstars, rank, world):
    """Contiguous block [lo, hi) of stars owned by `rank`; sizes differ by at most one."""
    base, extra = divmod(nstars, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def walker_slice(nactive, rank, world):
    return star_block(nactive, rank, world)


def run_star_sharded(engine, t, ys, yerr, init, nsteps, seed, rank, world):
    """Sample every star owned by this rank.  ys [nstars, N], init [nstars, W, ndim] are the GLOBAL
    arrays; returns (lo, hi, chain[hi-lo, W, nsteps, ndim], logp[hi-lo, nsteps, W])."""
    lo, hi = star_block(len(ys), rank, world)
    engine.clear_stars()
    for s in range(lo, hi):
        engine.upload_star(s - lo, t, ys[s], yerr)
    if hi == lo:
        return lo, hi, None, None
    engine.sampler_begin(init[lo:hi], seed=seed, nstars=hi - lo, star_ids=np.arange(lo, hi))
    try:
        engine.sampler_run(nsteps, keep=True)
        chain, logp = engine.sampler_read(0, nsteps)
    finally:
        engine.sampler_end()
    return lo, hi, chain, logp


class _DeviceArray:
    """Zero-copy view of engine-owned device memory for torch (``torch.as_tensor(_DeviceArray(...))``)."""

    def __init__(self, ptr, count):
        self.__cuda_array_interface__ = {"shape": (int(count),), "typestr": "<f8", "data": (int(ptr), False),
                                         "version": 3, "strides": None}


def run_walker_sharded(engine, dist, init, nsteps, seed=42, step0=0, a=2.0, timed=False, host_collective=False):
    """One star, the ensemble sharded by walker across the ranks of `dist` (NCCL, one process per GPU).

    The device-resident form of SURVEY.md section 8e(2): every rank keeps the whole ensemble on its GPU;
    per half-step it proposes for all active walkers, evaluates its slice of them, the slices of the
    proposal log-probabilities are all-gathered in place on the device (the only collective, twice per
    step, a few hundred bytes) and every rank applies the identical accept/reject.  The star must already
    be uploaded (slot 0) on every rank.  Returns (chain[W, nsteps, ndim], logp[nsteps, W]) -- the same
    on every rank and, up to the evaluation geometry's rounding, the single-GPU sampler's -- and, with
    timed, the device milliseconds of every step.  host_collective forces the half-step protocol around
    torch.distributed's all-gather (two host synchronisations per half-step)."""
    rank, world = (dist.get_rank(), dist.get_world_size()) if dist is not None else (0, 1)
    init = np.asarray(init, dtype=float)
    W = init.shape[0]
    H = W // 2
    # the whole loop on the engine's stream (the library's own NCCL communicator) when the half-ensemble
    # divides by the number of ranks; otherwise the half-step calls below around torch's collective
    on_stream = (world == 1 or (H % world == 0 and dist.get_backend() == "nccl")) and not host_collective
    if on_stream and world > 1 and getattr(engine, "_comm", None) != (rank, world):
        engine.comm_init(dist)
    engine.sampler_begin(init, seed=seed, step0=step0, a=a, nstars=1)
    try:
        engine.sampler_reserve(nsteps)
        if on_stream:
            ms = engine.sampler_run_sharded(nsteps, keep=True, timed=timed)
            chain, logp = engine.sampler_read(0, nsteps)
            return (chain[0], logp[0], ms) if timed else (chain[0], logp[0])
        staged = world > 1 and dist.get_backend() != "nccl"   # no device collective: the slices travel through the host
        if world > 1:
            import torch
            ptr, count = engine.sampler_lq_device()
            if not staged:
                lq = torch.as_tensor(_DeviceArray(ptr, count), device=torch.device("cuda", torch.cuda.current_device()))
        slices = [walker_slice(H, r, world) for r in range(world)]
        lo, hi = slices[rank]
        even = H % world == 0
        for _ in range(nsteps):
            for half in (0, 1):
                engine.sampler_half_eval(half, lo, hi)          # returns with the engine's stream drained
                if staged:
                    mine = np.empty(hi - lo)
                    if hi > lo:
                        engine.device_copy(mine, ptr + 8 * lo, to_device=False)
                    parts = [None] * world
                    dist.all_gather_object(parts, mine)
                    engine.device_copy(np.ascontiguousarray(np.concatenate(parts)), ptr, to_device=True)
                elif world > 1:
                    if even:
                        dist.all_gather_into_tensor(lq, lq[lo:hi].clone())
                    else:
                        for r, (l, h) in enumerate(slices):
                            if h > l:
                                dist.broadcast(lq[l:h], src=r)
                    torch.cuda.current_stream().synchronize()   # the accept runs on the engine's stream
                engine.sampler_half_accept(half, keep=True)
        chain, logp = engine.sampler_read(0, nsteps)
    finally:
        engine.sampler_end()
    return chain[0], logp[0]


def gather_chains(dist, local, nstars, rank, world):
    """All ranks -> rank 0: list of per-star arrays in global star order (host-side, once per run)."""
    out = [None] * world
    dist.all_gather_object(out, local)
    if rank != 0:
        return None
    chains = [None] * nstars
    for lo, hi, chain, logp in out:
        for k in range(hi - lo):
            chains[lo + k] = (chain[k], logp[k])
    return chains


# ---------------------------------------------------------------------------------------------
def _philox(ctr, key):
    c = [np.uint32(x) for x in ctr]
    k = [np.uint32(x) for x in key]
    m0, m1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
    for _ in range(10):
        p0 = m0 * np.uint64(c[0])
        p1 = m1 * np.uint64(c[2])
        c = [np.uint32(p1 >> np.uint64(32)) ^ c[1] ^ k[0], np.uint32(p1 & np.uint64(0xFFFFFFFF)),
             np.uint32(p0 >> np.uint64(32)) ^ c[3] ^ k[1], np.uint32(p0 & np.uint64(0xFFFFFFFF))]
        k = [np.uint32((int(k[0]) + 0x9E3779B9) & 0xFFFFFFFF), np.uint32((int(k[1]) + 0xBB67AE85) & 0xFFFFFFFF)]
    return [int(x) for x in c]


def rng_words(seed, star, step, walker, purpose):
    """Host copy of the device stream layout (csrc/device_math.cuh rng_words)."""
    ctr = [walker, step & 0xFFFFFFFF, ((step >> 32) & 0x0FFFFFFF) | (purpose << 28), star]
    return _philox(ctr, [seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF])


def u53(lo, hi):
    return float(((hi << 32) | lo) >> 11) / 9007199254740992.0


class HalfEnsembleSampler:
    """emcee's red/blue stretch move with the active half sharded by walker across ranks.

    log_prob_fn(theta[k, ndim]) -> lp[k] evaluates this rank's slice (the CUDA engine in
    production).  `dist` is torch.distributed (or None for a single process).  Every rank holds
    the full ensemble; after each half-step the slices of (proposal log-probabilities) are
    all-gathered and every rank applies the identical accept/reject, so the replicated state
    never diverges."""

    def __init__(self, log_prob_fn, init, seed=42, star=0, step0=0, a=2.0, dist=None):
        self.f = log_prob_fn
        self.x = np.array(init, dtype=float)
        self.W, self.ndim = self.x.shape
        if self.W % 2 or self.W < 2 * self.ndim:
            raise ValueError("need an even number of walkers, at least 2*ndim")
        self.seed, self.star, self.step, self.a, self.dist = seed, star, step0, a, dist
        self.rank = dist.get_rank() if dist is not None else 0
        self.world = dist.get_world_size() if dist is not None else 1
        self.lp = self._evaluate(self.x)
        self.naccept = np.zeros(self.W, dtype=np.int64)

    def _evaluate(self, theta):
        n = theta.shape[0]
        lo, hi = walker_slice(n, self.rank, self.world)
        mine = np.asarray(self.f(theta[lo:hi]), dtype=float) if hi > lo else np.zeros(0)
        if self.dist is None:
            return mine
        import torch
        sizes = [walker_slice(n, r, self.world) for r in range(self.world)]
        width = max(h - l for l, h in sizes)
        # NCCL (NVLink) moves device tensors; gloo (the CPU tests) host tensors
        dev = torch.device("cuda", torch.cuda.current_device()) if self.dist.get_backend() == "nccl" else "cpu"
        buf = torch.zeros(width, dtype=torch.float64, device=dev)
        buf[:hi - lo] = torch.from_numpy(mine).to(dev)
        out = [torch.zeros(width, dtype=torch.float64, device=dev) for _ in range(self.world)]
        self.dist.all_gather(out, buf)
        return np.concatenate([o[:h - l].cpu().numpy() for o, (l, h) in zip(out, sizes)])

    def step_once(self):
        W, H = self.W, self.W // 2
        keys = []
        for w in range(W):
            r = rng_words(self.seed, self.star, self.step, w, 0)
            keys.append((r[1] << 32) | r[0])
        order = sorted(range(W), key=lambda w: (keys[w], w))
        half_of = np.ones(W, dtype=int)
        half_of[order[:H]] = 0
        for half in (0, 1):
            act = [w for w in range(W) if half_of[w] == half]
            oth = [w for w in range(W) if half_of[w] != half]
            q = np.empty((H, self.ndim))
            fac = np.empty(H)
            for i, w in enumerate(act):
                r = rng_words(self.seed, self.star, self.step, w, 1)
                zz = ((self.a - 1.0) * u53(r[0], r[1]) + 1.0)
                zz = zz * zz / self.a
                fac[i] = (self.ndim - 1.0) * np.log(zz)
                j = oth[(rng_words(self.seed, self.star, self.step, w, 2)[0] * H) >> 32]
                q[i] = self.x[j] - (self.x[j] - self.x[w]) * zz
            lq = self._evaluate(q)
            for i, w in enumerate(act):
                r = rng_words(self.seed, self.star, self.step, w, 1)
                ua = u53(r[2], r[3])
                with np.errstate(divide="ignore", invalid="ignore"):
                    if fac[i] + lq[i] - self.lp[w] > np.log(ua):
                        self.x[w] = q[i]
                        self.lp[w] = lq[i]
                        self.naccept[w] += 1
        self.step += 1

    def run(self, nsteps):
        chain = np.empty((self.W, nsteps, self.ndim))
        logp = np.empty((nsteps, self.W))
        for s in range(nsteps):
            self.step_once()
            chain[:, s, :] = self.x
            logp[s] = self.lp
        return chain, logp
