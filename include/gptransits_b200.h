/*
 * gptransits_b200.h -- C ABI of the B200-native likelihood + sampling engine.
 *
 * The reference (Fill4/gptransits) has no FFI: its seams are Python call sites.
 * Each entry point below names the reference interface it replaces
 * (file:line into the reference tree).  INTEGRATION.md shows the ctypes stub a
 * maintainer of the reference would add.
 *
 * Conventions
 *   - every call returns 0 on success or a negative GPT_E_* code; nothing throws
 *     across the ABI; gpt_last_error() gives the message for the last failure.
 *   - all pointers are HOST memory owned by the caller unless the name ends in
 *     _device; the engine owns its device copies.  Calls are synchronous.
 *   - a per-walker numerical failure (prior out of support, non positive
 *     definite covariance, NaN) is NOT an error return: that walker's
 *     log-probability is -inf, exactly as model.py:302-303 and celerite's
 *     log_likelihood do.
 *   - all floating point is IEEE double.  theta is the flat walker vector
 *     [GP component parameters in config order ..., free transit parameters in
 *     (P,t0,Rrat,aR,cosi,e,w,u1,u2) order] (model.py:262,300-301,306,309).
 *   - one engine = one GPU + one model structure + any number of stars.
 *     One host thread per engine at a time.
 */
#ifndef GPTRANSITS_B200_H
#define GPTRANSITS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gpt_engine gpt_engine;

enum {
	GPT_OK = 0,
	GPT_E_INVALID = -1,   /* bad argument */
	GPT_E_CUDA = -2,      /* CUDA runtime failure (message has the cudaError) */
	GPT_E_NOMODEL = -3,   /* gpt_model_set not called */
	GPT_E_NOSTAR = -4,    /* star slot empty */
	GPT_E_UNSUPPORTED = -5
};

/* component.py:66,111,150 -- the three celerite term classes */
enum { GPT_COMP_GRANULATION = 0, GPT_COMP_OSCILLATION_BUMP = 1, GPT_COMP_WHITE_NOISE = 2 };
/* component.py:33-41, transit.py:56-64 -- scipy.stats frozen priors */
enum { GPT_PRIOR_UNIFORM = 0, GPT_PRIOR_NORMAL = 1, GPT_PRIOR_RECIPROCAL = 2 };
/* K and E of the Mandel-Agol flux: batman's Hastings polynomials (parity with
 * the reference) or iterated to machine precision */
enum { GPT_ELLIP_BATMAN = 0, GPT_ELLIP_EXACT = 1 };
/* How the 5th transit parameter reaches batman's `inc` (degrees).  The reference assigns the
 * value named `cosi` verbatim (transit.py:103); GPT_INC_FROM_COSI is the opt-in
 * inc = degrees(arccos(cosi)). */
enum { GPT_INC_VERBATIM = 0, GPT_INC_FROM_COSI = 1 };

/* (kind, a, b) of one free parameter: uniform(a, b-a) | norm(a, b) | reciprocal(a, b) */
typedef struct {
	int32_t kind;
	int32_t _pad;
	double a, b;
} gpt_prior;

/* The model structure: what GPModel.__init__ (gp.py:10-29) and
 * BatmanModel.__init__/init_model (transit.py:14-91) hold after construction. */
typedef struct {
	int32_t ncomp;              /* number of GP components (0 = transit only) */
	const int32_t *comp_type;   /* [ncomp] GPT_COMP_* in config order */
	int32_t has_transit;        /* 0 = GP only (model.py:315-325) */
	double tr_fixed[9];         /* values of fixed transit parameters (transit.py:52-55) */
	uint8_t tr_free[9];         /* BatmanModel.mask (transit.py:19,54) */
	int32_t supersample;        /* batman supersample_factor (model.py:174: 6) */
	double exp_time_days;       /* batman exp_time (transit.py:90) */
	int32_t ellip_mode;         /* GPT_ELLIP_* */
	int32_t inc_mode;           /* GPT_INC_* */
	int32_t ndim;               /* number of free parameters = len(priors) */
	const gpt_prior *priors;    /* [ndim] in theta order */
} gpt_model_desc;

/* ---- lifetime ----------------------------------------------------------- */
int gpt_engine_create(int device, gpt_engine **out);
void gpt_engine_destroy(gpt_engine *e);
const char *gpt_last_error(const gpt_engine *e);
const char *gpt_version(void);

/* Tunables: "halo" (fixed halo length, 0 = adaptive), "chunks" (0 = cost model), "verify", "verify_tol",
 * "halo_auto", "adapt_every" (sampler steps per block), "epoch_scan",
 * "fuse_scan" (epoch scan inside the prep kernel for small batches, default 1), "transit_strict" (batman's
 * operation order), "series_order_min" (0..4: lowest treatment of regular-cadence steps, 4 = general loop),
 * "pdl" (programmatic dependent launch, default 0), "speculate" (sampler blocks without repair launches,
 * replayed when a chunk was flagged: 0 never, 1 after a long clean record -- default --, 2 eagerly), "scratch_cap_mb" (default 8192: an evaluation
 * whose transit-model scratch, 16 bytes per walker and sample, would exceed it runs as sub-batches), "async_egress"
 * (gpt_sample streams finished step ranges of the chain to the host while sampling, default 1), "time_kernels", "reset_timers".
 * Counters: "launches", "evals", "pass2_chunks", "spec_replays", "verify_ratio", "halo", "chunks", "chunk_len",
 * "transit_items", "ms_<kernel>" / "n_<kernel>" for prep, transit, trscan, chunk, final, pass2. */
int gpt_set_option(gpt_engine *e, const char *key, double value);
int gpt_get_stat(gpt_engine *e, const char *key, double *value);

/* ---- model + data ------------------------------------------------------- */
/* Replaces Model.setup_models (model.py:168-201): GPModel(config['gp']),
 * BatmanModel(config['transit'][0]).init_model(time, cadence, 6). */
int gpt_model_set(gpt_engine *e, const gpt_model_desc *desc);

/* Replaces celerite.GP.compute(t*0.0864, yerr=...) (model.py:183-187) plus the
 * light-curve members read by the likelihood (model.py:309-310: lc.time,
 * lc.flux).  t in days, y and yerr in the flux unit the caller fits in (the
 * reference: ppm).  yerr == NULL -> celerite's default 1.123e-12.  A re-upload into the slot last uploaded with
 * bit-identical time stamps (new fluxes or error bars on the same grid) keeps the analysis of the grid. */
int gpt_star_upload(gpt_engine *e, int star, const double *t_days, const double *y,
		    const double *yerr, int64_t N);
/* The same for nstars stars observed on ONE time grid with one error array (a survey batch, slots star0 ..
 * star0+nstars-1): y [nstars][N].  The grid is analysed once and only the fluxes cross the bus per star. */
int gpt_stars_upload(gpt_engine *e, int star0, int nstars, const double *t_days, const double *y,
		     const double *yerr, int64_t N);
int gpt_star_clear(gpt_engine *e);

/* ---- posterior seam ----------------------------------------------------- */
/* Replaces Model.full_lnlike / gp_lnlike / transit_lnlike (model.py:299-343)
 * evaluated for W walkers at once: theta[W*ndim] row-major -> lp[W].
 * parts (optional, [W*4]) = lnprior, quad (r^T K^-1 r), logdet, lnL. */
int gpt_log_prob(gpt_engine *e, int star, const double *theta, int W, double *lp, double *parts);

/* Same for nstars consecutive stars (all of length N): theta[nstars*W*ndim]. */
int gpt_log_prob_multi(gpt_engine *e, int star0, int nstars, const double *theta, int W, double *lp,
		       double *parts);

/* Same with device pointers; asynchronous on the engine's stream. */
int gpt_log_prob_device(gpt_engine *e, int star0, int nstars, const double *theta_device, int W,
			double *lp_device, double *parts_device);

/* Replaces BatmanModel.compute's batman call (transit.py:124-127 ->
 * batman.TransitModel.light_curve): pars9[W*9] = per,t0,rp,a,inc[deg],ecc,w[deg],u1,u2
 * in batman's own semantics; flux[W*N] = supersampled, exposure-averaged
 * relative flux (1.0 out of transit). */
int gpt_transit_flux(gpt_engine *e, int star, const double *pars9, int W, double *flux);

/* ---- GP conditional prediction ---------------------------- */
/* Replaces celerite.GP(kernel).compute(t, yerr); .predict(flux - mean, x, return_var=True) as the
 * reference's light-curve figure calls it (plot.py:250-260,322-331).  theta[ndim] is ONE walker
 * vector (the transit part, if any, is subtracted from the data first, as plot.py:246-258 does);
 * x_days[M] are the prediction times; mu[M] and (optional) var[M] are in the data's flux unit.
 * Errors: GPT_E_INVALID when theta is outside the prior support or the matrix is not positive
 * definite (celerite raises LinAlgError). */
int gpt_gp_predict(gpt_engine *e, int star, const double *theta, const double *x_days, int64_t M, double *mu,
		   double *var);

/* ---- sampler seam ------------------------------------------------------- */
/* Replaces emcee.EnsembleSampler(W, ndim, lnlike).run_mcmc(init, nsteps)
 * (model.py:270-284) with the default StretchMove(a) for nstars independent
 * stars, each with its own ensemble:
 *   init  [nstars][W][ndim]
 *   chain [nstars][W][nsteps][ndim]   (sampler.chain, model.py:283)
 *   logp  [nstars][nsteps][W]         (sampler.get_log_prob(), model.py:284)
 *   naccept [nstars][W] (optional)
 * step0 is the number of iterations already done (resume, model.py:235-253):
 * the Philox stream is keyed by (seed, star id, step0 + i, walker) so results
 * do not depend on how the run is split or on the number of GPUs.  Every step is kept.
 * chain and logp are ordinary host buffers; finished ranges of steps are copied into them on a second
 * stream while the sampler runs (the reference appends every step to its backend, model.py:204-220), so
 * only the last range's copy is exposed. */
int gpt_sample(gpt_engine *e, int star0, int nstars, const double *init, int W, int64_t nsteps,
	       uint64_t seed, uint64_t step0, double a, double *chain, double *logp, int64_t *naccept);

/* Device-resident sampler, for callers that keep the chain on the GPU or shard
 * an ensemble across GPUs.  star_ids (optional, [nstars]) are the GLOBAL star
 * ids used in the RNG key (default star0 + i). */
int gpt_sampler_begin(gpt_engine *e, int star0, int nstars, const int32_t *star_ids, const double *init,
		      int W, uint64_t seed, uint64_t step0, double a);
/* advance nsteps; positions/log-probs stay on the device.  keep != 0 appends
 * every step to the device chain buffer (read with gpt_sampler_read). */
int gpt_sampler_run(gpt_engine *e, int64_t nsteps, int keep);
int gpt_sampler_read(gpt_engine *e, int64_t first, int64_t count, double *chain, double *logp);
int gpt_sampler_state(gpt_engine *e, double *coords, double *logp, int64_t *naccept);
/* Gelman-Rubin statistic V/W (the reference's convention: convergence.py:4-35, no square root) of the
 * kept steps [first, first + count) of the device chain, per star and parameter: rhat[nstars*ndim].
 * Replaces pulling the chain to the host for the check at model.py:417-424 while a run is in flight. */
int gpt_sampler_rhat(gpt_engine *e, int64_t first, int64_t count, double *rhat);
/* Geweke z-scores (convergence.py:75-99, used for the burn-in at model.py:387-402) of the kept steps
 * [first, first + count) of the device chain: for bin b the sub-chain from starts[b] (relative to first; the
 * reference's np.linspace(0, count/2, nbins).astype(int)) is cut into its first frac1 and its part after frac2,
 * z = |mean_head - mean_tail| / sqrt(var_head + var_tail).  z [nstars][nbins][W][ndim]. */
int gpt_sampler_geweke(gpt_engine *e, int64_t first, int64_t count, int nbins, const int64_t *starts, double frac1,
		       double frac2, double *z);
/* One star on several GPUs (one process each): the ensemble is sharded by walker.  Every rank
 * holds the whole ensemble; per half-step: gpt_sampler_half_eval proposes for all active walkers and
 * evaluates this rank's slice [lo, hi) of them into the device array lq[W/2] (gpt_sampler_lq_device), the
 * caller all-gathers lq (ncclAllGather on that pointer -- the only collective of the path, twice per step),
 * gpt_sampler_half_accept applies the same accept/reject on every rank.  gpt_sampler_reserve sizes the
 * device chain first.  Replaces the pool hook of emcee's EnsembleSampler (model.py:269-270). */
int gpt_sampler_reserve(gpt_engine *e, int64_t nsteps);
int gpt_sampler_lq_device(gpt_engine *e, void **lq, int64_t *count);
int gpt_sampler_half_eval(gpt_engine *e, int half, int lo, int hi);
int gpt_sampler_half_accept(gpt_engine *e, int half, int keep);
/* The same with nothing but device work inside a step: the ranks form an NCCL communicator of their own
 * (gpt_comm_id on one rank -> 128-byte id -> the caller hands it to every rank -> gpt_comm_init on all), and
 * gpt_sampler_run_sharded enqueues, per half-step and all on the engine's stream, the proposals, this rank's slice
 * of the posteriors, ncclAllGather of the proposal log-probabilities in place over NVLink, and the accept/reject.
 * The half-ensemble must divide by the number of ranks.  ms_per_step (may be null): device time of every step. */
int gpt_comm_id(void *id128);
int gpt_comm_init(gpt_engine *e, const void *id128, int rank, int world);
int gpt_comm_destroy(gpt_engine *e);
int gpt_sampler_run_sharded(gpt_engine *e, int64_t nsteps, int keep, float *ms_per_step);
int gpt_sampler_end(gpt_engine *e);

/* ---- measurement helpers (bench.py) ------------------------------------------------ */
/* gpt_sampler_run with one CUDA-event pair per step on the engine's stream; flush_l2 != 0
 * overwrites a 256 MiB buffer between steps (outside the timed intervals). */
int gpt_sampler_run_timed(gpt_engine *e, int64_t nsteps, int keep, int flush_l2, float *ms_per_step);
/* reps evaluations of a device-resident batch, one event pair each. */
int gpt_log_prob_device_timed(gpt_engine *e, int star0, int nstars, const double *theta_device, int W,
			      double *lp_device, int reps, int flush_l2, float *ms_per_rep);
/* plain device buffers for callers without their own CUDA allocator */
int gpt_device_alloc(gpt_engine *e, size_t bytes, void **out);
int gpt_device_free(gpt_engine *e, void *p);
int gpt_device_copy(gpt_engine *e, void *dst, const void *src, size_t bytes, int to_device);

/* FP64 FMA throughput of this GPU (TFLOP/s, FMA = 2), measured by a register
 * resident DFMA kernel -- the roofline denominator bench.py reports against. */
int gpt_measure_fp64_peak(gpt_engine *e, double *tflops);

#ifdef __cplusplus
}
#endif
#endif
