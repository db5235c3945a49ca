// engine.cu -- host side of the C ABI declared in include/gptransits_b200.h.
// Owns device copies of the light curves, the workspaces of one evaluation batch, and the
// device-resident sampler state.  No torch types, no CPU fallback: every result comes from the
// kernels in kernels.cuh.
#include "../../include/gptransits_b200.h"
#include "kernels.cuh"

#include <dlfcn.h>
#include <nccl.h>  // types only: the library is opened at run time (gpt_comm_*), never linked

#include <algorithm>
#include <cmath>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <thread>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

using namespace gpt;

#define CK(call)                                                                                   \
	do {                                                                                       \
		cudaError_t _e = (call);                                                           \
		if (_e != cudaSuccess) {                                                           \
			e->set_error(std::string(#call) + ": " + cudaGetErrorString(_e));          \
			return GPT_E_CUDA;                                                         \
		}                                                                                  \
	} while (0)

enum { TK_PREP = 0, TK_TRANSIT, TK_CHUNK, TK_FINAL, TK_PASS2, TK_TRSCAN, TK_COUNT };

namespace {

template <typename T>
struct DevBuf {
	T *p = nullptr;
	size_t cap = 0;
	cudaError_t ensure(size_t n)
	{
		if (n <= cap) return cudaSuccess;
		if (p) cudaFree(p);
		p = nullptr;
		cap = 0;
		size_t want = n + n / 8;
		cudaError_t err = cudaMalloc((void **)&p, want * sizeof(T));
		if (err == cudaSuccess) cap = want;
		return err;
	}
	void release()
	{
		if (p) cudaFree(p);
		p = nullptr;
		cap = 0;
	}
};

struct Star {
	char *blob = nullptr;     // one allocation per star (samp | tdays | tile_dt0 | irr), reused by re-uploads that fit
	size_t blob_cap = 0, blob_used = 0;
	double4 *samp = nullptr;
	double *tdays = nullptr;
	int *irr = nullptr;
	int nirr = 0;
	double *tile_dt0 = nullptr;
	int N = 0;
	double dt0 = 0, dreg = 0, dmax = 0, cdev = 0;
	bool has_err = false;
};

}  // namespace

struct gpt_engine {
	int device = 0;
	int num_sms = 148;       // cudaDevAttrMultiProcessorCount of the device (B200: 148)
	double sm_clock_ghz = 0.0;  // cudaDevAttrClockRate
	cudaStream_t stream = nullptr;
	std::string err;
	bool have_model = false;
	ModelDev model;
	bool may_generic = false;
	std::vector<Star> stars;
	DevBuf<StarDev> d_stars;
	bool stars_dirty = true;

	// options
	int opt_halo = 0;        // 0 = auto
	int opt_chunks = 0;      // 0 = auto
	int opt_verify = 1;
	int opt_halo_auto = 1;
	int opt_epoch_scan = 1;
	int opt_adapt_every = 32;  // sampler steps between halo adaptations
	int opt_transit_strict = 0;
	int opt_series_order_min = 0;
	int opt_speculate = 1;   // speculative sampler blocks: 0 never, 1 after a long clean record, 2 whenever the last block was clean
	long long spec_clean_steps = 0;   // sampler steps since a chunk was last flagged (or the halo last changed)
	long long spec_need_steps = 128;  // clean record needed before a block runs speculatively; doubles on a replay
	bool spec_last_clean = true;
	long long n_spec_replays = 0;
	int opt_scratch_cap_mb = 8192;  // most scratch (transit model + queue) one evaluation may hold; larger batches are sliced
	int opt_async_egress = 1;  // gpt_sample: the chain leaves for the host range by range while the sampler runs
	long long n_egress_jobs = 0;
	double opt_tol = 5e-10;   // budget (a bound, not an estimate) for the chunk-coupling error, relative to quad and to logdet
	int cur_halo = 0;         // 0: derive from the priors at the first evaluation
	double prior_cmin = 0.0;  // smallest celerite decay rate c the priors allow (rad per megasecond)

	// stats
	long long n_launch = 0, n_pass2_calls = 0, n_pass2_chunks = 0, n_evals = 0;
	double last_verr = 0.0;
	int last_chunks = 0, last_L = 0, last_H = 0;

	// evaluation workspaces
	DevBuf<double> d_theta, d_lp, d_parts, d_prep, d_mean, d_carryT, d_carryE, d_yup;
	DevBuf<double> d_zero;  // [1] 0.0: what the chunk kernel reads for a sample outside the transit window
	DevBuf<unsigned int> d_winmask;  // [nwalk][ceil(N/32)] in-window bits of the transit model
	DevBuf<int> d_flags, d_wstatus, d_wbad;
	DevBuf<ChunkPartial> d_partial;
	DevBuf<ChunkPartial> d_partial2;  // [nwalk] results of the sequential recomputation of flagged walkers
	DevBuf<unsigned char> d_redo;
	DevBuf<unsigned long long> d_queue;  // transit work queue: (walker << 32) | sample
	DevBuf<unsigned int> d_qcount;
	DevBuf<unsigned int> d_qfused;  // [0] item counter of the fused prep + scan path (zero between evaluations), [1] last count
	bool last_fused = false;
	int opt_fuse_scan = 1;
	DevBuf<int> d_nbad;      // [1]
	DevBuf<double> d_verr;   // [1]
	char *h_stage = nullptr;  // pinned staging buffer of gpt_star_upload (grow-only)
	size_t h_stage_cap = 0;
	int stage_star = -1;      // slot whose blob the staging buffer holds (-1: none)
	int64_t stage_N = 0;
	long long n_grid_reuse = 0;
	int *h_nbad = nullptr;   // pinned
	double *h_verr = nullptr;

	// NCCL communicator of the walker-sharded sampler (gpt_comm_init)
	ncclComm_t comm = nullptr;
	int comm_rank = 0, comm_world = 1;

	// sampler
	bool sampling = false;
	SamplerDev S;
	int s_star0 = 0;
	unsigned long long s_step = 0;
	long long s_kept = 0;
	DevBuf<double> s_x, s_lpx, s_q, s_lq, s_fac, s_chain, s_logp;
	DevBuf<int> s_act, s_ids;
	DevBuf<long long> s_nacc;
	DevBuf<double> s_snap_x, s_snap_lpx;  // ensemble at the start of a speculative block
	DevBuf<long long> s_snap_nacc;

	// running totals written by the finalize kernel: [0] flagged chunks, [1] max error ratio (bits)
	DevBuf<unsigned long long> d_acc;

	// optional per-kernel CUDA-event timing on the engine stream (option "time_kernels")
	int opt_time = 0;
	std::vector<std::pair<cudaEvent_t, cudaEvent_t>> tev[TK_COUNT];
	std::vector<cudaEvent_t> ev_pool;
	double t_ms[TK_COUNT] = {0};
	long long t_n[TK_COUNT] = {0};
	cudaEvent_t ev_get()
	{
		if (!ev_pool.empty()) {
			cudaEvent_t v = ev_pool.back();
			ev_pool.pop_back();
			return v;
		}
		cudaEvent_t v;
		cudaEventCreate(&v);
		return v;
	}
	void timer_begin(int k)
	{
		if (!opt_time) return;
		cudaEvent_t a = ev_get(), b = ev_get();
		cudaEventRecord(a, stream);
		tev[k].push_back({a, b});
	}
	void timer_end(int k)
	{
		if (!opt_time || tev[k].empty()) return;
		cudaEventRecord(tev[k].back().second, stream);
	}
	void timer_collect()
	{
		cudaStreamSynchronize(stream);
		for (int k = 0; k < TK_COUNT; k++) {
			for (auto &pr : tev[k]) {
				float ms = 0;
				if (cudaEventElapsedTime(&ms, pr.first, pr.second) == cudaSuccess) {
					t_ms[k] += ms;
					t_n[k]++;
				}
				ev_pool.push_back(pr.first);
				ev_pool.push_back(pr.second);
			}
			tev[k].clear();
		}
	}
	void adapt_halo(int H, bool failed, double ratio)
	{
		if (!opt_halo_auto || opt_halo != 0 || H <= 0) return;
		// geometric decay model: error/tolerance ratio r at halo H, O(1e12) at H = 0; aim at 0.3
		double r = failed ? 1e3 : std::max(ratio, 1e-30);
		double lr0 = std::log(1e12), lr = std::log(r), lt = std::log(0.3);
		double want = H * (lr0 - lt) / std::max(lr0 - lr, 1.0);
		int nh = (int)std::ceil(want / 16.0) * 16;
		nh = std::max(32, std::min(nh, 32768));
		// hysteresis: a halo at which a chunk was flagged is not tried again for a while (the model above
		// would otherwise shrink back to it and fail again, and a flagged block costs a repair or a replay)
		if (failed) {
			fail_halo = std::max(fail_halo, H);
			fail_age = 0;
		} else if (++fail_age > 64) {
			fail_halo = 0;
		}
		if (fail_halo > 0) nh = std::max(nh, fail_halo + 16);
		cur_halo = nh > cur_halo ? nh : std::max(nh, (int)(cur_halo * 0.75));
	}
	int fail_halo = 0, fail_age = 0;

	void set_error(const std::string &m) { err = m; }
};

// ---------------------------------------------------------------------------------------------
static int sync_stars(gpt_engine *e)
{
	if (!e->stars_dirty) return GPT_OK;
	std::vector<StarDev> h(e->stars.size());
	for (size_t i = 0; i < h.size(); i++) {
		h[i].samp = e->stars[i].samp;
		h[i].tdays = e->stars[i].tdays;
		h[i].tile_dt0 = e->stars[i].tile_dt0;
		h[i].irr = e->stars[i].irr;
		h[i].nirr = e->stars[i].nirr;
		h[i]._pad2 = 0;
		h[i].N = e->stars[i].N;
		h[i]._pad = 0;
		h[i].dt0 = e->stars[i].dt0;
		h[i].dreg = e->stars[i].dreg;
		h[i].dmax = e->stars[i].dmax;
		h[i].cdev = e->stars[i].cdev;
	}
	CK(e->d_stars.ensure(std::max<size_t>(1, h.size())));
	if (!h.empty())
		CK(cudaMemcpyAsync(e->d_stars.p, h.data(), h.size() * sizeof(StarDev), cudaMemcpyHostToDevice, e->stream));
	CK(cudaStreamSynchronize(e->stream));
	e->stars_dirty = false;
	return GPT_OK;
}

static int check_stars(gpt_engine *e, int star0, int nstars)
{
	if (!e->have_model) {
		e->set_error("no model: call gpt_model_set first");
		return GPT_E_NOMODEL;
	}
	if (star0 < 0 || nstars < 1 || (size_t)(star0 + nstars) > e->stars.size()) {
		e->set_error("star index out of range");
		return GPT_E_NOSTAR;
	}
	int N = e->stars[star0].N;
	for (int s = star0; s < star0 + nstars; s++) {
		if (!e->stars[s].samp) {
			e->set_error("star slot empty");
			return GPT_E_NOSTAR;
		}
		if (e->stars[s].N != N) {
			e->set_error("all stars of one batch must have the same number of samples");
			return GPT_E_INVALID;
		}
	}
	return GPT_OK;
}

// chunk geometry: halo from the priors / the adaptation, number of chunks from the measured cost model
static void pick_geometry(gpt_engine *e, int star0, int N, int W, int nstars, int &nchunks, int &L, int &H)
{
	if (e->opt_halo > 0) {
		H = e->opt_halo;
	} else {
		if (e->cur_halo <= 0) {
			// first evaluation: 10 decades of decay of the slowest mode the priors allow, exp(-c_min dt0)
			// per step (measured on C3: the criterion passes after ~9.5 decades); adapted afterwards
			double dt0 = std::fabs(e->stars[star0].dt0);
			double h = (e->prior_cmin > 0 && dt0 > 0) ? std::log(1e10) / (e->prior_cmin * dt0) : 192.0;
			e->cur_halo = (int)std::min(8192.0, std::max(64.0, std::ceil(h / 16.0) * 16.0));
		}
		H = e->cur_halo;
	}
	int nc;
	if (e->opt_chunks > 0) {
		nc = e->opt_chunks;
	} else {
		// Cost model (measured on B200): every thread runs N/nc + H dependent steps; up to one warp
		// per SM sub-partition they all run concurrently at the single-warp step time, more resident
		// warps share the sub-partition's issue slots and FP64 pipe with little gain.
		const int cthreads = std::min(CHUNK_THREADS, ((W + 31) / 32) * 32);
		const double wpl = (double)nstars * ((W + cthreads - 1) / cthreads) * (cthreads / 32);  // warps per chunk layer
		const double slots = (double)e->num_sms * 4.0;
		const int nc_max = std::max(1, N / 32);
		double best = 1e300;
		nc = 1;
		for (int cand = 1; cand <= nc_max; cand = cand < 16 ? cand + 1 : (int)std::ceil(cand * 1.08)) {
			const double h = cand == 1 ? 0.0 : (double)H;
			const double wps = cand * wpl / slots;
			double slow = 1.0;
			// (measured on C3 and on the C4 batch at 1/2/4/8 GPUs: a step function -- the kernel lasts as long as the
			// SM with the most CTAs, and k warps on a sub-partition take 0.885 k times one warp's time per step)
			if (wps > 1.0) slow = std::ceil(wps - 1e-9) * 0.885;
			const double cost = ((double)N / cand + h) * slow + 0.05 * cand;
			if (cost < best * 0.999) {
				best = cost;
				nc = cand;
			}
		}
		// snap to whole waves when close to one
		const double per_wave = slots / wpl;
		if (per_wave >= 1.0 && nc > 0.8 * per_wave && nc < 1.25 * per_wave) nc = (int)per_wave;
	}
	nc = std::max(1, std::min(nc, N));
	L = (N + nc - 1) / nc;
	nchunks = (N + L - 1) / L;
	if (nchunks == 1) H = 0;
}

// Launch with programmatic stream serialisation (see pdl_wait in kernels.cuh): the kernel may become resident
// while its predecessor in the stream is still running and blocks in griddepcontrol.wait until that one
// has completed.  Off by default (option "pdl"): measured on the C3 sampler step it gains nothing (0.669 ms
// per step with, 0.652 ms without) -- the stream's back-to-back launch gaps are already short and the
// early-resident CTAs cost more than the hidden prologues save.
static int g_pdl = 0;
template <typename... KArgs, typename... Args>
static void launch_kx(bool pdl, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args &&...args)
{
	cudaLaunchConfig_t cfg;
	memset(&cfg, 0, sizeof(cfg));
	cfg.gridDim = grid;
	cfg.blockDim = block;
	cfg.dynamicSmemBytes = smem;
	cfg.stream = st;
	cudaLaunchAttribute attr[1];
	attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
	attr[0].val.programmaticStreamSerializationAllowed = 1;
	cfg.attrs = attr;
	cfg.numAttrs = (g_pdl && pdl) ? 1 : 0;
	cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}
template <typename... KArgs, typename... Args>
static void launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args &&...args)
{
	launch_kx(true, kernel, grid, block, smem, st, std::forward<Args>(args)...);
}

// dynamic shared memory of the epoch scan (stand-alone or fused into the prep kernel) for light curves of N
// samples: the coarse time table and room for the walker's items -- an eighth of the light curve covers any
// realistic transit duty cycle (a walker that needs more takes the direct-append path)
static int scan_items_cap(int N) { return std::max(256, std::min(EPOCH_SCAN_CAP, N / 8)); }
static size_t scan_smem_bytes(int N)
{
	int cstride = SCAN_STRIDE;
	while ((N + cstride - 1) / cstride > SCAN_COARSE) cstride <<= 1;
	const int ncoarse = (N + cstride - 1) / cstride;
	return sizeof(double) * (size_t)ncoarse + sizeof(unsigned int) * (size_t)scan_items_cap(N);
}

// flux of the queued (walker, sample) items; "transit_strict" selects batman's own operation order
static void launch_transit_eval(gpt_engine *e, cudaStream_t st, const StarDev *stars, int W, int ss, double exp_time,
				int mode, const double *prep, const unsigned long long *queue, const unsigned int *qcount,
				double *out, int dense, unsigned int *winmask = nullptr, int nw32 = 0)
{
	const int grid = e->num_sms * 2 * TR_MINBLOCKS;
	if (e->opt_transit_strict)
		launch_k(transit_eval_kernel<true>, dim3(grid), dim3(TR_THREADS), 0, st, stars, W, ss, exp_time, mode, prep, queue, qcount, out, dense, winmask, nw32);
	else
		launch_k(transit_eval_kernel<false>, dim3(grid), dim3(TR_THREADS), 0, st, stars, W, ss, exp_time, mode, prep, queue, qcount, out, dense, winmask, nw32);
}

// threads per chunk CTA: one per walker of the star, a whole number of warps, at most 128
static int chunk_threads(int W) { return std::min(CHUNK_THREADS, ((W + 31) / 32) * 32); }

static void launch_finalize(int nt, int nwalk, cudaStream_t st, const FinalArgs &f)
{
	if (f.nchunks == 1 && !f.skip_if_clean) {
		launch_k(gp_finalize_wide_kernel, dim3((nwalk + FINW_THREADS / 32 - 1) / (FINW_THREADS / 32)), dim3(FINW_THREADS), 0, st, f);
		return;
	}
	// thousands of walkers cut into a few chunks: one warp per walker; otherwise one block per walker
	if (nwalk >= 1024 && f.nchunks <= 64) {
		const dim3 g((nwalk + FINW_WALKERS - 1) / FINW_WALKERS), b(32 * FINW_WALKERS);
		switch (nt) {
		case 1: launch_k(gp_finalize_kernel<1, 32>, g, b, 0, st, f); break;
		case 2: launch_k(gp_finalize_kernel<2, 32>, g, b, 0, st, f); break;
		case 3: launch_k(gp_finalize_kernel<3, 32>, g, b, 0, st, f); break;
		case 4: launch_k(gp_finalize_kernel<4, 32>, g, b, 0, st, f); break;
		default: break;
		}
		return;
	}
	switch (nt) {
	case 1: launch_k(gp_finalize_kernel<1, FIN_THREADS>, dim3(nwalk), dim3(FIN_THREADS), 0, st, f); break;
	case 2: launch_k(gp_finalize_kernel<2, FIN_THREADS>, dim3(nwalk), dim3(FIN_THREADS), 0, st, f); break;
	case 3: launch_k(gp_finalize_kernel<3, FIN_THREADS>, dim3(nwalk), dim3(FIN_THREADS), 0, st, f); break;
	case 4: launch_k(gp_finalize_kernel<4, FIN_THREADS>, dim3(nwalk), dim3(FIN_THREADS), 0, st, f); break;
	default: break;
	}
}

template <int NT>
static void launch_chunk_nt(bool transit, dim3 grid, int threads, cudaStream_t st, const ChunkArgs &a)
{
	// Pass 1 is launched stream-ordered: CTAs of a programmatic launch are placed as the preceding kernel
	// frees resources, which can put two chunk CTAs on one SM and none on another -- with a grid of one
	// CTA per SM that doubles the kernel's duration.  Pass 2 (normally an immediate exit) overlaps.
	const bool pdl = a.pass2 != 0;
	if (transit)
		launch_kx(pdl, gp_chunk_kernel<NT, true>, dim3(grid), dim3(threads), 0, st, a);
	else
		launch_kx(pdl, gp_chunk_kernel<NT, false>, dim3(grid), dim3(threads), 0, st, a);
}

static void launch_chunk(int nt, bool transit, dim3 grid, int threads, cudaStream_t st, const ChunkArgs &a)
{
	switch (nt) {
	case 1: launch_chunk_nt<1>(transit, grid, threads, st, a); break;
	case 2: launch_chunk_nt<2>(transit, grid, threads, st, a); break;
	case 3: launch_chunk_nt<3>(transit, grid, threads, st, a); break;
	case 4: launch_chunk_nt<4>(transit, grid, threads, st, a); break;
	default: break;
	}
}

// Enqueue one posterior evaluation of nstars*W walkers.  theta/lp/parts are device pointers.
// host_sync: read the verification verdict back and launch pass 2 only when needed; otherwise pass 2
// is always enqueued (it exits immediately when nothing was flagged) so the sequence is graph-safe.
static int eval_enqueue_batch(gpt_engine *e, int star0, int nstars, const double *d_theta, int W, double *d_lp,
			      double *d_parts, bool host_sync, const PrepPropose *propose, const FinalAccept *accept,
			      bool *accept_done, bool speculative);

// One posterior evaluation of nstars x W walkers.  The transit model and its work queue are scratch of
// 16 bytes per (walker, sample); a plain evaluation (no fused proposal / accept) whose scratch would exceed
// "scratch_cap_mb" is run as a sequence of sub-batches -- whole stars, or slices of the walkers of one star --
// that reuse the same scratch, so the footprint is bounded by the cap, not by nwalk x N.
static int eval_enqueue(gpt_engine *e, int star0, int nstars, const double *d_theta, int W, double *d_lp,
			double *d_parts, bool host_sync, const PrepPropose *propose = nullptr,
			const FinalAccept *accept = nullptr, bool *accept_done = nullptr, bool speculative = false)
{
	if (!propose && !accept && e->have_model && e->model.has_transit && star0 >= 0 && nstars >= 1 &&
	    (size_t)(star0 + nstars) <= e->stars.size() && e->stars[star0].N > 0) {
		const double per_walker = 16.0 * (double)e->stars[star0].N;
		const double cap = (double)e->opt_scratch_cap_mb * 1048576.0;
		if (per_walker * (double)nstars * (double)W > cap) {
			const int ndim = e->model.ndim;
			const long long fit = std::max<long long>(32, (long long)(cap / per_walker));  // walkers per sub-batch
			if (fit >= W) {
				const int sper = (int)std::max<long long>(1, fit / W);  // whole stars
				for (int s0 = 0; s0 < nstars; s0 += sper) {
					const int sn = std::min(sper, nstars - s0);
					int rc = eval_enqueue_batch(e, star0 + s0, sn, d_theta + (size_t)s0 * W * ndim, W, d_lp + (size_t)s0 * W,
								    d_parts ? d_parts + (size_t)s0 * W * 4 : nullptr, host_sync, nullptr, nullptr,
								    nullptr, speculative);
					if (rc) return rc;
				}
			} else {
				for (int s = 0; s < nstars; s++)
					for (int w0 = 0; w0 < W; w0 += (int)fit) {
						const int wn = (int)std::min<long long>(fit, W - w0);
						const size_t off = (size_t)s * W + w0;
						int rc = eval_enqueue_batch(e, star0 + s, 1, d_theta + off * ndim, wn, d_lp + off,
									    d_parts ? d_parts + off * 4 : nullptr, host_sync, nullptr, nullptr, nullptr,
									    speculative);
						if (rc) return rc;
					}
			}
			return GPT_OK;
		}
	}
	return eval_enqueue_batch(e, star0, nstars, d_theta, W, d_lp, d_parts, host_sync, propose, accept, accept_done, speculative);
}

static int eval_enqueue_batch(gpt_engine *e, int star0, int nstars, const double *d_theta, int W, double *d_lp,
			      double *d_parts, bool host_sync, const PrepPropose *propose, const FinalAccept *accept,
			      bool *accept_done, bool speculative)
{
	if (accept_done) *accept_done = false;
	int rc = check_stars(e, star0, nstars);
	if (rc) return rc;
	rc = sync_stars(e);
	if (rc) return rc;
	const ModelDev &m = e->model;
	const int N = e->stars[star0].N;
	const int nwalk = nstars * W;
	const int nt = m.nterms;
	const int nw32 = 4 * ((N + 127) / 128);  // in-window mask: four 32-bit words per 128-sample tile (rows 16-byte aligned)
	const StarDev *dst = e->d_stars.p + star0;
	cudaStream_t st = e->stream;

	CK(e->d_prep.ensure((size_t)nwalk * PREP_STRIDE));
	CK(e->d_flags.ensure(nwalk));
	CK(e->d_wstatus.ensure(nwalk));
	CK(e->d_nbad.ensure(1));
	CK(e->d_verr.ensure(1));
	if (!e->d_acc.p) {
		CK(e->d_acc.ensure(2));
		CK(cudaMemsetAsync(e->d_acc.p, 0, 2 * sizeof(unsigned long long), st));
	}
	if (m.has_transit) {
		CK(e->d_mean.ensure((size_t)nwalk * N));
		CK(e->d_queue.ensure((size_t)nwalk * N));
		CK(e->d_qcount.ensure(1));
		if (nt > 0) CK(e->d_winmask.ensure((size_t)nwalk * nw32));
		if (!e->d_zero.p) {
			CK(e->d_zero.ensure(1));
			CK(cudaMemsetAsync(e->d_zero.p, 0, sizeof(double), st));
		}
	}
	unsigned int *winmask = (m.has_transit && nt > 0) ? e->d_winmask.p : nullptr;

	// small batches: the epoch scan rides on the prep kernel (one launch, the window never leaves the CTA)
	const bool fuse_scan = m.has_transit && nwalk <= 1024 && e->opt_epoch_scan && e->opt_fuse_scan;
	int prep_threads = PREP_THREADS;
	if (fuse_scan) {
		// one thread per transit epoch: a few per 256 samples at most for any realistic period
		int et = 32;
		while (et < PREP_SCAN_THREADS && et * GPT_SCAN_SAMPLES_PER_THREAD < N) et *= 2;
		prep_threads = std::max(PREP_THREADS, et);
	}
	if (fuse_scan && !e->d_qfused.p) {
		CK(e->d_qfused.ensure(2));
		CK(cudaMemsetAsync(e->d_qfused.p, 0, 2 * sizeof(unsigned int), st));
	}
	e->last_fused = fuse_scan;
	unsigned int *qcount = fuse_scan ? e->d_qfused.p : e->d_qcount.p;
	e->timer_begin(TK_PREP);
	PrepReset reset;
	reset.wstatus = e->d_wstatus.p; reset.nbad = e->d_nbad.p; reset.verr = e->d_verr.p;
	reset.qcount = (m.has_transit && !fuse_scan) ? e->d_qcount.p : nullptr;
	PrepPropose pr;
	memset(&pr, 0, sizeof(pr));
	if (propose) pr = *propose;
	PrepScan sc;
	sc.enabled = fuse_scan ? 1 : 0;
	sc.items_cap = scan_items_cap(N);
	sc.queue = e->d_queue.p;
	sc.qcount = qcount;
	sc.winmask = fuse_scan ? winmask : nullptr;  // cleared by whichever kernel runs the scan
	sc.nw32 = nw32;
	if (nwalk <= 4096 && fuse_scan)
		launch_k(prep_kernel<true>, dim3(nwalk), dim3(prep_threads), scan_smem_bytes(N), st, m, dst, W, d_theta, nwalk,
			 e->d_prep.p, e->d_flags.p, reset, pr, sc);
	else if (nwalk <= 4096)
		launch_k(prep_kernel<false>, dim3(nwalk), dim3(PREP_THREADS), 0, st, m, dst, W, d_theta, nwalk,
			 e->d_prep.p, e->d_flags.p, reset, pr, sc);
	else
		launch_k(prep_wide_kernel, dim3((nwalk + 63) / 64), dim3(64), 0, st, m, dst, W, d_theta, nwalk, e->d_prep.p, e->d_flags.p, reset, pr);
	e->timer_end(TK_PREP);
	e->n_launch++;
	if (m.has_transit) {
		dim3 g(nwalk, (N + TR_THREADS - 1) / TR_THREADS);
		e->timer_begin(TK_TRANSIT);
		if (!fuse_scan) {
			e->timer_begin(TK_TRSCAN);
			if (e->opt_epoch_scan) {
				// one thread per transit epoch: a few per 256 samples at most for any realistic period
				int et = 32;
				while (et < TR_THREADS && et * 256 < N) et *= 2;
				launch_k(transit_epoch_scan_kernel, dim3(nwalk), dim3(et), scan_smem_bytes(N), st, dst, W, e->d_prep.p, e->d_flags.p,
					 e->d_queue.p, qcount, scan_items_cap(N), winmask, nw32);
			} else {
				launch_k(transit_scan_kernel, dim3(g), dim3(TR_THREADS), 0, st, dst, W, e->d_prep.p, e->d_flags.p, e->d_queue.p, qcount,
					 e->d_mean.p, 0, winmask, nw32);
			}
			e->timer_end(TK_TRSCAN);
			e->n_launch++;
		}
		launch_transit_eval(e, st, dst, W, m.supersample, m.exp_time, m.ellip_mode, e->d_prep.p, e->d_queue.p, qcount,
				    e->d_mean.p, 0, winmask, nw32);
		e->timer_end(TK_TRANSIT);
		e->n_launch++;
	}
	e->n_evals += nwalk;

	if (nt == 0) {
		DiagArgs a;
		a.stars = dst; a.prep = e->d_prep.p; a.flags = e->d_flags.p; a.mean = e->d_mean.p;
		a.lp = d_lp; a.parts = d_parts; a.W = W; a.nwalk = nwalk; a.N = N;
		a.gp = m.ncomp > 0; a.transit = m.has_transit; a.has_err = e->stars[star0].has_err;
		a.qreset = fuse_scan ? e->d_qfused.p : nullptr;
		launch_k(diag_kernel, dim3((nwalk * 32 + 255) / 256), dim3(256), 0, st, a);
		e->n_launch++;
		CK(cudaGetLastError());
		return GPT_OK;
	}

	int nchunks, L, H;
	pick_geometry(e, star0, N, W, nstars, nchunks, L, H);
	e->last_chunks = nchunks;
	e->last_L = L;
	e->last_H = H;
	const int J = 2 * nt, CS = J * (J + 1) / 2 + J;
	CK(e->d_carryT.ensure((size_t)nwalk * nchunks * CS));
	CK(e->d_carryE.ensure((size_t)nwalk * nchunks * CS));
	CK(e->d_partial.ensure((size_t)nwalk * nchunks));
	CK(e->d_redo.ensure((size_t)nwalk * nchunks));
	CK(e->d_partial2.ensure((size_t)nwalk));
	CK(e->d_wbad.ensure(nwalk));

	ChunkArgs a;
	a.stars = dst; a.prep = e->d_prep.p; a.flags = e->d_flags.p; a.mean = e->d_mean.p;
	a.winmask = winmask; a.nw32 = nw32; a.zero = e->d_zero.p;
	a.carryT = e->d_carryT.p; a.carryE = e->d_carryE.p; a.partial = e->d_partial.p;
	a.wstatus = e->d_wstatus.p; a.redo = e->d_redo.p; a.nbad = e->d_nbad.p; a.wbad = e->d_wbad.p; a.partial2 = e->d_partial2.p;
	a.W = W; a.N = N; a.nchunks = nchunks; a.L = L; a.H = H; a.pass2 = 0;
	a.min_order = e->opt_series_order_min;
	const int cthreads = chunk_threads(W);
	dim3 grid(nchunks, (W + cthreads - 1) / cthreads, nstars);
	e->timer_begin(TK_CHUNK);
	launch_chunk(nt, m.has_transit != 0, grid, cthreads, st, a);
	e->timer_end(TK_CHUNK);
	e->n_launch++;

	FinalArgs f;
	f.prep = e->d_prep.p; f.flags = e->d_flags.p; f.wstatus = e->d_wstatus.p;
	f.carryT = e->d_carryT.p; f.carryE = e->d_carryE.p; f.partial = e->d_partial.p; f.redo = e->d_redo.p;
	f.nbad = e->d_nbad.p; f.lp = d_lp; f.parts = d_parts; f.verr = e->d_verr.p;
	f.nwalk = nwalk; f.N = N; f.nchunks = nchunks; f.H = H; f.nterms = nt;
	f.verify = (nchunks > 1 && e->opt_verify) ? 1 : 0;
	f.skip_if_clean = 0;
	f.acc = e->d_acc.p;
	f.wbad = e->d_wbad.p;
	f.partial2 = e->d_partial2.p;
	f.L = L;
	f.tol = e->opt_tol;
	f.qreset = fuse_scan ? e->d_qfused.p : nullptr;
	memset(&f.AC, 0, sizeof(f.AC));
	// the accept/reject can ride on the finalize kernel when every walker's result comes from it
	if (accept && !e->may_generic && !host_sync) {
		f.AC = *accept;
		if (accept_done) *accept_done = true;
	}
	e->timer_begin(TK_FINAL);
	launch_finalize(nt, nwalk, st, f);
	e->timer_end(TK_FINAL);
	e->n_launch++;

	if (e->may_generic) {
		GenericArgs g;
		g.stars = dst; g.prep = e->d_prep.p; g.flags = e->d_flags.p; g.mean = e->d_mean.p;
		g.lp = d_lp; g.parts = d_parts; g.W = W; g.nwalk = nwalk; g.N = N; g.nterms = nt;
		g.transit = m.has_transit;
		launch_k(gp_generic_kernel, dim3((nwalk + 31) / 32), dim3(32), 0, st, g);
		e->n_launch++;
	}

	// speculative: the caller (the sampler) checks the accumulated verdicts once per block of steps and
	// replays the block with pass 2 when a chunk was flagged, so nothing is enqueued here
	if (f.verify && !speculative) {
		bool need = true;
		if (host_sync) {
			CK(cudaMemcpyAsync(e->h_nbad, e->d_nbad.p, sizeof(int), cudaMemcpyDeviceToHost, st));
			CK(cudaMemcpyAsync(e->h_verr, e->d_verr.p, sizeof(double), cudaMemcpyDeviceToHost, st));
			CK(cudaStreamSynchronize(st));
			need = *e->h_nbad > 0;
			e->last_verr = *e->h_verr;
			if (need) {
				e->n_pass2_calls++;
				e->n_pass2_chunks += *e->h_nbad;
			}
			e->adapt_halo(H, need, e->last_verr);
		}
		if (need) {
			// Pass 2 exits at once on the device when nothing was flagged, so the asynchronous path can enqueue
			// it unconditionally (no host round trip inside a sampler step).  A flagged chunk is redone exactly
			// from the carry its predecessor left in pass 1; walkers for which that carry cannot be trusted
			// (the predecessor flagged as well, by more than a chunk's worth of decay makes up for) are recomputed whole,
			// sequentially, by the CTAs of chunk 0 of the same launch (gp_finalize_kernel decides).
			a.pass2 = 1;
			e->timer_begin(TK_PASS2);
			launch_chunk(nt, m.has_transit != 0, grid, cthreads, st, a);
			f.verify = 0;
			f.skip_if_clean = 1;
			launch_finalize(nt, nwalk, st, f);
			e->timer_end(TK_PASS2);
			e->n_launch += 2;
		}
	}
	CK(cudaGetLastError());
	return GPT_OK;
}

// ---------------------------------------------------------------------------------------------
extern "C" {

#ifndef GPT_SRC_HASH
#define GPT_SRC_HASH "unknown"
#endif
const char *gpt_version(void) { return "gptransits_b200 0.1 (sm_100a, src=" GPT_SRC_HASH ")"; }

int gpt_engine_create(int device, gpt_engine **out)
{
	if (!out) return GPT_E_INVALID;
	*out = nullptr;
	int ndev = 0;
	if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) return GPT_E_CUDA;
	if (cudaSetDevice(device) != cudaSuccess) return GPT_E_CUDA;
	gpt_engine *e = new gpt_engine();
	e->device = device;
	{
		int v = 0;
		if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device) == cudaSuccess && v > 0) e->num_sms = v;
		if (cudaDeviceGetAttribute(&v, cudaDevAttrClockRate, device) == cudaSuccess && v > 0) e->sm_clock_ghz = v * 1e-6;
	}
	if (cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking) != cudaSuccess ||
	    cudaMallocHost((void **)&e->h_nbad, sizeof(int)) != cudaSuccess ||
	    cudaMallocHost((void **)&e->h_verr, sizeof(double)) != cudaSuccess) {
		delete e;
		return GPT_E_CUDA;
	}
	memset(&e->model, 0, sizeof(e->model));
	memset(&e->S, 0, sizeof(e->S));
	*out = e;
	return GPT_OK;
}

int gpt_star_clear(gpt_engine *e)
{
	if (!e) return GPT_E_INVALID;
	cudaSetDevice(e->device);
	cudaStreamSynchronize(e->stream);
	for (auto &s : e->stars)
		if (s.blob) cudaFree(s.blob);
	e->stars.clear();
	e->stage_star = -1;
	e->stars_dirty = true;
	e->spec_clean_steps = 0;  // new data: the clean record of the speculative sampler starts over
	return GPT_OK;
}

void gpt_engine_destroy(gpt_engine *e)
{
	if (e && e->comm) gpt_comm_destroy(e);
	if (!e) return;
	cudaSetDevice(e->device);
	cudaStreamSynchronize(e->stream);
	gpt_star_clear(e);
	e->d_stars.release();
	e->d_theta.release(); e->d_lp.release(); e->d_parts.release(); e->d_prep.release(); e->d_mean.release(); e->d_yup.release(); e->d_winmask.release(); e->d_zero.release();
	e->d_carryT.release(); e->d_carryE.release(); e->d_flags.release(); e->d_wstatus.release();
	e->d_partial.release(); e->d_partial2.release(); e->d_redo.release(); e->d_nbad.release(); e->d_verr.release(); e->d_acc.release();
	e->d_queue.release(); e->d_qcount.release(); e->d_qfused.release(); e->d_wbad.release();
	e->s_x.release(); e->s_lpx.release(); e->s_q.release(); e->s_lq.release(); e->s_fac.release();
	e->s_chain.release(); e->s_logp.release(); e->s_act.release(); e->s_ids.release(); e->s_nacc.release();
	e->s_snap_x.release(); e->s_snap_lpx.release(); e->s_snap_nacc.release();
	if (e->h_stage) cudaFreeHost(e->h_stage);
	if (e->h_nbad) cudaFreeHost(e->h_nbad);
	if (e->h_verr) cudaFreeHost(e->h_verr);
	cudaStreamDestroy(e->stream);
	delete e;
}

const char *gpt_last_error(const gpt_engine *e) { return e ? e->err.c_str() : "null engine"; }

int gpt_set_option(gpt_engine *e, const char *key, double value)
{
	if (!e || !key) return GPT_E_INVALID;
	std::string k(key);
	if (k == "halo") {
		e->opt_halo = (int)value;
	} else if (k == "chunks") {
		e->opt_chunks = (int)value;
	} else if (k == "verify") {
		e->opt_verify = value != 0;
	} else if (k == "transit_strict") {
		e->opt_transit_strict = value != 0;
	} else if (k == "series_order_min") {
		e->opt_series_order_min = std::min(4, std::max(0, (int)value));
	} else if (k == "fuse_scan") {
		e->opt_fuse_scan = value != 0;
	} else if (k == "speculate") {
		e->opt_speculate = std::min(2, std::max(0, (int)value));
		e->spec_last_clean = true;
	} else if (k == "scratch_cap_mb") {
		e->opt_scratch_cap_mb = std::max(1, (int)value);
	} else if (k == "async_egress") {
		e->opt_async_egress = value != 0;
	} else if (k == "pdl") {
		g_pdl = value != 0;
	} else if (k == "adapt_every") {
		e->opt_adapt_every = (int)value;
	} else if (k == "epoch_scan") {
		e->opt_epoch_scan = value != 0;
	} else if (k == "halo_auto") {
		e->opt_halo_auto = value != 0;
	} else if (k == "verify_tol") {
		e->opt_tol = value;
	} else if (k == "time_kernels") {
		e->opt_time = value != 0;
		if (!e->opt_time) e->timer_collect();
	} else if (k == "reset_timers") {
		e->timer_collect();
		for (int i = 0; i < TK_COUNT; i++) {
			e->t_ms[i] = 0;
			e->t_n[i] = 0;
		}
	} else if (k == "halo_start") {
		e->cur_halo = std::max(0, (int)value);
	} else {
		e->set_error("unknown option: " + k);
		return GPT_E_INVALID;
	}
	return GPT_OK;
}

int gpt_get_stat(gpt_engine *e, const char *key, double *value)
{
	if (!e || !key || !value) return GPT_E_INVALID;
	std::string k(key);
	if (k == "launches") *value = (double)e->n_launch;
	else if (k == "pass2_calls") *value = (double)e->n_pass2_calls;
	else if (k == "pass2_chunks") *value = (double)e->n_pass2_chunks;
	else if (k == "spec_replays") *value = (double)e->n_spec_replays;
	else if (k == "egress_jobs") *value = (double)e->n_egress_jobs;
	else if (k == "grid_reuse") *value = (double)e->n_grid_reuse;
	else if (k == "evals") *value = (double)e->n_evals;
	else if (k == "verify_ratio") *value = e->last_verr;
	else if (k == "chunks") *value = e->last_chunks;
	else if (k == "chunk_len") *value = e->last_L;
	else if (k == "halo") *value = e->last_H;
	else if (k == "halo_next") *value = e->opt_halo > 0 ? e->opt_halo : e->cur_halo;
	else if (k == "nstars") *value = (double)e->stars.size();
	else if (k == "num_sms") *value = e->num_sms;
	// nominal FP64 FMA peak at the device's maximum SM clock: SMs x 64 lanes x 2 FLOP x GHz (TFLOP/s)
	else if (k == "fp64_nominal_tflops") *value = e->num_sms * 64.0 * 2.0 * e->sm_clock_ghz * 1e-3;
	else if (k == "transit_items") {
		unsigned int q = 0;
		const unsigned int *src = e->last_fused ? (e->d_qfused.p ? e->d_qfused.p + 1 : nullptr) : e->d_qcount.p;
		if (src) {
			cudaStreamSynchronize(e->stream);
			cudaMemcpy(&q, src, sizeof(q), cudaMemcpyDeviceToHost);
		}
		*value = q;
	}
	else if (k.rfind("ms_", 0) == 0 || k.rfind("n_", 0) == 0) {
		static const char *names[TK_COUNT] = {"prep", "transit", "chunk", "final", "pass2", "trscan"};
		e->timer_collect();
		bool ms = k[0] == 'm';
		std::string which = k.substr(ms ? 3 : 2);
		for (int i = 0; i < TK_COUNT; i++)
			if (which == names[i]) {
				*value = ms ? e->t_ms[i] : (double)e->t_n[i];
				return GPT_OK;
			}
		e->set_error("unknown stat: " + k);
		return GPT_E_INVALID;
	}
	else {
		e->set_error("unknown stat: " + k);
		return GPT_E_INVALID;
	}
	return GPT_OK;
}

int gpt_model_set(gpt_engine *e, const gpt_model_desc *d)
{
	if (!e || !d) return GPT_E_INVALID;
	if (d->ncomp < 0 || d->ncomp > MAX_COMP || d->ndim < 1 || d->ndim > MAX_NDIM) {
		e->set_error("model: ncomp or ndim out of range");
		return GPT_E_INVALID;
	}
	if (d->ncomp == 0 && !d->has_transit) {
		e->set_error("model: need at least a gp or transit model");  // model.py:199-201
		return GPT_E_INVALID;
	}
	ModelDev m;
	memset(&m, 0, sizeof(m));
	m.ncomp = d->ncomp;
	int ndim_gp = 0, nterms = 0;
	bool may_generic = false;
	for (int c = 0; c < d->ncomp; c++) {
		int t = d->comp_type[c];
		if (t < 0 || t > 2) {
			e->set_error("model: component type not recognized");  // gp.py:26-27
			return GPT_E_INVALID;
		}
		m.comp_type[c] = t;
		m.comp_off[c] = ndim_gp;
		m.comp_term[c] = t != GPT_COMP_WHITE_NOISE ? nterms : -1;
		if (t == GPT_COMP_OSCILLATION_BUMP) {
			// Q is the second parameter; Q < 1/2 turns the term into two real celerite terms
			if (ndim_gp + 1 < d->ndim) {
				const gpt_prior &pq = d->priors[ndim_gp + 1];
				if (pq.kind == GPT_PRIOR_NORMAL || pq.a < 0.5) may_generic = true;
			}
		}
		ndim_gp += t == 0 ? 2 : (t == 1 ? 3 : 1);
		if (t != GPT_COMP_WHITE_NOISE) nterms++;
	}
	if (nterms > MAX_TERMS) {
		e->set_error("model: more than 4 SHO terms are not supported");
		return GPT_E_UNSUPPORTED;
	}
	int nfree = 0;
	if (d->has_transit)
		for (int i = 0; i < 9; i++) nfree += d->tr_free[i] ? 1 : 0;
	if (ndim_gp + nfree != d->ndim) {
		e->set_error("model: ndim does not match components + free transit parameters");
		return GPT_E_INVALID;
	}
	m.nterms = nterms;
	m.ndim = d->ndim;
	m.ndim_gp = ndim_gp;
	m.has_transit = d->has_transit ? 1 : 0;
	m.supersample = d->supersample < 1 ? 1 : d->supersample;
	m.exp_time = d->exp_time_days;
	m.ellip_mode = d->ellip_mode;
	m.inc_mode = d->inc_mode;
	for (int i = 0; i < d->ndim; i++) {
		if (d->priors[i].kind < 0 || d->priors[i].kind > 2) {
			e->set_error("model: prior distribution not recognized");  // component.py:42-44
			return GPT_E_INVALID;
		}
		m.prior_kind[i] = d->priors[i].kind;
		m.prior_a[i] = d->priors[i].a;
		m.prior_b[i] = d->priors[i].b;
	}
	for (int i = 0; i < 9; i++) {
		m.tr_fixed[i] = d->tr_fixed[i];
		m.tr_free[i] = d->tr_free[i];
	}
	// slowest decay rate inside the prior box: granulation c = 2 pi b / sqrt(2); bump c = 2 pi nu / (2 Q)
	{
		auto lo = [&](int i) { const gpt_prior &p = d->priors[i]; return p.kind == GPT_PRIOR_NORMAL ? p.a - 3 * p.b : p.a; };
		auto hi = [&](int i) { const gpt_prior &p = d->priors[i]; return p.kind == GPT_PRIOR_NORMAL ? p.a + 3 * p.b : p.b; };
		double cmin = 0.0;
		int i = 0;
		for (int c = 0; c < d->ncomp; c++) {
			double cc = 0.0;
			if (m.comp_type[c] == 0) {
				cc = 2 * kPi * std::max(lo(i + 1), 1e-300) / std::sqrt(2.0);
				i += 2;
			} else if (m.comp_type[c] == 1) {
				cc = 2 * kPi * std::max(lo(i + 2), 1e-300) / (2.0 * std::max(hi(i + 1), 0.5));
				i += 3;
			} else {
				i += 1;
				continue;
			}
			if (cc > 0 && (cmin == 0.0 || cc < cmin)) cmin = cc;
		}
		e->prior_cmin = cmin;
		e->cur_halo = 0;
		e->fail_halo = 0;
		e->fail_age = 0;
		e->spec_clean_steps = 0;
		e->spec_need_steps = 128;
		e->spec_last_clean = true;
	}
	e->model = m;
	e->may_generic = may_generic;
	e->have_model = true;
	return GPT_OK;
}

int gpt_star_upload(gpt_engine *e, int star, const double *t_days, const double *y, const double *yerr, int64_t N)
{
	if (!e || !t_days || !y || N < 1 || N > (int64_t)1 << 30 || star < 0) return GPT_E_INVALID;
	cudaSetDevice(e->device);
	if ((size_t)star >= e->stars.size()) e->stars.resize(star + 1);
	Star &s = e->stars[star];
	CK(cudaStreamSynchronize(e->stream));
	// layout of the star's blob; everything is assembled in pinned host memory and sent with one copy
	const int64_t ntiles = (N + TILE - 1) / TILE;
	const size_t off_samp = 0, off_tdays = sizeof(double4) * (size_t)N, off_tile = off_tdays + sizeof(double) * (size_t)N;
	const size_t off_irr = off_tile + sizeof(double) * (size_t)ntiles;
	const size_t stage_need = off_irr + sizeof(int) * ((size_t)N + 1);  // irr can hold every sample
	if (e->h_stage_cap < stage_need) {
		if (e->h_stage) cudaFreeHost(e->h_stage);
		e->h_stage = nullptr;
		e->h_stage_cap = 0;
		CK(cudaMallocHost((void **)&e->h_stage, stage_need + stage_need / 8));
		e->h_stage_cap = stage_need + stage_need / 8;
	}
	double4 *h = reinterpret_cast<double4 *>(e->h_stage + off_samp);
	double *h_tdays = reinterpret_cast<double *>(e->h_stage + off_tdays);
	double *tile_dt0 = reinterpret_cast<double *>(e->h_stage + off_tile);
	int *h_irr = reinterpret_cast<int *>(e->h_stage + off_irr);
	// The staging buffer still holds the blob of the last star sent.  New fluxes / error bars on the SAME time
	// stamps in the same slot (injection-recovery runs, detrending variants) keep the analysis of the time grid:
	// only the two changed fields of every sample are rewritten and the blob is sent again.
	if (e->stage_star == star && e->stage_N == N && s.samp && s.N == (int)N && s.has_err == (yerr != nullptr) &&
	    s.blob_used > 0 && memcmp(h_tdays, t_days, sizeof(double) * (size_t)N) == 0) {
		for (int64_t n = 0; n < N; n++) {
			const double err = yerr ? yerr[n] : 1.123e-12;
			h[n].y = y[n];
			h[n].z = err * err;
		}
		CK(cudaMemcpyAsync(s.blob, e->h_stage, s.blob_used, cudaMemcpyHostToDevice, e->stream));
		CK(cudaStreamSynchronize(e->stream));
		e->spec_clean_steps = 0;
		e->n_grid_reuse++;
		return GPT_OK;
	}
	e->stage_star = -1;  // (set again once the blob below is complete)
	memcpy(h_tdays, t_days, sizeof(double) * (size_t)N);
	std::vector<double> dts;
	dts.reserve(N);
	double xprev = 0.0;
	for (int64_t n = 0; n < N; n++) {
		double x = t_days[n] * kDaysToMs;  // model.py:183-187
		double dt = n ? x - xprev : 0.0;
		xprev = x;
		double err = yerr ? yerr[n] : 1.123e-12;  // celerite GP.compute default
		h[n].x = dt;
		h[n].y = y[n];
		h[n].z = err * err;
		h[n].w = t_days[n];
		if (n) dts.push_back(dt);
	}
	double dt0 = 0.0;
	if (!dts.empty()) {
		std::nth_element(dts.begin(), dts.begin() + dts.size() / 2, dts.end());
		dt0 = dts[dts.size() / 2];
	}
	// Regular steps are within 1e-3 of the median cadence.  Real timestamps drift slowly (barycentric
	// correction, ~1e-4 relative over a season), so the series of the fast loop expand around a PER-TILE
	// base cadence -- the mean regular step of the tile -- and dmax is the largest within-tile deviation.
	// cdev is the largest accumulated deviation |sum (dt_n - tile_dt0)| any run of regular steps inside a
	// tile can build up (range of the running sum, restarted at irregular samples): when max(c, d) * cdev
	// is negligible the loop advances by the tile's constant cadence and ignores the per-step deviation.
	double dreg = 1e-3 * std::fabs(dt0), dmax = 0.0, cdev = 0.0;
	for (int64_t tl = 0; tl < ntiles; tl++) tile_dt0[tl] = dt0;
	double prev_base = 0.0;
	bool have_prev = false;
	for (int64_t tl = 0; tl < ntiles; tl++) {
		const int64_t n_lo = tl * TILE, n_hi = std::min<int64_t>(N, n_lo + TILE);
		double sum = 0.0;
		int64_t cnt = 0;
		for (int64_t n = std::max<int64_t>(n_lo, 1); n < n_hi; n++)
			if (std::fabs(h[n].x - dt0) <= dreg) {
				sum += h[n].x - dt0;
				cnt++;
			}
		// largest accumulated deviation and largest single deviation of the tile for a base cadence
		auto tile_dev = [&](double base, double &dsingle) {
			double run = 0.0, lo = 0.0, hi = 0.0, worst = 0.0;
			dsingle = 0.0;
			for (int64_t n = n_lo; n < n_hi; n++) {
				if (n > 0 && std::fabs(h[n].x - dt0) <= dreg) {
					dsingle = std::max(dsingle, std::fabs(h[n].x - base));
					run += h[n].x - base;
					lo = std::min(lo, run);
					hi = std::max(hi, run);
					worst = std::max(worst, hi - lo);
				} else {
					run = lo = hi = 0.0;
				}
			}
			return worst;
		};
		double base = cnt ? dt0 + sum / (double)cnt : dt0;
		double ds = 0.0, dev = tile_dev(base, ds);
		// keep the previous tile's base when it serves this tile about as well (on an exact grid the tile
		// means differ by rounding only): the kernels recompute their per-walker constants only on a change
		if (have_prev && prev_base != base) {
			double ds_p = 0.0;
			const double dev_p = tile_dev(prev_base, ds_p);
			if (dev_p <= 1.5 * dev) {
				base = prev_base;
				dev = dev_p;
				ds = ds_p;
			}
		}
		tile_dt0[tl] = base;
		prev_base = base;
		have_prev = true;
		dmax = std::max(dmax, ds);
		cdev = std::max(cdev, dev);
	}
	s.N = (int)N;
	s.dt0 = dt0;
	s.dreg = dreg;
	s.dmax = dmax;
	s.cdev = cdev;
	s.has_err = yerr != nullptr;
	int nirr = 0;
	for (int64_t n = 0; n < N; n++)
		if (!(std::fabs(h[n].x - dt0) <= dreg)) h_irr[nirr++] = (int)n;
	s.nirr = nirr;
	const size_t total = off_irr + sizeof(int) * (size_t)std::max(1, nirr);
	if (s.blob_cap < total) {
		if (s.blob) cudaFree(s.blob);
		s.blob = nullptr;
		s.blob_cap = 0;
		CK(cudaMalloc((void **)&s.blob, total));
		s.blob_cap = total;
	}
	s.blob_used = total;
	s.samp = reinterpret_cast<double4 *>(s.blob + off_samp);
	s.tdays = reinterpret_cast<double *>(s.blob + off_tdays);
	s.tile_dt0 = reinterpret_cast<double *>(s.blob + off_tile);
	s.irr = reinterpret_cast<int *>(s.blob + off_irr);
	CK(cudaMemcpyAsync(s.blob, e->h_stage, total, cudaMemcpyHostToDevice, e->stream));
	CK(cudaStreamSynchronize(e->stream));
	e->stage_star = star;
	e->stage_N = N;
	e->stars_dirty = true;
	e->spec_clean_steps = 0;  // new data: the clean record of the speculative sampler starts over
	return GPT_OK;
}

// Many stars on one time grid (a survey batch: SURVEY.md 8 e1).  The grid is analysed once (gpt_star_upload of the
// first star); the other stars receive a device-to-device copy of its blob and their own fluxes, which cross the
// bus as one [nstars-1][N] matrix and are scattered into the samples by a kernel.
__global__ void star_fill_y_kernel(const StarDev *stars, int nstars, int N, const double *y)
{
	const int s = blockIdx.y;
	double4 *samp = const_cast<double4 *>(stars[s].samp);
	for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < N; n += gridDim.x * blockDim.x) samp[n].y = y[(size_t)s * N + n];
}

int gpt_stars_upload(gpt_engine *e, int star0, int nstars, const double *t_days, const double *y, const double *yerr, int64_t N)
{
	if (!e || !t_days || !y || nstars < 1 || star0 < 0) return GPT_E_INVALID;
	int rc = gpt_star_upload(e, star0, t_days, y, yerr, N);
	if (rc || nstars == 1) return rc;
	cudaSetDevice(e->device);
	if ((size_t)(star0 + nstars) > e->stars.size()) e->stars.resize(star0 + nstars);
	const Star proto = e->stars[star0];
	const size_t total = proto.blob_used;
	for (int i = 1; i < nstars; i++) {
		Star &d = e->stars[star0 + i];
		if (d.blob_cap < total) {
			if (d.blob) cudaFree(d.blob);
			d.blob = nullptr;
			d.blob_cap = 0;
			CK(cudaMalloc((void **)&d.blob, total));
			d.blob_cap = total;
		}
		d.blob_used = total;
		d.samp = reinterpret_cast<double4 *>(d.blob + ((char *)proto.samp - proto.blob));
		d.tdays = reinterpret_cast<double *>(d.blob + ((char *)proto.tdays - proto.blob));
		d.tile_dt0 = reinterpret_cast<double *>(d.blob + ((char *)proto.tile_dt0 - proto.blob));
		d.irr = reinterpret_cast<int *>(d.blob + ((char *)proto.irr - proto.blob));
		d.nirr = proto.nirr;
		d.N = proto.N;
		d.dt0 = proto.dt0;
		d.dreg = proto.dreg;
		d.dmax = proto.dmax;
		d.cdev = proto.cdev;
		d.has_err = proto.has_err;
		CK(cudaMemcpyAsync(d.blob, proto.blob, total, cudaMemcpyDeviceToDevice, e->stream));
	}
	CK(e->d_yup.ensure((size_t)(nstars - 1) * N));
	CK(cudaMemcpyAsync(e->d_yup.p, y + N, sizeof(double) * (size_t)(nstars - 1) * N, cudaMemcpyHostToDevice, e->stream));
	e->stars_dirty = true;
	rc = sync_stars(e);
	if (rc) return rc;
	star_fill_y_kernel<<<dim3((unsigned)std::min<int64_t>((N + 255) / 256, 64), nstars - 1), 256, 0, e->stream>>>(
		e->d_stars.p + star0 + 1, nstars - 1, (int)N, e->d_yup.p);
	e->n_launch++;
	CK(cudaGetLastError());
	CK(cudaStreamSynchronize(e->stream));
	return GPT_OK;
}

int gpt_log_prob_device(gpt_engine *e, int star0, int nstars, const double *theta_device, int W, double *lp_device,
			double *parts_device)
{
	if (!e || !theta_device || !lp_device || W < 1) return GPT_E_INVALID;
	cudaSetDevice(e->device);
	return eval_enqueue(e, star0, nstars, theta_device, W, lp_device, parts_device, false);
}

int gpt_log_prob_multi(gpt_engine *e, int star0, int nstars, const double *theta, int W, double *lp, double *parts)
{
	if (!e || !theta || !lp || W < 1 || nstars < 1) return GPT_E_INVALID;
	cudaSetDevice(e->device);
	if (!e->have_model) {
		e->set_error("no model: call gpt_model_set first");
		return GPT_E_NOMODEL;
	}
	const size_t nwalk = (size_t)nstars * W;
	const int ndim = e->model.ndim;
	CK(e->d_theta.ensure(nwalk * ndim));
	CK(e->d_lp.ensure(nwalk));
	CK(e->d_parts.ensure(nwalk * 4));
	CK(cudaMemcpyAsync(e->d_theta.p, theta, sizeof(double) * nwalk * ndim, cudaMemcpyHostToDevice, e->stream));
	int rc = eval_enqueue(e, star0, nstars, e->d_theta.p, W, e->d_lp.p, parts ? e->d_parts.p : nullptr, true);
	if (rc) return rc;
	CK(cudaMemcpyAsync(lp, e->d_lp.p, sizeof(double) * nwalk, cudaMemcpyDeviceToHost, e->stream));
	if (parts)
		CK(cudaMemcpyAsync(parts, e->d_parts.p, sizeof(double) * nwalk * 4, cudaMemcpyDeviceToHost, e->stream));
	CK(cudaStreamSynchronize(e->stream));
	return GPT_OK;
}

int gpt_log_prob(gpt_engine *e, int star, const double *theta, int W, double *lp, double *parts)
{
	return gpt_log_prob_multi(e, star, 1, theta, W, lp, parts);
}

int gpt_transit_flux(gpt_engine *e, int star, const double *pars9, int W, double *flux)
{
	if (!e || !pars9 || !flux || W < 1) return GPT_E_INVALID;
	cudaSetDevice(e->device);
	if (star < 0 || (size_t)star >= e->stars.size() || !e->stars[star].samp) {
		e->set_error("star slot empty");
		return GPT_E_NOSTAR;
	}
	if (!e->have_model) {
		e->set_error("no model: call gpt_model_set first");
		return GPT_E_NOMODEL;
	}
	int rc = sync_stars(e);
	if (rc) return rc;
	const int N = e->stars[star].N;
	CK(e->d_theta.ensure((size_t)W * 9));
	CK(e->d_prep.ensure((size_t)W * PREP_STRIDE));
	CK(e->d_flags.ensure(W));
	CK(e->d_mean.ensure((size_t)W * N));
	CK(e->d_queue.ensure((size_t)W * N));
	CK(e->d_qcount.ensure(1));
	CK(cudaMemsetAsync(e->d_qcount.p, 0, sizeof(unsigned int), e->stream));
	CK(cudaMemcpyAsync(e->d_theta.p, pars9, sizeof(double) * W * 9, cudaMemcpyHostToDevice, e->stream));
	prep_transit_direct_kernel<<<(W + 127) / 128, 128, 0, e->stream>>>(e->model.exp_time, e->d_theta.p, W, e->d_prep.p,
									    e->d_flags.p);
	dim3 g(W, (N + TR_THREADS - 1) / TR_THREADS);
	launch_k(transit_scan_kernel, dim3(g), dim3(TR_THREADS), 0, e->stream, e->d_stars.p + star, W, e->d_prep.p, e->d_flags.p, e->d_queue.p,
							     e->d_qcount.p, e->d_mean.p, 1, (unsigned int *)nullptr, 0);
	launch_transit_eval(e, e->stream, e->d_stars.p + star, W, e->model.supersample,
								   e->model.exp_time, e->model.ellip_mode, e->d_prep.p,
								   e->d_queue.p, e->d_qcount.p, e->d_mean.p, 1);
	e->n_launch += 3;
	CK(cudaGetLastError());
	CK(cudaMemcpyAsync(flux, e->d_mean.p, sizeof(double) * (size_t)W * N, cudaMemcpyDeviceToHost, e->stream));
	CK(cudaStreamSynchronize(e->stream));
	return GPT_OK;
}

// ------------------------------------------------------------------------------------- sampler
int gpt_sampler_begin(gpt_engine *e, int star0, int nstars, const int32_t *star_ids, const double *init, int W,
		      uint64_t seed, uint64_t step0, double a)
{
	if (!e || !init || W < 2 || (W & 1) || nstars < 1) {
		if (e) e->set_error("sampler: W must be even and >= 2");
		return GPT_E_INVALID;
	}
	cudaSetDevice(e->device);
	int rc = check_stars(e, star0, nstars);
	if (rc) return rc;
	const int ndim = e->model.ndim;
	// emcee refuses W < 2*ndim unless live_dangerously; keep the same guard
	if (W < 2 * ndim) {
		e->set_error("sampler: need at least 2*ndim walkers (emcee RedBlueMove)");
		return GPT_E_INVALID;
	}
	const size_t nw = (size_t)nstars * W, nh = (size_t)nstars * (W / 2);
	CK(e->s_x.ensure(nw * ndim));
	CK(e->s_lpx.ensure(nw));
	CK(e->s_q.ensure(nh * ndim));
	CK(e->s_lq.ensure(nh));
	CK(e->s_fac.ensure(nh));
	CK(e->s_act.ensure(nw));
	CK(e->s_ids.ensure(nstars));
	CK(e->s_nacc.ensure(nw));
	std::vector<int> ids(nstars);
	for (int i = 0; i < nstars; i++) ids[i] = star_ids ? star_ids[i] : star0 + i;
	CK(cudaMemcpyAsync(e->s_ids.p, ids.data(), sizeof(int) * nstars, cudaMemcpyHostToDevice, e->stream));
	CK(cudaMemcpyAsync(e->s_x.p, init, sizeof(double) * nw * ndim, cudaMemcpyHostToDevice, e->stream));
	CK(cudaMemsetAsync(e->s_nacc.p, 0, sizeof(long long) * nw, e->stream));
	CK(cudaStreamSynchronize(e->stream));
	SamplerDev &S = e->S;
	S.x = e->s_x.p; S.lpx = e->s_lpx.p; S.q = e->s_q.p; S.lq = e->s_lq.p; S.fac = e->s_fac.p;
	S.act = e->s_act.p; S.nacc = e->s_nacc.p; S.star_ids = e->s_ids.p;
	S.chain = nullptr; S.logp = nullptr; S.cap = 0;
	S.W = W; S.ndim = ndim; S.nstars = nstars; S.seed = seed; S.a = a;
	e->s_star0 = star0;
	e->s_step = step0;
	e->s_kept = 0;
	// initial log-probabilities of the whole ensemble
	rc = eval_enqueue(e, star0, nstars, S.x, W, S.lpx, nullptr, true);
	if (rc) return rc;
	CK(cudaStreamSynchronize(e->stream));
	e->sampling = true;
	return GPT_OK;
}

// ---- chain egress (the reference appends every step to its backend, model.py:204-220) ----------
// Finished step ranges of the device chain leave for the caller's host buffers while the sampler goes on: the
// sampling stream records an event at the end of a range, a helper thread makes the copy stream wait for it and
// issues the two strided device-to-host copies there (the caller's buffers are ordinary pageable memory: the
// copy call blocks its thread, which is why it is not the sampling thread), so only the last range is exposed.
struct ChainEgress {
	struct Job {
		cudaEvent_t ev;
		int64_t first, count;
	};
	gpt_engine *e;
	double *chain, *logp;  // host [ns][W][total][ndim], [ns][total][W]
	int64_t total, every;
	int64_t mark = 0;      // first kept step not yet handed over
	cudaStream_t cs = nullptr;
	std::thread th;
	std::mutex m;
	std::condition_variable cv;
	std::deque<Job> q;
	bool closing = false;
	cudaError_t err = cudaSuccess;
	long long njobs = 0;

	ChainEgress(gpt_engine *e_, double *chain_, double *logp_, int64_t total_, int64_t every_)
		: e(e_), chain(chain_), logp(logp_), total(total_), every(std::max<int64_t>(1, every_))
	{
	}
	cudaError_t start()
	{
		cudaError_t r = cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking);
		if (r != cudaSuccess) return r;
		th = std::thread([this] { loop(); });
		return cudaSuccess;
	}
	void loop()
	{
		cudaSetDevice(e->device);
		for (;;) {
			Job j;
			{
				std::unique_lock<std::mutex> lk(m);
				cv.wait(lk, [this] { return closing || !q.empty(); });
				if (q.empty()) return;
				j = q.front();
				q.pop_front();
			}
			const SamplerDev &S = e->S;
			const size_t W = S.W, ndim = S.ndim, ns = S.nstars, cap = S.cap;
			cudaError_t r = cudaStreamWaitEvent(cs, j.ev, 0);
			if (r == cudaSuccess && chain)
				r = cudaMemcpy2DAsync(chain + j.first * ndim, total * ndim * sizeof(double), S.chain + j.first * ndim,
						      cap * ndim * sizeof(double), j.count * ndim * sizeof(double), ns * W,
						      cudaMemcpyDeviceToHost, cs);
			if (r == cudaSuccess && logp)
				r = cudaMemcpy2DAsync(logp + j.first * W, total * W * sizeof(double), S.logp + j.first * W,
						      cap * W * sizeof(double), j.count * W * sizeof(double), ns, cudaMemcpyDeviceToHost, cs);
			if (r == cudaSuccess) r = cudaStreamSynchronize(cs);
			cudaEventDestroy(j.ev);
			if (r != cudaSuccess) {
				std::lock_guard<std::mutex> lk(m);
				if (err == cudaSuccess) err = r;
			}
		}
	}
	// the kept steps [mark, upto) are final on the sampling stream from this point of its queue on
	cudaError_t push(cudaStream_t st, int64_t upto)
	{
		if (upto <= mark) return cudaSuccess;
		Job j;
		cudaError_t r = cudaEventCreateWithFlags(&j.ev, cudaEventDisableTiming);
		if (r != cudaSuccess) return r;
		r = cudaEventRecord(j.ev, st);
		if (r != cudaSuccess) return r;
		j.first = mark;
		j.count = upto - mark;
		mark = upto;
		{
			std::lock_guard<std::mutex> lk(m);
			q.push_back(j);
			njobs++;
		}
		cv.notify_one();
		return cudaSuccess;
	}
	cudaError_t finish()
	{
		if (th.joinable()) {
			{
				std::lock_guard<std::mutex> lk(m);
				closing = true;
			}
			cv.notify_one();
			th.join();
		}
		if (cs) {
			cudaStreamDestroy(cs);
			cs = nullptr;
		}
		return err;
	}
	~ChainEgress() { finish(); }
};

static int sampler_run_impl(gpt_engine *e, int64_t nsteps, int keep, int flush_l2, float *ms_per_step, ChainEgress *eg = nullptr)
{
	if (!e || nsteps < 0) return GPT_E_INVALID;
	if (!e->sampling) {
		e->set_error("sampler: gpt_sampler_begin not called");
		return GPT_E_INVALID;
	}
	cudaSetDevice(e->device);
	SamplerDev &S = e->S;
	const int W = S.W, H = W / 2, ndim = S.ndim, ns = S.nstars;
	if (keep) {
		size_t nw = (size_t)ns * W;
		CK(e->s_chain.ensure(nw * nsteps * ndim));
		CK(e->s_logp.ensure(nw * nsteps));
		S.chain = e->s_chain.p;
		S.logp = e->s_logp.p;
		S.cap = nsteps;
		e->s_kept = 0;
	}
	cudaStream_t st = e->stream;
	const int nh = ns * H;
	std::vector<cudaEvent_t> ev;
	DevBuf<char> flush;
	const size_t flush_bytes = (size_t)256 << 20;  // > 126 MB L2
	if (ms_per_step) {
		ev.resize(2 * nsteps);
		for (auto &v : ev) CK(cudaEventCreate(&v));
		if (flush_l2) CK(flush.ensure(flush_bytes));
	}
	CK(e->d_acc.ensure(2));
	CK(cudaMemsetAsync(e->d_acc.p, 0, 2 * sizeof(unsigned long long), st));
	// verification verdicts accumulate on the device; every "adapt_every" steps (and at the end) they
	// are read back -- one small synchronisation -- to steer the halo length
	auto adapt = [&](bool *flagged, bool speculative) -> int {
		unsigned long long acc[2] = {0, 0};
		CK(cudaMemcpyAsync(acc, e->d_acc.p, sizeof(acc), cudaMemcpyDeviceToHost, st));
		CK(cudaMemsetAsync(e->d_acc.p, 0, 2 * sizeof(unsigned long long), st));
		CK(cudaStreamSynchronize(st));
		double ratio;
		memcpy(&ratio, &acc[1], sizeof(double));
		e->last_verr = ratio;
		if (flagged) *flagged = acc[0] > 0;
		if (acc[0] && !speculative) {
			e->n_pass2_calls++;
			e->n_pass2_chunks += (long long)acc[0];
		}
		if (e->last_chunks > 1) e->adapt_halo(e->last_H, acc[0] > 0, ratio);
		return GPT_OK;
	};
	auto run_steps = [&](int64_t sa, int64_t sb, bool speculative) -> int {
		CK(e->s_act.ensure((size_t)(sb - sa) * ns * W));
		for (int64_t s = sa; s < sb; s++) {
			if (ms_per_step) {
				if (flush_l2) CK(cudaMemsetAsync(flush.p, (int)(s & 0xff), flush_bytes, st));
				CK(cudaEventRecord(ev[2 * s], st));
			}
			if (s == sa) {
				// the random splits of all steps of the block: one launch (they depend on the RNG stream
				// only); inside the first step's timed region
				S.act = e->s_act.p;
				launch_k(sampler_split_kernel, dim3(ns, (unsigned)(sb - sa)), dim3(256), (size_t)W * (sizeof(unsigned long long) + sizeof(int)), st, S, e->s_step);
				e->n_launch++;
			}
			S.act = e->s_act.p + (size_t)(s - sa) * ns * W;
			for (int half = 0; half < 2; half++) {
				PrepPropose pr;
				pr.enabled = 1;
				pr.half = half;
				pr.step = e->s_step;
				pr.S = S;
				FinalAccept ac;
				ac.enabled = 1;
				ac.half = half;
				ac.step = e->s_step;
				ac.slot = keep ? e->s_kept : -1LL;
				ac.S = S;
				bool fused = false;
				int rc = eval_enqueue(e, e->s_star0, ns, S.q, H, S.lq, nullptr, false, &pr, &ac, &fused, speculative);
				if (rc) return rc;
				if (!fused) {
					launch_k(sampler_accept_kernel, dim3((nh + 127) / 128), dim3(128), 0, st, S, e->s_step, half, ac.slot);
					e->n_launch++;
				}
			}
			if (keep) e->s_kept++;
			if (ms_per_step) CK(cudaEventRecord(ev[2 * s + 1], st));
			e->s_step++;
			// a step of a block with its repair pass in line is final once enqueued: hand finished ranges over
			if (eg && keep && !speculative && (e->s_kept - eg->mark >= eg->every || s + 1 == sb)) CK(eg->push(st, e->s_kept));
		}
		CK(cudaGetLastError());
		return GPT_OK;
	};
	auto add_times = [&](int64_t sa, int64_t sb) -> int {
		if (!ms_per_step) return GPT_OK;
		for (int64_t s = sa; s < sb; s++) {
			float t = 0.f;
			CK(cudaEventElapsedTime(&t, ev[2 * s], ev[2 * s + 1]));
			ms_per_step[s] += t;
		}
		return GPT_OK;
	};
	if (ms_per_step)
		for (int64_t s = 0; s < nsteps; s++) ms_per_step[s] = 0.f;
	// Blocks of "adapt_every" steps.  The two repair launches per half-step are idle as long as no chunk is
	// flagged (C3: none in 3000 steps) and cost 3-4 % of a step, so once the sampler has a clean record of
	// spec_need_steps steps at the current halo a block runs speculatively: no repair pass is enqueued, the
	// chunk verification only accumulates its verdicts on the device.  They are read once per block (one small
	// synchronisation, also used to steer the halo); if any chunk was flagged the ensemble is rolled back
	// to the start of the block and the block is replayed with the repair pass -- the RNG is counter
	// based, so the replay reproduces the same moves and the result does not depend on the speculation --
	// and the record needed doubles.  Light curves whose chunks are flagged now and then (the bundled
	// examples) therefore never leave the repaired mode.  Option "speculate": 0 never, 1 this policy,
	// 2 speculate whenever the previous block was clean (tests).
	const bool can_speculate = e->opt_speculate && !e->may_generic && e->opt_verify;
	const int64_t block = e->opt_adapt_every > 0 ? std::min(e->opt_adapt_every, 1024) : 64;
	const size_t nw = (size_t)ns * W;
	for (int64_t sa = 0; sa < nsteps; sa += block) {
		const int64_t sb = std::min<int64_t>(nsteps, sa + block);
		const unsigned long long step_at = e->s_step;
		const long long kept_at = e->s_kept;
		const int halo_before = e->cur_halo;
		const bool spec = can_speculate && (e->opt_speculate == 2 ? e->spec_last_clean
									    : e->spec_clean_steps >= e->spec_need_steps);
		if (spec) {
			CK(e->s_snap_x.ensure(nw * ndim));
			CK(e->s_snap_lpx.ensure(nw));
			CK(e->s_snap_nacc.ensure(nw));
			CK(cudaMemcpyAsync(e->s_snap_x.p, S.x, nw * ndim * sizeof(double), cudaMemcpyDeviceToDevice, st));
			CK(cudaMemcpyAsync(e->s_snap_lpx.p, S.lpx, nw * sizeof(double), cudaMemcpyDeviceToDevice, st));
			CK(cudaMemcpyAsync(e->s_snap_nacc.p, S.nacc, nw * sizeof(long long), cudaMemcpyDeviceToDevice, st));
		}
		int rc = run_steps(sa, sb, spec);
		if (rc) return rc;
		bool flagged = false;
		rc = adapt(&flagged, spec);
		if (rc) return rc;
		rc = add_times(sa, sb);
		if (rc) return rc;
		if (eg && keep && spec && !flagged) CK(eg->push(st, e->s_kept));  // a speculative block is final once its record is clean
		e->spec_last_clean = !flagged;
		if (flagged || e->cur_halo != halo_before) e->spec_clean_steps = 0;
		else e->spec_clean_steps += sb - sa;
		if (spec && flagged) {
			e->n_spec_replays++;
			e->spec_need_steps = std::min<long long>(2 * e->spec_need_steps, 32768);
			CK(cudaMemcpyAsync(S.x, e->s_snap_x.p, nw * ndim * sizeof(double), cudaMemcpyDeviceToDevice, st));
			CK(cudaMemcpyAsync(S.lpx, e->s_snap_lpx.p, nw * sizeof(double), cudaMemcpyDeviceToDevice, st));
			CK(cudaMemcpyAsync(S.nacc, e->s_snap_nacc.p, nw * sizeof(long long), cudaMemcpyDeviceToDevice, st));
			e->s_step = step_at;
			e->s_kept = kept_at;
			rc = run_steps(sa, sb, false);
			if (rc) return rc;
			rc = adapt(nullptr, false);
			if (rc) return rc;
			rc = add_times(sa, sb);
			if (rc) return rc;
		}
	}
	CK(cudaStreamSynchronize(st));
	if (ms_per_step) {
		for (auto &v : ev) cudaEventDestroy(v);
		flush.release();
	}
	return GPT_OK;
}

int gpt_sampler_run(gpt_engine *e, int64_t nsteps, int keep) { return sampler_run_impl(e, nsteps, keep, 0, nullptr); }

int gpt_sampler_run_timed(gpt_engine *e, int64_t nsteps, int keep, int flush_l2, float *ms_per_step)
{
	if (!ms_per_step) return GPT_E_INVALID;
	return sampler_run_impl(e, nsteps, keep, flush_l2, ms_per_step);
}

int gpt_log_prob_device_timed(gpt_engine *e, int star0, int nstars, const double *theta_device, int W,
			      double *lp_device, int reps, int flush_l2, float *ms_per_rep)
{
	if (!e || !theta_device || !lp_device || W < 1 || reps < 1 || !ms_per_rep) return GPT_E_INVALID;
	cudaSetDevice(e->device);
	std::vector<cudaEvent_t> ev(2 * reps);
	for (auto &v : ev) CK(cudaEventCreate(&v));
	DevBuf<char> flush;
	const size_t flush_bytes = (size_t)256 << 20;
	if (flush_l2) CK(flush.ensure(flush_bytes));
	for (int r = 0; r < reps; r++) {
		if (flush_l2) CK(cudaMemsetAsync(flush.p, r & 0xff, flush_bytes, e->stream));
		CK(cudaEventRecord(ev[2 * r], e->stream));
		int rc = eval_enqueue(e, star0, nstars, theta_device, W, lp_device, nullptr, false);
		if (rc) return rc;
		CK(cudaEventRecord(ev[2 * r + 1], e->stream));
	}
	CK(cudaStreamSynchronize(e->stream));
	for (int r = 0; r < reps; r++) CK(cudaEventElapsedTime(&ms_per_rep[r], ev[2 * r], ev[2 * r + 1]));
	for (auto &v : ev) cudaEventDestroy(v);
	flush.release();
	return GPT_OK;
}

int gpt_device_alloc(gpt_engine *e, size_t bytes, void **out)
{
	if (!e || !out) return GPT_E_INVALID;
	cudaSetDevice(e->device);
	CK(cudaMalloc(out, bytes));
	return GPT_OK;
}
int gpt_device_free(gpt_engine *e, void *p)
{
	if (!e) return GPT_E_INVALID;
	cudaSetDevice(e->device);
	CK(cudaFree(p));
	return GPT_OK;
}
int gpt_device_copy(gpt_engine *e, void *dst, const void *src, size_t bytes, int to_device)
{
	if (!e) return GPT_E_INVALID;
	cudaSetDevice(e->device);
	CK(cudaMemcpyAsync(dst, src, bytes, to_device ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost, e->stream));
	CK(cudaStreamSynchronize(e->stream));
	return GPT_OK;
}

int gpt_sampler_read(gpt_engine *e, int64_t first, int64_t count, double *chain, double *logp)
{
	if (!e || !e->sampling || first < 0 || count < 0 || first + count > e->s_kept) {
		if (e) e->set_error("sampler_read: range outside the kept steps");
		return GPT_E_INVALID;
	}
	cudaSetDevice(e->device);
	const SamplerDev &S = e->S;
	const size_t W = S.W, ndim = S.ndim, ns = S.nstars, cap = S.cap;
	// chain [ns][W][count][ndim] <- device [ns][W][cap][ndim]
	if (chain)
		CK(cudaMemcpy2DAsync(chain, count * ndim * sizeof(double), S.chain + first * ndim, cap * ndim * sizeof(double),
				     count * ndim * sizeof(double), ns * W, cudaMemcpyDeviceToHost, e->stream));
	if (logp)
		CK(cudaMemcpy2DAsync(logp, count * W * sizeof(double), S.logp + first * W, cap * W * sizeof(double),
				     count * W * sizeof(double), ns, cudaMemcpyDeviceToHost, e->stream));
	CK(cudaStreamSynchronize(e->stream));
	return GPT_OK;
}

int gpt_sampler_state(gpt_engine *e, double *coords, double *logp, int64_t *naccept)
{
	if (!e || !e->sampling) return GPT_E_INVALID;
	cudaSetDevice(e->device);
	const SamplerDev &S = e->S;
	const size_t nw = (size_t)S.nstars * S.W;
	if (coords) CK(cudaMemcpyAsync(coords, S.x, sizeof(double) * nw * S.ndim, cudaMemcpyDeviceToHost, e->stream));
	if (logp) CK(cudaMemcpyAsync(logp, S.lpx, sizeof(double) * nw, cudaMemcpyDeviceToHost, e->stream));
	if (naccept) {
		static_assert(sizeof(long long) == sizeof(int64_t), "");
		CK(cudaMemcpyAsync(naccept, S.nacc, sizeof(int64_t) * nw, cudaMemcpyDeviceToHost, e->stream));
	}
	CK(cudaStreamSynchronize(e->stream));
	return GPT_OK;
}

int gpt_sampler_rhat(gpt_engine *e, int64_t first, int64_t count, double *rhat)
{
	if (!e || !rhat || !e->sampling || first < 0 || count < 2 || first + count > e->s_kept) {
		if (e) e->set_error("sampler_rhat: need at least 2 kept steps inside the kept range");
		return GPT_E_INVALID;
	}
	cudaSetDevice(e->device);
	const SamplerDev &S = e->S;
	if (S.W > 1024) {
		e->set_error("sampler_rhat: more than 1024 walkers per star are not supported");
		return GPT_E_UNSUPPORTED;
	}
	const int nb = S.nstars * S.ndim;
	DevBuf<double> out;
	CK(out.ensure(nb));
	int threads = 32;
	while (threads < S.W && threads < 256) threads *= 2;
	sampler_rhat_kernel<<<nb, threads, 2 * sizeof(double) * std::max(threads, S.W), e->stream>>>(S, first, count, out.p);
	e->n_launch++;
	CK(cudaGetLastError());
	CK(cudaMemcpyAsync(rhat, out.p, sizeof(double) * nb, cudaMemcpyDeviceToHost, e->stream));
	CK(cudaStreamSynchronize(e->stream));
	out.release();
	return GPT_OK;
}

int gpt_sampler_geweke(gpt_engine *e, int64_t first, int64_t count, int nbins, const int64_t *starts, double frac1,
		       double frac2, double *z)
{
	if (!e || !z || !starts || !e->sampling || first < 0 || count < 2 || nbins < 1 || first + count > e->s_kept) {
		if (e) e->set_error("sampler_geweke: range outside the kept steps");
		return GPT_E_INVALID;
	}
	for (int b = 0; b < nbins; b++)
		if (starts[b] < 0 || starts[b] >= count) {
			e->set_error("sampler_geweke: a bin starts outside the range");
			return GPT_E_INVALID;
		}
	cudaSetDevice(e->device);
	const SamplerDev &S = e->S;
	const size_t nz = (size_t)S.nstars * nbins * S.W * S.ndim;
	DevBuf<double> out;
	DevBuf<long long> st;
	CK(out.ensure(nz));
	CK(st.ensure(nbins));
	static_assert(sizeof(long long) == sizeof(int64_t), "");
	CK(cudaMemcpyAsync(st.p, starts, sizeof(int64_t) * nbins, cudaMemcpyHostToDevice, e->stream));
	sampler_geweke_kernel<<<(unsigned)((nz + 127) / 128), 128, 0, e->stream>>>(S, first, count, nbins, st.p, frac1, frac2, out.p);
	e->n_launch++;
	CK(cudaGetLastError());
	CK(cudaMemcpyAsync(z, out.p, sizeof(double) * nz, cudaMemcpyDeviceToHost, e->stream));
	CK(cudaStreamSynchronize(e->stream));
	out.release();
	st.release();
	return GPT_OK;
}

// ---- single star, ensemble sharded by walker across ranks (one process per GPU) ------------------
// Every rank holds the whole ensemble.  A half-step is: gpt_sampler_half_eval (proposals of all active
// walkers; log-probabilities of this rank's slice [lo, hi) of them into lq) -> all-gather of lq by the
// caller (NCCL on the device pointer gpt_sampler_lq_device returns) -> gpt_sampler_half_accept (identical
// accept/reject on every rank: the replicated state never diverges).
int gpt_sampler_reserve(gpt_engine *e, int64_t nsteps)
{
	if (!e || !e->sampling || nsteps < 0) return GPT_E_INVALID;
	cudaSetDevice(e->device);
	SamplerDev &S = e->S;
	const size_t nw = (size_t)S.nstars * S.W;
	CK(e->s_chain.ensure(nw * std::max<int64_t>(nsteps, 1) * S.ndim));
	CK(e->s_logp.ensure(nw * std::max<int64_t>(nsteps, 1)));
	S.chain = e->s_chain.p;
	S.logp = e->s_logp.p;
	S.cap = nsteps;
	e->s_kept = 0;
	return GPT_OK;
}

int gpt_sampler_lq_device(gpt_engine *e, void **lq, int64_t *count)
{
	if (!e || !e->sampling || !lq || !count) return GPT_E_INVALID;
	*lq = e->S.lq;
	*count = (int64_t)e->S.nstars * (e->S.W / 2);
	return GPT_OK;
}

int gpt_sampler_half_eval(gpt_engine *e, int half, int lo, int hi)
{
	if (!e || !e->sampling || half < 0 || half > 1) return GPT_E_INVALID;
	SamplerDev &S = e->S;
	const int H = S.W / 2;
	if (S.nstars != 1 || lo < 0 || hi > H || lo > hi) {
		e->set_error("sampler_half_eval: one star, 0 <= lo <= hi <= W/2");
		return GPT_E_INVALID;
	}
	cudaSetDevice(e->device);
	cudaStream_t st = e->stream;
	if (half == 0) {
		CK(e->s_act.ensure((size_t)S.W));
		S.act = e->s_act.p;
		launch_k(sampler_split_kernel, dim3(1, 1), dim3(256), (size_t)S.W * (sizeof(unsigned long long) + sizeof(int)), st, S, e->s_step);
		e->n_launch++;
	}
	launch_k(sampler_propose_kernel, dim3((H + 127) / 128), dim3(128), 0, st, S, e->s_step, half);
	e->n_launch++;
	if (hi > lo) {
		int rc = eval_enqueue(e, e->s_star0, 1, S.q + (size_t)lo * S.ndim, hi - lo, S.lq + lo, nullptr, false);
		if (rc) return rc;
	}
	CK(cudaGetLastError());
	CK(cudaStreamSynchronize(st));  // the caller's collective runs on its own stream
	return GPT_OK;
}

int gpt_sampler_half_accept(gpt_engine *e, int half, int keep)
{
	if (!e || !e->sampling || half < 0 || half > 1) return GPT_E_INVALID;
	SamplerDev &S = e->S;
	if (keep && (!S.chain || e->s_kept >= S.cap)) {
		e->set_error("sampler_half_accept: gpt_sampler_reserve first (chain full)");
		return GPT_E_INVALID;
	}
	cudaSetDevice(e->device);
	const int nh = S.nstars * (S.W / 2);
	launch_k(sampler_accept_kernel, dim3((nh + 127) / 128), dim3(128), 0, e->stream, S, e->s_step, half, keep ? e->s_kept : -1LL);
	e->n_launch++;
	if (half == 1) {
		if (keep) e->s_kept++;
		e->s_step++;
	}
	CK(cudaGetLastError());
	return GPT_OK;
}

// ---- walker-sharded sampling of one star across GPUs: the all-gather runs on the engine's stream -------------
namespace {
struct NcclApi {
	void *handle = nullptr;
	ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
	ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
	ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
	const char *(*GetErrorString)(ncclResult_t) = nullptr;
	bool ok = false;
};
static NcclApi &nccl_api()
{
	static NcclApi api;
	if (api.handle) return api;
	// the copy a host framework has already loaded is found first; otherwise the system library
	for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
		api.handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
		if (api.handle) break;
	}
	if (!api.handle) return api;
	api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(api.handle, "ncclGetUniqueId");
	api.CommInitRank = (decltype(api.CommInitRank))dlsym(api.handle, "ncclCommInitRank");
	api.AllGather = (decltype(api.AllGather))dlsym(api.handle, "ncclAllGather");
	api.CommDestroy = (decltype(api.CommDestroy))dlsym(api.handle, "ncclCommDestroy");
	api.GetErrorString = (decltype(api.GetErrorString))dlsym(api.handle, "ncclGetErrorString");
	api.ok = api.GetUniqueId && api.CommInitRank && api.AllGather && api.CommDestroy && api.GetErrorString;
	return api;
}
}  // namespace

int gpt_comm_id(void *id128)
{
	if (!id128) return GPT_E_INVALID;
	NcclApi &n = nccl_api();
	if (!n.ok) return GPT_E_CUDA;
	static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
	ncclUniqueId id;
	if (n.GetUniqueId(&id) != ncclSuccess) return GPT_E_CUDA;
	memcpy(id128, &id, sizeof(id));
	return GPT_OK;
}

int gpt_comm_init(gpt_engine *e, const void *id128, int rank, int world)
{
	if (!e || !id128 || world < 1 || rank < 0 || rank >= world) return GPT_E_INVALID;
	NcclApi &n = nccl_api();
	if (!n.ok) {
		e->set_error("gpt_comm_init: libnccl.so.2 not found");
		return GPT_E_CUDA;
	}
	cudaSetDevice(e->device);
	if (e->comm) {
		n.CommDestroy(e->comm);
		e->comm = nullptr;
	}
	ncclUniqueId id;
	memcpy(&id, id128, sizeof(id));
	ncclResult_t r = n.CommInitRank(&e->comm, world, id, rank);
	if (r != ncclSuccess) {
		e->set_error(std::string("ncclCommInitRank: ") + n.GetErrorString(r));
		e->comm = nullptr;
		return GPT_E_CUDA;
	}
	e->comm_rank = rank;
	e->comm_world = world;
	return GPT_OK;
}

int gpt_comm_destroy(gpt_engine *e)
{
	if (!e) return GPT_E_INVALID;
	if (e->comm) {
		cudaSetDevice(e->device);
		cudaStreamSynchronize(e->stream);
		nccl_api().CommDestroy(e->comm);
		e->comm = nullptr;
	}
	e->comm_rank = 0;
	e->comm_world = 1;
	return GPT_OK;
}

int gpt_sampler_run_sharded(gpt_engine *e, int64_t nsteps, int keep, float *ms_per_step)
{
	if (!e || !e->sampling || nsteps < 0) return GPT_E_INVALID;
	SamplerDev &S = e->S;
	const int H = S.W / 2, world = e->comm_world, rank = e->comm_rank;
	if (S.nstars != 1) {
		e->set_error("sampler_run_sharded: one star");
		return GPT_E_INVALID;
	}
	if (world > 1 && (!e->comm || H % world)) {
		e->set_error("sampler_run_sharded: gpt_comm_init first, and the half-ensemble must divide by the number of ranks");
		return GPT_E_INVALID;
	}
	if (keep && (!S.chain || e->s_kept + nsteps > S.cap)) {
		e->set_error("sampler_run_sharded: gpt_sampler_reserve first (chain too short)");
		return GPT_E_INVALID;
	}
	cudaSetDevice(e->device);
	cudaStream_t st = e->stream;
	NcclApi &n = nccl_api();
	const int per = H / world, lo = rank * per;
	std::vector<cudaEvent_t> ev;
	if (ms_per_step) {
		ev.resize(2 * nsteps);
		for (auto &v : ev) CK(cudaEventCreate(&v));
	}
	CK(e->s_act.ensure((size_t)S.W));
	S.act = e->s_act.p;
	if (!e->d_acc.p) {
		CK(e->d_acc.ensure(2));
		CK(cudaMemsetAsync(e->d_acc.p, 0, 2 * sizeof(unsigned long long), st));
	}
	for (int64_t s = 0; s < nsteps; s++) {
		if (ms_per_step) CK(cudaEventRecord(ev[2 * s], st));
		launch_k(sampler_split_kernel, dim3(1, 1), dim3(256), (size_t)S.W * (sizeof(unsigned long long) + sizeof(int)), st, S, e->s_step);
		e->n_launch++;
		for (int half = 0; half < 2; half++) {
			// every rank proposes for the whole active half (the same Philox stream everywhere) ...
			launch_k(sampler_propose_kernel, dim3((H + 127) / 128), dim3(128), 0, st, S, e->s_step, half);
			// ... evaluates its slice of the proposals ...
			int rc = eval_enqueue(e, e->s_star0, 1, S.q + (size_t)lo * S.ndim, per, S.lq + lo, nullptr, false);
			if (rc) return rc;
			// ... the slices are gathered in place over NVLink (the only collective of the path) ...
			if (world > 1) {
				ncclResult_t r = n.AllGather(S.lq + lo, S.lq, (size_t)per, ncclDouble, e->comm, st);
				if (r != ncclSuccess) {
					e->set_error(std::string("ncclAllGather: ") + n.GetErrorString(r));
					return GPT_E_CUDA;
				}
			}
			// ... and every rank applies the identical accept/reject
			launch_k(sampler_accept_kernel, dim3((H + 127) / 128), dim3(128), 0, st, S, e->s_step, half, keep ? e->s_kept : -1LL);
			e->n_launch += 2;
		}
		if (keep) e->s_kept++;
		e->s_step++;
		if (ms_per_step) CK(cudaEventRecord(ev[2 * s + 1], st));
	}
	CK(cudaGetLastError());
	// the verification verdicts of this rank's evaluations steer its halo (one synchronisation per call)
	unsigned long long acc[2] = {0, 0};
	CK(cudaMemcpyAsync(acc, e->d_acc.p, sizeof(acc), cudaMemcpyDeviceToHost, st));
	CK(cudaMemsetAsync(e->d_acc.p, 0, sizeof(acc), st));
	CK(cudaStreamSynchronize(st));
	{
		double ratio;
		memcpy(&ratio, &acc[1], sizeof(double));
		e->last_verr = ratio;
		if (acc[0]) {
			e->n_pass2_calls++;
			e->n_pass2_chunks += (long long)acc[0];
		}
		if (e->last_H > 0) e->adapt_halo(e->last_H, acc[0] != 0, ratio);
	}
	if (ms_per_step) {
		for (int64_t s = 0; s < nsteps; s++) CK(cudaEventElapsedTime(&ms_per_step[s], ev[2 * s], ev[2 * s + 1]));
		for (auto &v : ev) cudaEventDestroy(v);
	}
	return GPT_OK;
}

int gpt_sampler_end(gpt_engine *e)
{
	if (!e) return GPT_E_INVALID;
	e->sampling = false;
	return GPT_OK;
}

int gpt_sample(gpt_engine *e, int star0, int nstars, const double *init, int W, int64_t nsteps, uint64_t seed,
	       uint64_t step0, double a, double *chain, double *logp, int64_t *naccept)
{
	if (!e || !chain || !logp) return GPT_E_INVALID;
	int rc = gpt_sampler_begin(e, star0, nstars, nullptr, init, W, seed, step0, a);
	if (rc) return rc;
	// the chain leaves for the caller's buffers range by range while the sampler runs (ranges of at least ~1 KB per
	// walker row, at most an eighth of the run, so that the strided copies stay efficient and the last one is short)
	{
		const int64_t row = std::max<int64_t>(1, 1024 / ((int64_t)e->S.ndim * (int64_t)sizeof(double)));
		ChainEgress eg(e, chain, logp, nsteps, std::max<int64_t>(row, (nsteps + 7) / 8));
		if (e->opt_async_egress) CK(eg.start());
		rc = sampler_run_impl(e, nsteps, 1, 0, nullptr, e->opt_async_egress ? &eg : nullptr);
		cudaError_t er = eg.finish();
		if (rc) return rc;
		if (er != cudaSuccess) {
			e->set_error(std::string("chain egress: ") + cudaGetErrorString(er));
			return GPT_E_CUDA;
		}
		if (!e->opt_async_egress || eg.mark < nsteps) {
			rc = gpt_sampler_read(e, 0, nsteps, chain, logp);
			if (rc) return rc;
		}
		e->n_egress_jobs = eg.njobs;
	}
	if (naccept) {
		rc = gpt_sampler_state(e, nullptr, nullptr, naccept);
		if (rc) return rc;
	}
	return gpt_sampler_end(e);
}

int gpt_gp_predict(gpt_engine *e, int star, const double *theta, const double *x_days, int64_t M, double *mu,
		   double *var)
{
	if (!e || !theta || !x_days || !mu || M < 1) return GPT_E_INVALID;
	cudaSetDevice(e->device);
	int rc = check_stars(e, star, 1);
	if (rc) return rc;
	rc = sync_stars(e);
	if (rc) return rc;
	const ModelDev &m = e->model;
	if (m.nterms < 1) {
		e->set_error("gp_predict: the model has no correlated (SHO) term");
		return GPT_E_UNSUPPORTED;
	}
	const int N = e->stars[star].N, J = 2 * m.nterms;
	if ((double)N * (double)M > 4e9) {
		e->set_error("gp_predict: N * M too large for the dense prediction kernel");
		return GPT_E_UNSUPPORTED;
	}
	cudaStream_t st = e->stream;
	CK(e->d_theta.ensure(std::max<size_t>(m.ndim, (size_t)M)));
	CK(e->d_prep.ensure(PREP_STRIDE));
	CK(e->d_flags.ensure(1));
	CK(e->d_wstatus.ensure(1));
	CK(e->d_nbad.ensure(1));
	CK(e->d_verr.ensure(1));
	DevBuf<double> fac, xs, out;
	CK(fac.ensure((size_t)N * (2 + 3 * J)));
	CK(xs.ensure(M));
	CK(out.ensure(2 * (size_t)M));
	CK(cudaMemcpyAsync(e->d_theta.p, theta, sizeof(double) * m.ndim, cudaMemcpyHostToDevice, st));
	CK(cudaMemcpyAsync(xs.p, x_days, sizeof(double) * M, cudaMemcpyHostToDevice, st));
	const StarDev *dst = e->d_stars.p + star;
	if (m.has_transit) {
		CK(e->d_mean.ensure(N));
		CK(e->d_queue.ensure(N));
		CK(e->d_qcount.ensure(1));
	}
	{
		PrepReset reset;
		reset.wstatus = e->d_wstatus.p; reset.nbad = e->d_nbad.p; reset.verr = e->d_verr.p;
		reset.qcount = m.has_transit ? e->d_qcount.p : nullptr;
		PrepPropose pr;
		memset(&pr, 0, sizeof(pr));
		PrepScan sc;
		memset(&sc, 0, sizeof(sc));
		launch_k(prep_kernel<false>, dim3(1), dim3(PREP_THREADS), 0, st, m, dst, 1, e->d_theta.p, 1, e->d_prep.p, e->d_flags.p, reset, pr, sc);
	}
	if (m.has_transit) {
		dim3 g(1, (N + TR_THREADS - 1) / TR_THREADS);
		launch_k(transit_scan_kernel, dim3(g), dim3(TR_THREADS), 0, st, dst, 1, e->d_prep.p, e->d_flags.p, e->d_queue.p, e->d_qcount.p,
							      e->d_mean.p, 0, (unsigned int *)nullptr, 0);
		launch_transit_eval(e, st, dst, 1, m.supersample, m.exp_time, m.ellip_mode,
										e->d_prep.p, e->d_queue.p, e->d_qcount.p, e->d_mean.p, 0);
	}
	PredictArgs a;
	a.star = dst; a.prep = e->d_prep.p; a.flags = e->d_flags.p; a.mean = e->d_mean.p;
	a.nterms = m.nterms; a.N = N; a.transit = m.has_transit;
	a.D = fac.p; a.alpha = fac.p + N; a.W = fac.p + 2 * (size_t)N; a.U = a.W + (size_t)N * J; a.phi = a.U + (size_t)N * J;
	a.status = e->d_wstatus.p;
	predict_factor_kernel<<<1, 32, 0, st>>>(a);
	predict_points_kernel<<<(unsigned)((M + 127) / 128), 128, 0, st>>>(a, xs.p, (long long)M, out.p, var ? out.p + M : nullptr);
	e->n_launch += 3 + (m.has_transit ? 2 : 0);
	CK(cudaGetLastError());
	int status = 0, flags = 0;
	CK(cudaMemcpyAsync(&status, e->d_wstatus.p, sizeof(int), cudaMemcpyDeviceToHost, st));
	CK(cudaMemcpyAsync(&flags, e->d_flags.p, sizeof(int), cudaMemcpyDeviceToHost, st));
	CK(cudaMemcpyAsync(mu, out.p, sizeof(double) * M, cudaMemcpyDeviceToHost, st));
	if (var) CK(cudaMemcpyAsync(var, out.p + M, sizeof(double) * M, cudaMemcpyDeviceToHost, st));
	CK(cudaStreamSynchronize(st));
	fac.release();
	xs.release();
	out.release();
	if (!(flags & WF_VALID)) {
		e->set_error("gp_predict: parameter vector outside the prior support");
		return GPT_E_INVALID;
	}
	if (status & WF_NOTPD) {
		e->set_error("gp_predict: covariance matrix is not positive definite");  // celerite: LinAlgError
		return GPT_E_INVALID;
	}
	return GPT_OK;
}

int gpt_measure_fp64_peak(gpt_engine *e, double *tflops)
{
	if (!e || !tflops) return GPT_E_INVALID;
	cudaSetDevice(e->device);
	const int blocks = e->num_sms * 8, threads = 256, iters = 1 << 14;
	DevBuf<double> out;
	CK(out.ensure((size_t)blocks * threads));
	cudaEvent_t a, b;
	CK(cudaEventCreate(&a));
	CK(cudaEventCreate(&b));
	double best = 0.0;
	for (int rep = 0; rep < 5; rep++) {
		CK(cudaEventRecord(a, e->stream));
		dfma_peak_kernel<<<blocks, threads, 0, e->stream>>>(out.p, iters);
		CK(cudaEventRecord(b, e->stream));
		CK(cudaEventSynchronize(b));
		float ms = 0;
		CK(cudaEventElapsedTime(&ms, a, b));
		double fl = 2.0 * 8.0 * (double)iters * blocks * threads;
		if (rep > 0) best = std::max(best, fl / (ms * 1e-3) / 1e12);
	}
	e->n_launch += 5;
	cudaEventDestroy(a);
	cudaEventDestroy(b);
	out.release();
	*tflops = best;
	return GPT_OK;
}

}  // extern "C"
