// device_math.cuh -- per-walker scalar math of the gptransits posterior, FP64, sm_100a.
//
//   priors              component.py:33-51, transit.py:56-64,120-122 (scipy.stats frozen log-pdfs)
//   GP transforms       component.py:83-101,128-140,168-172 + celerite SHOTerm/JitterTerm coefficients
//   transit geometry    batman _rsky.c as called from transit.py:91,126
//   limb-darkened flux  batman _quadratic_ld.c (Mandel & Agol 2002), Hastings or exact K/E
//   Philox4x32-10       counter-based RNG of the stretch-move sampler (replaces emcee's MT19937)
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include <math_constants.h>

namespace gpt {

constexpr int MAX_TERMS = 4;   // SHO terms handled by the register-resident kernels (J = 2*terms)
constexpr int MAX_COMP = 8;    // GP components per model
constexpr int MAX_NDIM = 32;   // free parameters per walker

constexpr double kPi = 3.14159265358979323846;
constexpr double kTwoPi = 6.28318530717958647692;
constexpr double kDaysToMs = 0.0864;  // model.py:183: (24*3600)/1e6, days -> megaseconds

// walker flag bits
constexpr int WF_VALID = 1;        // priors finite
constexpr int WF_SLOWTRIG = 2;     // cadence jitter too large for the series -> direct exp/sincos every step
constexpr int WF_GENERIC = 4;      // has a Q<1/2 term (two real celerite terms) -> generic kernel
constexpr int WF_NOTPD = 8;        // D_n <= 0 or NaN met
constexpr int WF_TR_ALL = 16;      // transit window covers everything

// per-walker prepared parameters: fixed slots in a double[PREP_STRIDE] row
constexpr int PREP_STRIDE = 40;
constexpr int PP_LNPRIOR = 0;
constexpr int PP_ASUM = 1;     // jitter^2 sum + sum of a coefficients (celerite a_sum)
constexpr int PP_JIT2 = 2;
constexpr int PP_TERMS = 4;    // a,b,c,d per term, MAX_TERMS terms -> 4..19 ; real pair: a+,a-,c+,c-
constexpr int PP_TR = 20;      // per,t0,rp,a,inc_rad,ecc,w_rad,u1,u2 -> 20..28
constexpr int PP_TP = 29;      // time of periastron
constexpr int PP_WCEN = 30;    // transit window centre (days)
constexpr int PP_WHALF = 31;   // transit window half-width in phase units (>=0.5: everything)
constexpr int PP_TYPEMASK = 32; // bit k set: term k is a real pair (stored as double)
constexpr int PP_INVPER = 33;   // 1 / per (window test uses a multiply, identically in every kernel)
constexpr int PP_MAXCD = 34;
constexpr int PP_SININC = 35;   // sin(inc)
constexpr int PP_RPINV = 36;    // 1 / |rp|
constexpr int PP_INVOMEGA = 37; // 1 / (1 - u1/3 - u2/6)
constexpr int PP_WOVERPI = 38;  // w / pi    // max |c|, |d| over terms (selects the cadence-jitter series order)

struct ModelDev {
	int ncomp;
	int comp_type[MAX_COMP];
	int comp_off[MAX_COMP];   // offset of the component's first parameter in theta
	int comp_term[MAX_COMP];  // index of its SHO term, -1 for white noise
	int nterms;    // number of SHO terms (granulation + oscillation components)
	int ndim, ndim_gp;
	int has_transit;
	int supersample;
	int ellip_mode;
	int inc_mode;
	int prior_kind[MAX_NDIM];
	double prior_a[MAX_NDIM], prior_b[MAX_NDIM];
	double tr_fixed[9];
	unsigned char tr_free[9];
	double exp_time;
};

// ---------------------------------------------------------------- Philox4x32-10
__host__ __device__ inline void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
					      uint32_t k1, uint32_t out[4])
{
#pragma unroll
	for (int r = 0; r < 10; r++) {
		uint64_t p0 = (uint64_t)0xD2511F53u * c0;
		uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
		uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
		uint32_t n1 = (uint32_t)p1;
		uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
		uint32_t n3 = (uint32_t)p0;
		c0 = n0; c1 = n1; c2 = n2; c3 = n3;
		k0 += 0x9E3779B9u;
		k1 += 0xBB67AE85u;
	}
	out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// stream layout (DESIGN.md "RNG streams"): key = seed; counter = (walker, step lo, step hi | purpose<<28, star)
__host__ __device__ inline void rng_words(uint64_t seed, uint32_t star, uint64_t step, uint32_t walker,
					  uint32_t purpose, uint32_t out[4])
{
	philox4x32_10(walker, (uint32_t)step, (uint32_t)((step >> 32) & 0x0FFFFFFFu) | (purpose << 28), star,
		      (uint32_t)seed, (uint32_t)(seed >> 32), out);
}

__host__ __device__ inline double u53(uint32_t lo, uint32_t hi)
{
	uint64_t v = ((uint64_t)hi << 32) | lo;
	return (double)(v >> 11) * (1.0 / 9007199254740992.0);
}

// ---------------------------------------------------------------- priors
__device__ inline double prior_logpdf(int kind, double a, double b, double x)
{
	if (isnan(x)) return CUDART_NAN;
	if (kind == 0) {
		double scale = b - a;
		double y = (x - a) / scale;
		if (!(scale > 0.0)) return CUDART_NAN;
		if (y >= 0.0 && y <= 1.0) return -log(scale);
		return -CUDART_INF;
	} else if (kind == 1) {
		double y = (x - a) / b;
		if (!(b > 0.0)) return CUDART_NAN;
		return -y * y / 2.0 - 0.91893853320467274178 - log(b);
	} else if (kind == 2) {
		if (!(a > 0.0 && b > a)) return CUDART_NAN;
		if (x >= a && x <= b) return -log(x) - log(log(b) - log(a));
		return -CUDART_INF;
	}
	return CUDART_NAN;
}

// ---------------------------------------------------------------- GP coefficients
// celerite SHOTerm with the reference's log/exp round trip.  Writes (a,b,c,d) of a complex term, or
// (a+,a-,c+,c-) of a real pair (returns true) when Q < 1/2.
__device__ inline bool sho_coeffs(double log_S0, double log_Q, double log_w0, double out[4])
{
	double Q = exp(log_Q), S0 = exp(log_S0), w0 = exp(log_w0);
	if (Q < 0.5) {
		double f = sqrt(1.0 - 4.0 * Q * Q);
		out[0] = 0.5 * S0 * w0 * Q * (1.0 + 1.0 / f);
		out[1] = 0.5 * S0 * w0 * Q * (1.0 - 1.0 / f);
		out[2] = 0.5 * w0 / Q * (1.0 - f);
		out[3] = 0.5 * w0 / Q * (1.0 + f);
		return true;
	}
	double f = sqrt(4.0 * Q * Q - 1.0);
	out[0] = S0 * w0 * Q;
	out[1] = S0 * w0 * Q / f;
	out[2] = 0.5 * w0 / Q;
	out[3] = 0.5 * w0 / Q * f;
	return false;
}

// ---------------------------------------------------------------- transit: elliptic integrals
__device__ inline double ellk_hastings(double k)
{
	double m1 = 1.0 - k * k;
	double ek1 = 1.38629436112 + m1 * (0.09666344259 + m1 * (0.03590092383 + m1 * (0.03742563713 + m1 * 0.01451196212)));
	double ek2 = (0.5 + m1 * (0.12498593597 + m1 * (0.06880248576 + m1 * (0.03328355346 + m1 * 0.00441787012)))) * log(m1);
	return ek1 - ek2;
}

__device__ inline double ellec_hastings(double k)
{
	double m1 = 1.0 - k * k;
	double ee1 = 1.0 + m1 * (0.44325141463 + m1 * (0.06260601220 + m1 * (0.04757383546 + m1 * 0.01736506451)));
	double ee2 = m1 * (0.24998368310 + m1 * (0.09200180037 + m1 * (0.04069697526 + m1 * 0.00526449639))) * log(1.0 / m1);
	return ee1 + ee2;
}

// Bulirsch cel(kc, 1, a, b): complete elliptic integral by the AGM, relative stop tol (error ~ tol^2)
__device__ inline double cel_p1(double kc, double a, double b, double tol)
{
	if (kc == 0.0) return CUDART_INF;
	kc = fabs(kc);
	double e = kc, m = 1.0, p = 1.0, f, g;
	for (int it = 0; it < 64; it++) {
		f = a;
		a = b / p + a;
		g = e / p;
		b = f * g + b;
		b = b + b;
		p = g + p;
		g = m;
		m = kc + m;
		if (fabs(g - kc) <= g * tol) break;
		kc = sqrt(e);
		kc = kc + kc;
		e = kc * m;
	}
	return 0.5 * kPi * (a * m + b) / (m * (m + p));
}

// batman ellpic_bulirsch(n, k)
__device__ inline double ellpic_bulirsch(double n, double k, double tol)
{
	double kc = sqrt(1. - k * k);
	double p = sqrt(n + 1.);
	double m0 = 1., c = 1., d = 1. / p, e = kc, f, g;
	for (int nit = 0; nit < 10000; nit++) {
		f = c;
		c = d / p + c;
		g = e / p;
		d = 2. * (f * g + d);
		p = g + p;
		g = m0;
		m0 = kc + m0;
		if (fabs(1. - kc / g) > tol) {
			kc = 2. * sqrt(e);
			e = kc * m0;
		} else {
			return 0.5 * kPi * (c * m0 + d) / (m0 * (m0 + p));
		}
	}
	return 0.0;
}

__device__ inline double ellk_dev(double k, int mode)
{
	return mode == 0 ? ellk_hastings(k) : cel_p1(sqrt(1.0 - k * k), 1.0, 1.0, 1e-10);
}
__device__ inline double ellec_dev(double k, int mode)
{
	if (mode == 0) return ellec_hastings(k);
	double kc2 = 1.0 - k * k;
	if (kc2 == 0.0) return 1.0;
	return cel_p1(sqrt(kc2), 1.0, kc2, 1e-10);
}
__device__ inline double ellpi_dev(double n, double k, int mode)
{
	return ellpic_bulirsch(n, k, mode == 0 ? 1.0e-8 : 1.0e-10);
}

// ---------------------------------------------------------------- transit: Mandel-Agol flux for one d
__device__ inline double ma_flux(double d, double p, double c1, double c2, int mode)
{
	const double pi = kPi;
	const double omega = 1.0 - c1 / 3.0 - c2 / 6.0;
	const double tol = 1.0e-14;
	double lambdad = 0.0, etad = 0.0, lambdae = 0.0, kap0 = 0.0, kap1 = 0.0;
	double x1, x2, x3, q, Kk, Ek, Pk, nn;
	d = fabs(d);
	if (fabs(p - d) < tol) d = p;
	if (fabs(p - 1.0 - d) < tol) d = p - 1.0;
	if (fabs(1.0 - p - d) < tol) d = 1.0 - p;
	if (d < tol) d = 0.0;

	x1 = (p - d) * (p - d);
	x2 = (p + d) * (p + d);
	x3 = p * p - d * d;

	if (d >= 1.0 + p) return 1.0;
	if (p >= 1.0 && d <= p - 1.0) {
		lambdad = 0.0;
		etad = 0.5;
		lambdae = 1.0;
		return 1.0 - ((1.0 - c1 - 2.0 * c2) * lambdae + (c1 + 2.0 * c2) * (lambdad + 2.0 / 3.0) + c2 * etad) / omega;
	}
	if (d >= fabs(1.0 - p) && d <= 1.0 + p) {
		kap1 = acos(fmin((1.0 - p * p + d * d) / 2.0 / d, 1.0));
		kap0 = acos(fmin((p * p + d * d - 1.0) / 2.0 / p / d, 1.0));
		lambdae = p * p * kap0 + kap1;
		double tt = 1.0 + d * d - p * p;
		lambdae = (lambdae - 0.50 * sqrt(fmax(4.0 * d * d - tt * tt, 0.0))) / pi;
	}
	if (d == p) {
		if (d < 0.5) {
			q = 2.0 * p;
			Kk = ellk_dev(q, mode);
			Ek = ellec_dev(q, mode);
			lambdad = 1.0 / 3.0 + 2.0 / 9.0 / pi * (4.0 * (2.0 * p * p - 1.0) * Ek + (1.0 - 4.0 * p * p) * Kk);
			etad = p * p / 2.0 * (p * p + 2.0 * d * d);
			lambdae = p * p;  // batman leaves it unset on this branch; Mandel-Agol value
		} else if (d > 0.5) {
			q = 0.5 / p;
			Kk = ellk_dev(q, mode);
			Ek = ellec_dev(q, mode);
			double p4 = (p * p) * (p * p);
			lambdad = 1.0 / 3.0 + 16.0 * p / 9.0 / pi * (2.0 * p * p - 1.0) * Ek -
				  (32.0 * p4 - 20.0 * p * p + 3.0) / 9.0 / pi / p * Kk;
			etad = 1.0 / 2.0 / pi * (kap1 + p * p * (p * p + 2.0 * d * d) * kap0 -
					       (1.0 + 5.0 * p * p + d * d) / 4.0 * sqrt((1.0 - x1) * (x2 - 1.0)));
		} else {
			lambdad = 1.0 / 3.0 - 4.0 / pi / 9.0;
			etad = 3.0 / 32.0;
		}
		return 1.0 - ((1.0 - c1 - 2.0 * c2) * lambdae + (c1 + 2.0 * c2) * lambdad + c2 * etad) / omega;
	}
	if ((d > 0.5 + fabs(p - 0.5) && d < 1.0 + p) || (p > 0.5 && d > fabs(1.0 - p) && d < p)) {
		q = sqrt((1.0 - x1) / 4.0 / d / p);
		Kk = ellk_dev(q, mode);
		Ek = ellec_dev(q, mode);
		nn = 1.0 / x1 - 1.0;
		Pk = ellpi_dev(nn, q, mode);
		lambdad = 1.0 / 9.0 / pi / sqrt(p * d) *
			  (((1.0 - x2) * (2.0 * x2 + x1 - 3.0) - 3.0 * x3 * (x2 - 2.0)) * Kk +
			   4.0 * p * d * (d * d + 7.0 * p * p - 4.0) * Ek - 3.0 * x3 / x1 * Pk);
		if (d < p) lambdad += 2.0 / 3.0;
		etad = 1.0 / 2.0 / pi * (kap1 + p * p * (p * p + 2.0 * d * d) * kap0 -
				       (1.0 + 5.0 * p * p + d * d) / 4.0 * sqrt((1.0 - x1) * (x2 - 1.0)));
		return 1.0 - ((1.0 - c1 - 2.0 * c2) * lambdae + (c1 + 2.0 * c2) * lambdad + c2 * etad) / omega;
	}
	if (p <= 1.0 && d <= (1.0 - p)) {
		etad = p * p / 2.0 * (p * p + 2.0 * d * d);
		lambdae = p * p;
		q = sqrt((x2 - x1) / (1.0 - x1));
		Kk = ellk_dev(q, mode);
		Ek = ellec_dev(q, mode);
		nn = x2 / x1 - 1.0;
		Pk = ellpi_dev(nn, q, mode);
		lambdad = 2.0 / 9.0 / pi / sqrt(1.0 - x1) *
			  ((1.0 - 5.0 * d * d + p * p + x3 * x3) * Kk + (1.0 - x1) * (d * d + 7.0 * p * p - 4.0) * Ek -
			   3.0 * x3 / x1 * Pk);
		if (fabs(p + d - 1.0) <= tol) {
			lambdad = 2.0 / 3.0 / pi * acos(1.0 - 2.0 * p) -
				  4.0 / 9.0 / pi * sqrt(p * (1.0 - p)) * (3.0 + 2.0 * p - 8.0 * p * p);
		}
		if (d < p) lambdad += 2.0 / 3.0;
	}
	return 1.0 - ((1.0 - c1 - 2.0 * c2) * lambdae + (c1 + 2.0 * c2) * lambdad + c2 * etad) / omega;
}

// ---------------------------------------------------------------- transit: throughput variant
// Same case analysis and formulas as ma_flux, arranged for the GPU: reciprocals are shared
// (MUFU.RCP64H + Newton) instead of chained IEEE divisions, K and E share one log, the Bulirsch
// iteration does one reciprocal and one square root per pass.  Results agree with ma_flux to a few
// ulp (the tests hold both to 1e-9 against the oracle).
__device__ __forceinline__ double rcp_nr(double d)
{
	double r;
	asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
	double e = fma(-d, r, 1.0);
	r = fma(r, e, r);
	e = fma(-d, r, 1.0);
	return fma(r, e, r);
}

__device__ __forceinline__ double rsqrt_nr(double x)
{
	double y;
	asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
	const double h = 0.5 * x;
	double e = fma(-h * y, y, 0.5);
	y = fma(y, e, y);
	e = fma(-h * y, y, 0.5);
	return fma(y, e, y);
}
// sqrt(x) for finite x without the library routine's slow-path branch: x * rsqrt(x), <= 2 ulp
// (0 -> 0, negative or NaN -> NaN)
__device__ __forceinline__ double sqrt_nr(double x)
{
	const double r = x * rsqrt_nr(x);
	return x == 0.0 ? 0.0 : r;
}

__device__ inline void ellke_fast(double k, int mode, double &Kk, double &Ek, double &m1_out)
{
	m1_out = 1.0 - k * k;
	if (mode == 0) {
		const double m1 = m1_out;
		const double lg = log(m1);
		const double ek1 = 1.38629436112 + m1 * (0.09666344259 + m1 * (0.03590092383 + m1 * (0.03742563713 + m1 * 0.01451196212)));
		const double ek2 = (0.5 + m1 * (0.12498593597 + m1 * (0.06880248576 + m1 * (0.03328355346 + m1 * 0.00441787012)))) * lg;
		const double ee1 = 1.0 + m1 * (0.44325141463 + m1 * (0.06260601220 + m1 * (0.04757383546 + m1 * 0.01736506451)));
		const double ee2 = m1 * (0.24998368310 + m1 * (0.09200180037 + m1 * (0.04069697526 + m1 * 0.00526449639))) * (-lg);
		Kk = ek1 - ek2;
		Ek = ee1 + ee2;
	} else {
		Kk = ellk_dev(k, 1);
		Ek = ellec_dev(k, 1);
	}
}

// Bulirsch iteration of batman's ellpic_bulirsch(n, k) with kc^2 = 1 - k^2, p0 = sqrt(n + 1) and
// 1/p0 supplied by the caller (both are closed forms of the geometry, no square root needed)
__device__ inline double ellpic_fast(double kc2, double p0, double rp0, double tol)
{
	double kc = sqrt_nr(kc2);
	double p = p0;
	double m0 = 1., c = 1., d = rp0, e = kc, f, g;
	for (int nit = 0; nit < 10000; nit++) {
		const double rp = rcp_nr(p);
		f = c;
		c = fma(d, rp, c);
		g = e * rp;
		d = 2. * fma(f, g, d);
		p = g + p;
		g = m0;
		m0 = kc + m0;
		if (fabs(g - kc) > tol * g) {
			kc = 2. * sqrt_nr(e);
			e = kc * m0;
		} else {
			return 0.5 * kPi * fma(c, m0, d) * rcp_nr(m0 * (m0 + p));
		}
	}
	return 0.0;
}


// rp_inv = 1/p, inv_omega = 1/(1 - c1/3 - c2/6): per-walker constants.
// Structure: the case analysis only selects coefficients (q^2, n, prefactor, cK, cE, cP, lambda_e,
// eta_d); K, E, Pi and the combination lambda_d = pref (cK K + cE E - cP Pi) run once, converged,
// for every in-transit lane of the warp.
__device__ inline double ma_flux_fast(double d, double p, double rp_inv, double c1, double c2, double inv_omega, int mode)
{
	const double inv_pi = 0.31830988618379067154;
	const double tol = 1.0e-14;
	d = fabs(d);
	if (fabs(p - d) < tol) d = p;
	if (fabs(p - 1.0 - d) < tol) d = p - 1.0;
	if (fabs(1.0 - p - d) < tol) d = 1.0 - p;
	if (d < tol) d = 0.0;
	const double p2 = p * p, d2 = d * d;
	const double x1 = (p - d) * (p - d), x2 = (p + d) * (p + d), x3 = p2 - d2;
	const double w1 = 1.0 - c1 - 2.0 * c2, w2 = c1 + 2.0 * c2;
	const double pitol = mode == 0 ? 1.0e-8 : 1.0e-10;

	double lambdad = 0.0, etad = 0.0, lambdae = 0.0, kap0 = 0.0, kap1 = 0.0;
	double q2 = 0.0, ep0 = 1.0, erp0 = 1.0, pref = 0.0, cK = 0.0, cE = 0.0, cP = 0.0, add = 0.0;
	bool elliptic = false, done = false;
	double result = 1.0;

	if (d >= 1.0 + p) {
		done = true;  // unocculted
	} else if (p >= 1.0 && d <= p - 1.0) {
		result = 1.0 - (w1 + w2 * (2.0 / 3.0) + c2 * 0.5) * inv_omega;  // fully occulted
		done = true;
	} else {
		if (d >= fabs(1.0 - p) && d <= 1.0 + p) {
			// planet crosses the limb: uniform-source term from the two arcs
			const double rd = rcp_nr(d);
			kap1 = acos(fmin((1.0 - p2 + d2) * 0.5 * rd, 1.0));
			kap0 = acos(fmin((p2 + d2 - 1.0) * 0.5 * rp_inv * rd, 1.0));
			const double tt = 1.0 + d2 - p2;
			lambdae = (p2 * kap0 + kap1 - 0.50 * sqrt(fmax(4.0 * d2 - tt * tt, 0.0))) * inv_pi;
		}
		if (d == p) {
			// planet edge at the stellar centre (measure-zero set): closed forms in K, E only
			double Kk, Ek, m1_;
			if (d < 0.5) {
				ellke_fast(2.0 * p, mode, Kk, Ek, m1_);
				lambdad = 1.0 / 3.0 + (2.0 / 9.0) * inv_pi * (4.0 * (2.0 * p2 - 1.0) * Ek + (1.0 - 4.0 * p2) * Kk);
				etad = p2 * 0.5 * (p2 + 2.0 * d2);
				lambdae = p2;
			} else if (d > 0.5) {
				ellke_fast(0.5 * rp_inv, mode, Kk, Ek, m1_);
				lambdad = 1.0 / 3.0 + (16.0 / 9.0) * p * inv_pi * (2.0 * p2 - 1.0) * Ek -
					  (32.0 * p2 * p2 - 20.0 * p2 + 3.0) * (1.0 / 9.0) * inv_pi * rp_inv * Kk;
				etad = 0.5 * inv_pi * (kap1 + p2 * (p2 + 2.0 * d2) * kap0 - (1.0 + 5.0 * p2 + d2) * 0.25 * sqrt((1.0 - x1) * (x2 - 1.0)));
			} else {
				lambdad = 1.0 / 3.0 - 4.0 * inv_pi / 9.0;
				etad = 3.0 / 32.0;
			}
		} else if ((d > 0.5 + fabs(p - 0.5) && d < 1.0 + p) || (p > 0.5 && d > fabs(1.0 - p) && d < p)) {
			// ingress / egress
			const double rsdp = rsqrt_nr(d * p);
			const double rdp = rsdp * rsdp;
			q2 = (1.0 - x1) * 0.25 * rdp;
			const double rpd = rcp_nr(fabs(p - d));   // n + 1 = 1 / x1  ->  sqrt(n + 1) = 1 / |p - d|
			const double rx1 = rpd * rpd;
			ep0 = rpd;
			erp0 = fabs(p - d);
			pref = (1.0 / 9.0) * inv_pi * rsdp;
			cK = (1.0 - x2) * (2.0 * x2 + x1 - 3.0) - 3.0 * x3 * (x2 - 2.0);
			cE = 4.0 * p * d * (d2 + 7.0 * p2 - 4.0);
			cP = 3.0 * x3 * rx1;
			etad = 0.5 * inv_pi * (kap1 + p2 * (p2 + 2.0 * d2) * kap0 - (1.0 + 5.0 * p2 + d2) * 0.25 * sqrt((1.0 - x1) * (x2 - 1.0)));
			add = (d < p) ? 2.0 / 3.0 : 0.0;
			elliptic = true;
		} else if (p <= 1.0 && d <= (1.0 - p)) {
			// planet fully inside the disc
			etad = p2 * 0.5 * (p2 + 2.0 * d2);
			lambdae = p2;
			const double rs1 = rsqrt_nr(1.0 - x1);
			const double r1 = rs1 * rs1;
			q2 = (x2 - x1) * r1;
			const double rpd = rcp_nr(fabs(p - d));   // n + 1 = x2 / x1  ->  sqrt(n + 1) = (p + d) / |p - d|
			const double rx1 = rpd * rpd;
			ep0 = (p + d) * rpd;
			erp0 = fabs(p - d) * rcp_nr(p + d);
			pref = (2.0 / 9.0) * inv_pi * rs1;
			cK = 1.0 - 5.0 * d2 + p2 + x3 * x3;
			cE = (1.0 - x1) * (d2 + 7.0 * p2 - 4.0);
			cP = 3.0 * x3 * rx1;
			add = (d < p) ? 2.0 / 3.0 : 0.0;
			elliptic = true;
		}
	}
	// converged part: every lane of the warp that needs K, E, Pi computes them here together
	if (elliptic) {
		const double q = sqrt_nr(q2);
		double Kk, Ek, m1;
		ellke_fast(q, mode, Kk, Ek, m1);
		const double Pk = ellpic_fast(m1, ep0, erp0, pitol);
		lambdad = pref * (cK * Kk + cE * Ek - cP * Pk);
		if (p <= 1.0 && d <= (1.0 - p) && fabs(p + d - 1.0) <= tol)  // planet edge touches the limb from inside
			lambdad = (2.0 / 3.0) * inv_pi * acos(1.0 - 2.0 * p) - (4.0 / 9.0) * inv_pi * sqrt(p * (1.0 - p)) * (3.0 + 2.0 * p - 8.0 * p2);
		lambdad += add;
	}
	if (!done) result = 1.0 - (w1 * lambdae + w2 * lambdad + c2 * etad) * inv_omega;
	return result;
}

// ---------------------------------------------------------------- transit: geometry
struct TransitPars {
	double per, t0, rp, a, inc, ecc, w, u1, u2;  // batman semantics, angles in radians
	double tp;                                   // time of periastron (batman _rsky.c)
	double wcen, whalf, inv_per;                 // conservative in-transit window
	double sin_inc, rp_inv, inv_omega, w_over_pi; // per-walker constants of the throughput variant
	bool inverse;                                // rp < 0: batman's inverse transit, light curve 2 - lc
};

__device__ inline double kepler_E(double M, double e)
{
	double E = M;
	const double eps = 1.0e-7;
	int guard = 0;
	while (fmod(fabs(E - e * sin(E) - M), kTwoPi) > eps) {
		double fe = fmod(E - e * sin(E) - M, kTwoPi);
		double fs = fmod(1 - e * cos(E), kTwoPi);
		E = E - fe / fs;
		if (++guard > 200) return CUDART_NAN;
	}
	return E;
}

__device__ inline double periastron_time(double tc, double per, double ecc, double omega)
{
	double fc = kPi / 2. - omega;
	double Ec = 2. * atan(sqrt((1. - ecc) / (1. + ecc)) * tan(fc / 2.));
	return tc - per / 2. / kPi * (Ec - ecc * sin(Ec));
}

// sky-projected separation at time t (primary transit only; 100 when the planet is behind the star)
__device__ inline double rsky_one(double t, const TransitPars &tr)
{
	double f;
	if (tr.ecc < 1.0e-5) {
		double ph = (t - tr.tp) / tr.per;
		f = (ph - trunc(ph)) * 2. * kPi;
	} else {
		double M = (2. * kPi / tr.per) * (t - tr.tp);
		double E = kepler_E(M, tr.ecc);
		f = 2. * atan(sqrt((1. + tr.ecc) / (1. - tr.ecc)) * tan(E / 2.));
	}
	double sf, cf;
	sincos(tr.w + f, &sf, &cf);
	double si = sin(tr.inc);
	double d = tr.a * (1.0 - tr.ecc * tr.ecc) / (1.0 + tr.ecc * cos(f)) * sqrt(1.0 - sf * sf * si * si);
	return (si * sf > 0.) ? d : 100.;
}

// throughput variant of rsky_one: reciprocal period, per-walker sin(inc), no division for e == 0
__device__ inline double rsky_fast(double t, const TransitPars &tr)
{
	double sf, r = tr.a;
	const double si = tr.sin_inc;
	if (tr.ecc < 1.0e-5) {
		const double ph = (t - tr.tp) * tr.inv_per;
		const double fr = ph - trunc(ph);              // f = 2 pi fr
		sf = sinpi(fma(2.0, fr, tr.w_over_pi));         // sin(w + f)
		if (tr.ecc != 0.0) r = tr.a * (1.0 - tr.ecc * tr.ecc) / (1.0 + tr.ecc * cospi(2.0 * fr));
	} else {
		const double M = (2. * kPi * tr.inv_per) * (t - tr.tp);
		const double E = kepler_E(M, tr.ecc);
		const double f = 2. * atan(sqrt((1. + tr.ecc) / (1. - tr.ecc)) * tan(E / 2.));
		sf = sin(tr.w + f);
		r = tr.a * (1.0 - tr.ecc * tr.ecc) / (1.0 + tr.ecc * cos(f));
	}
	const double d = r * sqrt_nr(1.0 - sf * sf * si * si);
	return (si * sf > 0.) ? d : 100.;
}

// Mean anomaly of true anomaly f (closed form)
__device__ inline double mean_anomaly_of_f(double f, double e)
{
	double sf, cf;
	sincos(f, &sf, &cf);
	double E = atan2(sqrt(fmax(1.0 - e * e, 0.0)) * sf, e + cf);
	return E - e * sin(E);
}

__device__ inline double wrap_pi(double x)
{
	x = fmod(x, kTwoPi);
	if (x > kPi) x -= kTwoPi;
	if (x <= -kPi) x += kTwoPi;
	return x;
}

// Conservative in-transit window: every time at which any supersample can have d < 1 + p lies within
// |frac((t - wcen)/per)| <= whalf.  whalf >= 0.5 means "everything".  (Derivation in DESIGN.md.)
__device__ inline void transit_window(TransitPars &tr, double exp_time)
{
	tr.wcen = tr.t0;
	tr.whalf = 1.0;
	tr.inv_per = 1.0 / tr.per;
	tr.sin_inc = sin(tr.inc);
	tr.rp_inv = 1.0 / fabs(tr.rp);
	tr.inv_omega = 1.0 / (1.0 - tr.u1 / 3.0 - tr.u2 / 6.0);
	tr.w_over_pi = tr.w / kPi;
	double p = fabs(tr.rp);
	double rmin = tr.a * (1.0 - tr.ecc);
	if (!(tr.ecc >= 0.0 && tr.ecc < 0.99) || !(rmin > 0.0) || !(p < 90.0) || !(fabs(tr.per) > 0.0) ||
	    !isfinite(tr.per) || !isfinite(tr.t0) || !isfinite(tr.w) || !isfinite(tr.inc))
		return;
	double sw = (1.0 + p) / rmin * (1.0 + 1e-9);
	double si = sin(tr.inc), ci = cos(tr.inc);
	// d = r sqrt(sin^2 psi + cos^2 psi cos^2 i) with psi the angle from conjunction and r >= rmin:
	// d < 1 + p needs sin^2 psi < (sw^2 - cos^2 i) / sin^2 i
	double num = sw * sw - ci * ci, den = si * si;
	if (!(den > 0.0)) return;
	if (num <= 0.0) {
		tr.whalf = -1.0;  // never closer than 1 + p: no sample is in transit
		return;
	}
	double s2 = num / den * (1.0 + 1e-9);
	if (!(s2 < 1.0)) return;
	double df = asin(sqrt(s2));
	double fmid = (si >= 0.0 ? 0.5 * kPi : 1.5 * kPi) - tr.w;
	double Mm = mean_anomaly_of_f(fmid, tr.ecc);
	double M1 = mean_anomaly_of_f(fmid - df, tr.ecc);
	double M2 = mean_anomaly_of_f(fmid + df, tr.ecc);
	double d1 = fabs(wrap_pi(Mm - M1)), d2 = fabs(wrap_pi(M2 - Mm));
	double half = fmax(d1, d2) + 1.0e-4;  // pad: e<1e-5 branch treats f = M; Kepler stop 1e-7
	tr.wcen = tr.tp + tr.per * (Mm / kTwoPi);
	tr.whalf = half / kTwoPi + fabs(0.5 * exp_time / tr.per) + 1e-9;
}

__device__ __forceinline__ bool in_transit_window(double t, double wcen, double inv_per, double whalf)
{
	// |ph| <= 0.5 after the rint, so whalf >= 0.5 means "always"; NaN -> inside (compute, propagate NaN)
	double ph = (t - wcen) * inv_per;
	ph -= rint(ph);
	return !(fabs(ph) > whalf);
}
__device__ __forceinline__ bool in_transit_window(double t, const TransitPars &tr)
{
	return in_transit_window(t, tr.wcen, tr.inv_per, tr.whalf);
}

// np.linspace(-e/2, e/2, s)[i] from start = -e/2, step = e/(s-1) (the last point is exactly +e/2)
__device__ __forceinline__ double supersample_offset(int i, int ss, double start, double step)
{
	return (i == ss - 1) ? -start : fma((double)i, step, start);
}

}  // namespace gpt
