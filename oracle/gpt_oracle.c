/*
 * gpt_oracle.c -- CPU restatement of the gptransits posterior hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under gptransits_b200/ may import, link or
 * call this file; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, as the checker (never as the
 * thing shipped).
 *
 * PARITY STATUS: "parity unpinned".  The reference (Fill4/gptransits) ships no
 * tests or golden vectors, and the arithmetic lives in three un-vendored,
 * un-pinned PyPI dependencies that are absent from /root/reference and from
 * this image (setup.py:35-47: celerite, batman-package, emcee).  This file
 * restates their published algorithms:
 *   - celerite  (>=0.3, Foreman-Mackey et al. 2017, AJ 154:220, Sec. 5;
 *     CholeskySolver::compute / dot_solve / log_determinant)
 *   - batman-package (2.4.x; Kreidberg 2015; _rsky.c, _quadratic_ld.c incl. the
 *     Hastings K/E polynomials and the Bulirsch Pi iteration)
 *   - emcee (3.x; RedBlueMove + StretchMove) with the RNG replaced by a
 *     counter-based Philox4x32-10 stream (documented in DESIGN.md).
 * and follows the reference's own call sites for everything around them:
 *   gptransits/component.py:33-51,83-101,128-140,168-172
 *   gptransits/gp.py:34-40,51-57
 *   gptransits/transit.py:56-64,88-106,120-127
 *   gptransits/model.py:164,183-187,299-325
 * It is pinned against (tests/test_oracle_*.py): scipy.stats log-pdfs, a dense
 * Cholesky factorisation, scipy.special elliptic integrals + brute-force
 * quadrature of the limb-darkened disc, the Random123 Philox known-answer
 * vectors, golden vectors produced by running the reference's own
 * component.py/gp.py/transit.py (tests/golden/make_golden.py), and the
 * independently derived numbers recorded in SURVEY.md section 8(c).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

#define ORC_MAX_TERMS 16 /* max SHO terms */
#define ORC_MAXJ (2 * ORC_MAX_TERMS)

/* ------------------------------------------------------------------------- */
/* P1: priors.  scipy.stats frozen distributions as built at                  */
/* component.py:33-41 / transit.py:56-64, evaluated by logpdf and summed       */
/* (component.py:50-51, transit.py:120-122).                                   */
/* kind: 0 uniform(loc=a, scale=b-a); 1 norm(loc=a, scale=b); 2 reciprocal(a,b)*/
/* ------------------------------------------------------------------------- */
double orc_logpdf(int kind, double a, double b, double x)
{
	if (isnan(x)) return NAN;
	if (kind == 0) {
		double scale = b - a;
		double y = (x - a) / scale;
		if (!(scale > 0.0)) return NAN;
		if (y >= 0.0 && y <= 1.0) return -log(scale);
		return -INFINITY;
	} else if (kind == 1) {
		double y = (x - a) / b;
		if (!(b > 0.0)) return NAN;
		/* log(sqrt(2 pi)) */
		return -y * y / 2.0 - 0.91893853320467274178 - log(b);
	} else if (kind == 2) {
		if (!(a > 0.0 && b > a)) return NAN;
		if (x >= a && x <= b) return -log(x) - log(log(b) - log(a));
		return -INFINITY;
	}
	return NAN;
}

double orc_lnprior(const int *kind, const double *pa, const double *pb, const double *x, int n)
{
	double s = 0.0;
	for (int i = 0; i < n; i++) s += orc_logpdf(kind[i], pa[i], pb[i], x[i]);
	return s;
}

/* ------------------------------------------------------------------------- */
/* P2: physical -> celerite parameters -> (a,b,c,d) coefficients.             */
/* comp type: 0 granulation (component.py:83-109), 1 oscillation bump          */
/* (component.py:128-148), 2 white noise (component.py:168-179).               */
/* The reference hands celerite the LOG of the parameters and celerite         */
/* exponentiates them again (celerite terms.SHOTerm / JitterTerm); that        */
/* round trip is kept.  celerite's TermSum lists all real terms first, then    */
/* all complex terms, each in component order.                                 */
/* ------------------------------------------------------------------------- */
typedef struct {
	double jitter; /* sum of exp(2 log_sigma) over JitterTerms */
	int n_real;
	double a_real[ORC_MAXJ], c_real[ORC_MAXJ];
	int n_comp;
	double a_comp[ORC_MAX_TERMS], b_comp[ORC_MAX_TERMS], c_comp[ORC_MAX_TERMS], d_comp[ORC_MAX_TERMS];
} orc_coeffs;

static void sho_coeffs(double log_S0, double log_Q, double log_w0, orc_coeffs *k)
{
	double Q = exp(log_Q), S0 = exp(log_S0), w0 = exp(log_w0);
	if (Q < 0.5) {
		double f = sqrt(1.0 - 4.0 * Q * Q);
		int i = k->n_real;
		k->a_real[i] = 0.5 * S0 * w0 * Q * (1.0 + 1.0 / f);
		k->c_real[i] = 0.5 * w0 / Q * (1.0 - f);
		k->a_real[i + 1] = 0.5 * S0 * w0 * Q * (1.0 - 1.0 / f);
		k->c_real[i + 1] = 0.5 * w0 / Q * (1.0 + f);
		k->n_real += 2;
	} else {
		double f = sqrt(4.0 * Q * Q - 1.0);
		int i = k->n_comp;
		k->a_comp[i] = S0 * w0 * Q;
		k->b_comp[i] = S0 * w0 * Q / f;
		k->c_comp[i] = 0.5 * w0 / Q;
		k->d_comp[i] = 0.5 * w0 / Q * f;
		k->n_comp += 1;
	}
}

static int comp_npars(int type) { return type == 0 ? 2 : (type == 1 ? 3 : 1); }

/* celerite log-parameter vector, gp.py:34-40 */
int orc_gp_celerite_params(const int *comp_type, int ncomp, const double *theta, double *out)
{
	int i = 0, o = 0;
	for (int c = 0; c < ncomp; c++) {
		if (comp_type[c] == 0) {
			double a = theta[i], b = theta[i + 1];
			double w = b * (2 * M_PI);
			double S = (sqrt(2.0) / (2 * M_PI)) * (a * a / b);
			out[o++] = log(S);
			out[o++] = log(w);
		} else if (comp_type[c] == 1) {
			double Pg = theta[i], Q = theta[i + 1], numax = theta[i + 2];
			double S = Pg / (4 * (Q * Q));
			double w = numax * (2 * M_PI);
			out[o++] = log(S);
			out[o++] = log(Q);
			out[o++] = log(w);
		} else {
			out[o++] = log(theta[i]);
		}
		i += comp_npars(comp_type[c]);
	}
	return o;
}

static void gp_coeffs(const int *comp_type, int ncomp, const double *theta, orc_coeffs *k)
{
	double lp[4 * ORC_MAX_TERMS];
	memset(k, 0, sizeof(*k));
	orc_gp_celerite_params(comp_type, ncomp, theta, lp);
	int o = 0;
	for (int c = 0; c < ncomp; c++) {
		if (comp_type[c] == 0) {
			/* Q = 1/sqrt(2) frozen, component.py:106-108 */
			sho_coeffs(lp[o], log(1.0 / sqrt(2.0)), lp[o + 1], k);
			o += 2;
		} else if (comp_type[c] == 1) {
			sho_coeffs(lp[o], lp[o + 1], lp[o + 2], k);
			o += 3;
		} else {
			k->jitter += exp(2.0 * lp[o]);
			o += 1;
		}
	}
}

/* exported flat view: out = [jitter, n_real, n_comp, a_real.., c_real.., a.., b.., c.., d..] */
int orc_gp_coeffs(const int *comp_type, int ncomp, const double *theta, double *out)
{
	orc_coeffs k;
	gp_coeffs(comp_type, ncomp, theta, &k);
	int o = 0;
	out[o++] = k.jitter;
	out[o++] = k.n_real;
	out[o++] = k.n_comp;
	for (int i = 0; i < k.n_real; i++) out[o++] = k.a_real[i];
	for (int i = 0; i < k.n_real; i++) out[o++] = k.c_real[i];
	for (int i = 0; i < k.n_comp; i++) out[o++] = k.a_comp[i];
	for (int i = 0; i < k.n_comp; i++) out[o++] = k.b_comp[i];
	for (int i = 0; i < k.n_comp; i++) out[o++] = k.c_comp[i];
	for (int i = 0; i < k.n_comp; i++) out[o++] = k.d_comp[i];
	return o;
}

/* ------------------------------------------------------------------------- */
/* G1 + G2: celerite CholeskySolver.  compute() builds U, X(=W), phi, D        */
/* column by column; dot_solve() is the forward substitution returning         */
/* sum z^2/D; log_determinant() is sum log D.  Called as at model.py:183-187   */
/* (compute) and model.py:310,323 (log_likelihood).                            */
/* Workspace layout: D[N], X[J*N], U[J*N], phi[J*(N-1)] (column n contiguous).  */
/* Returns 0, or 1 when the matrix is not positive definite (celerite raises   */
/* LinAlgError when D_n < 0).                                                  */
/* ------------------------------------------------------------------------- */
static int celerite_compute(const orc_coeffs *k, int N, const double *x, const double *diag,
			    double *D, double *X, double *U, double *phi)
{
	const int Jr = k->n_real, Jc = k->n_comp, J = Jr + 2 * Jc;
	double S[ORC_MAXJ * ORC_MAXJ];
	double a_sum = k->jitter;
	for (int j = 0; j < Jr; j++) a_sum += k->a_real[j];
	for (int j = 0; j < Jc; j++) a_sum += k->a_comp[j];

	if (N <= 0) return 0;
	D[0] = diag[0] + a_sum;
	double value = 1.0 / D[0];
	for (int j = 0; j < Jr; j++) {
		U[j] = k->a_real[j];
		X[j] = value;
	}
	for (int j = 0; j < Jc; j++) {
		double cd = cos(k->d_comp[j] * x[0]), sd = sin(k->d_comp[j] * x[0]);
		U[Jr + 2 * j] = k->a_comp[j] * cd + k->b_comp[j] * sd;
		U[Jr + 2 * j + 1] = k->a_comp[j] * sd - k->b_comp[j] * cd;
		X[Jr + 2 * j] = cd * value;
		X[Jr + 2 * j + 1] = sd * value;
	}
	memset(S, 0, sizeof(S));
	if (D[0] < 0.0) return 1;

	for (int n = 1; n < N; n++) {
		double t = x[n], dx = t - x[n - 1];
		double *Un = U + (size_t)J * n, *Xn = X + (size_t)J * n;
		double *Xp = X + (size_t)J * (n - 1), *ph = phi + (size_t)J * (n - 1);
		for (int j = 0; j < Jr; j++) {
			ph[j] = exp(-k->c_real[j] * dx);
			Un[j] = k->a_real[j];
			Xn[j] = 1.0;
		}
		for (int j = 0; j < Jc; j++) {
			double v = exp(-k->c_comp[j] * dx);
			ph[Jr + 2 * j] = v;
			ph[Jr + 2 * j + 1] = v;
			double cd = cos(k->d_comp[j] * t), sd = sin(k->d_comp[j] * t);
			Un[Jr + 2 * j] = k->a_comp[j] * cd + k->b_comp[j] * sd;
			Un[Jr + 2 * j + 1] = k->a_comp[j] * sd - k->b_comp[j] * cd;
			Xn[Jr + 2 * j] = cd;
			Xn[Jr + 2 * j + 1] = sd;
		}
		/* S <- phi phi^T o (S + D_{n-1} X_{n-1} X_{n-1}^T), upper triangle */
		double Dn = D[n - 1];
		for (int j = 0; j < J; j++) {
			double phij = ph[j], xj = Dn * Xp[j];
			for (int kk = 0; kk <= j; kk++)
				S[kk * J + j] = phij * (ph[kk] * (S[kk * J + j] + xj * Xp[kk]));
		}
		/* D_n = A_n - U S U^T ; X_n = (V_n - U_n S) / D_n */
		Dn = diag[n] + a_sum;
		for (int j = 0; j < J; j++) {
			double uj = Un[j], xj = Xn[j];
			for (int kk = 0; kk < j; kk++) {
				double tmp = Un[kk] * S[kk * J + j];
				Dn -= 2.0 * (uj * tmp);
				xj -= tmp;
				Xn[kk] -= uj * S[kk * J + j];
			}
			double tmp = uj * S[j * J + j];
			Dn -= uj * tmp;
			Xn[j] = xj - tmp;
		}
		if (Dn < 0.0) return 1;
		D[n] = Dn;
		for (int j = 0; j < J; j++) Xn[j] /= Dn;
	}
	return 0;
}

static double celerite_dot_solve(int J, int N, const double *D, const double *X, const double *U,
				 const double *phi, const double *b)
{
	double F[ORC_MAXJ];
	for (int j = 0; j < J; j++) F[j] = 0.0;
	if (N <= 0) return 0.0;
	double z = b[0], T = z * z / D[0];
	for (int n = 1; n < N; n++) {
		const double *Xp = X + (size_t)J * (n - 1), *ph = phi + (size_t)J * (n - 1);
		const double *Un = U + (size_t)J * n;
		double acc = 0.0;
		for (int j = 0; j < J; j++) {
			F[j] = ph[j] * (F[j] + Xp[j] * z);
			acc += Un[j] * F[j];
		}
		z = b[n] - acc;
		T += z * z / D[n];
	}
	return T;
}

/* quad, logdet for one parameter vector; status 1 = not positive definite */
int orc_celerite_loglike(const int *comp_type, int ncomp, const double *theta_gp, int N,
			 const double *x, const double *diag, const double *resid, double *quad,
			 double *logdet)
{
	orc_coeffs k;
	gp_coeffs(comp_type, ncomp, theta_gp, &k);
	int J = k.n_real + 2 * k.n_comp;
	int Jm = J > 0 ? J : 1;
	double *D = malloc(sizeof(double) * (size_t)N);
	double *X = malloc(sizeof(double) * (size_t)N * Jm);
	double *U = malloc(sizeof(double) * (size_t)N * Jm);
	double *phi = malloc(sizeof(double) * (size_t)N * Jm);
	int st = celerite_compute(&k, N, x, diag, D, X, U, phi);
	if (st == 0) {
		double ld = 0.0;
		for (int n = 0; n < N; n++) ld += log(D[n]);
		*logdet = ld;
		*quad = celerite_dot_solve(J, N, D, X, U, phi, resid);
	} else {
		*logdet = NAN;
		*quad = NAN;
	}
	free(D);
	free(X);
	free(U);
	free(phi);
	return st;
}

/* celerite factor outputs for inspection by tests (D only) */
int orc_celerite_factor_D(const int *comp_type, int ncomp, const double *theta_gp, int N,
			  const double *x, const double *diag, double *D)
{
	orc_coeffs k;
	gp_coeffs(comp_type, ncomp, theta_gp, &k);
	int J = k.n_real + 2 * k.n_comp;
	int Jm = J > 0 ? J : 1;
	double *X = malloc(sizeof(double) * (size_t)N * Jm);
	double *U = malloc(sizeof(double) * (size_t)N * Jm);
	double *phi = malloc(sizeof(double) * (size_t)N * Jm);
	int st = celerite_compute(&k, N, x, diag, D, X, U, phi);
	free(X);
	free(U);
	free(phi);
	return st;
}

/* ------------------------------------------------------------------------- */
/* T2: batman _rsky.c -- sky-projected separation, primary transit only.      */
/* inc and w in RADIANS here (TransitModel converts degrees before the call).  */
/* ------------------------------------------------------------------------- */
static double getE(double M, double e)
{
	double E = M, eps = 1.0e-7;
	int guard = 0;
	while (fmod(fabs(E - e * sin(E) - M), 2. * M_PI) > eps) {
		double fe = fmod(E - e * sin(E) - M, 2. * M_PI);
		double fs = fmod(1 - e * cos(E), 2. * M_PI);
		E = E - fe / fs;
		if (++guard > 10000) return NAN; /* batman would spin; e>=1 or NaN input */
	}
	return E;
}

void orc_rsky(const double *t, int n, double tc, double per, double a, double inc, double ecc,
	      double omega, double *out)
{
	const double BIGD = 100.;
	const double nm = 2. * M_PI / per;
	for (int i = 0; i < n; i++) {
		double fc = M_PI / 2. - omega;
		double Ec = 2. * atan(sqrt((1. - ecc) / (1. + ecc)) * tan(fc / 2.));
		double tp = tc - per / 2. / M_PI * (Ec - ecc * sin(Ec));
		double f;
		if (ecc < 1.0e-5) {
			double ph = (t[i] - tp) / per;
			f = (ph - (int)ph) * 2. * M_PI;
		} else {
			double M = nm * (t[i] - tp);
			double E = getE(M, ecc);
			f = 2. * atan(sqrt((1. + ecc) / (1. - ecc)) * tan(E / 2.));
		}
		double d = a * (1.0 - ecc * ecc) / (1.0 + ecc * cos(f)) *
			   sqrt(1.0 - sin(omega + f) * sin(omega + f) * sin(inc) * sin(inc));
		if (sin(inc) * sin(omega + f) > 0.)
			out[i] = d;
		else
			out[i] = BIGD;
	}
}

/* ------------------------------------------------------------------------- */
/* T3: batman _quadratic_ld.c -- Mandel & Agol (2002) quadratic limb           */
/* darkening.  mode 0: batman's numerics (Hastings polynomial K and E,         */
/* Bulirsch iteration for Pi stopped at 1e-8).  mode 1: same case analysis     */
/* with K, E, Pi iterated to machine precision (AGM / Bulirsch cel).           */
/* ------------------------------------------------------------------------- */
static double ellk_hastings(double k)
{
	double m1 = 1.0 - k * k;
	double a0 = 1.38629436112, a1 = 0.09666344259, a2 = 0.03590092383, a3 = 0.03742563713,
	       a4 = 0.01451196212;
	double b0 = 0.5, b1 = 0.12498593597, b2 = 0.06880248576, b3 = 0.03328355346,
	       b4 = 0.00441787012;
	double ek1 = a0 + m1 * (a1 + m1 * (a2 + m1 * (a3 + m1 * a4)));
	double ek2 = (b0 + m1 * (b1 + m1 * (b2 + m1 * (b3 + m1 * b4)))) * log(m1);
	return ek1 - ek2;
}

static double ellec_hastings(double k)
{
	double m1 = 1.0 - k * k;
	double a1 = 0.44325141463, a2 = 0.06260601220, a3 = 0.04757383546, a4 = 0.01736506451;
	double b1 = 0.24998368310, b2 = 0.09200180037, b3 = 0.04069697526, b4 = 0.00526449639;
	double ee1 = 1.0 + m1 * (a1 + m1 * (a2 + m1 * (a3 + m1 * a4)));
	double ee2 = m1 * (b1 + m1 * (b2 + m1 * (b3 + m1 * b4))) * log(1.0 / m1);
	return ee1 + ee2;
}

/* Bulirsch general complete elliptic integral cel(kc, p, a, b).  The AGM converges
 * quadratically, so stopping at relative tol leaves an error of order tol^2. */
static double cel_bulirsch(double kc, double p, double a, double b, double tol)
{
	if (kc == 0.0) return INFINITY;
	kc = fabs(kc);
	double e = kc, m = 1.0, f, g, q;
	if (p > 0.0) {
		p = sqrt(p);
		b = b / p;
	} else {
		f = kc * kc;
		q = 1.0 - f;
		g = 1.0 - p;
		f = f - p;
		q = q * (b - a * p);
		p = sqrt(f / g);
		a = (a - b) / g;
		b = -q / (g * g * p) + a * p;
	}
	for (int it = 0; it < 10000; it++) {
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
	return 0.5 * M_PI * (a * m + b) / (m * (m + p));
}

/* batman's ellpic_bulirsch(n, k): Pi(-n... ) in M&A's convention, 1e-8 stop */
static double ellpic_bulirsch(double n, double k, double tol)
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
			return 0.5 * M_PI * (c * m0 + d) / (m0 * (m0 + p));
		}
	}
	return 0;
}

static double ellk(double k, int mode)
{
	if (mode == 0) return ellk_hastings(k);
	return cel_bulirsch(sqrt(1.0 - k * k), 1.0, 1.0, 1.0, 1e-10);
}

static double ellec(double k, int mode)
{
	if (mode == 0) return ellec_hastings(k);
	double kc2 = 1.0 - k * k;
	if (kc2 == 0.0) return 1.0;
	return cel_bulirsch(sqrt(kc2), 1.0, 1.0, kc2, 1e-10);
}

static double ellpi(double n, double k, int mode)
{
	return ellpic_bulirsch(n, k, mode == 0 ? 1.0e-8 : 1.0e-10);
}

#define ORC_MIN(x, y) (((x) < (y)) ? (x) : (y))
#define ORC_MAX(x, y) (((x) > (y)) ? (x) : (y))

void orc_quadratic_ld(const double *ds, int n, double p, double c1, double c2, int mode, double *flux)
{
	const double pi = M_PI;
	const double omega = 1.0 - c1 / 3.0 - c2 / 6.0;
	const double tol = 1.0e-14;
	for (int i = 0; i < n; i++) {
		double lambdad = 0.0, etad = 0.0, lambdae = 0.0, kap0 = 0.0, kap1 = 0.0;
		double x1, x2, x3, q, Kk, Ek, Pk, nn;
		double d = fabs(ds[i]);
		if (fabs(p - d) < tol) d = p;
		if (fabs(p - 1.0 - d) < tol) d = p - 1.0;
		if (fabs(1.0 - p - d) < tol) d = 1.0 - p;
		if (d < tol) d = 0.0;

		x1 = pow((p - d), 2.0);
		x2 = pow((p + d), 2.0);
		x3 = p * p - d * d;

		/* source is unocculted */
		if (d >= 1.0 + p) {
			flux[i] = 1.0;
			continue;
		}
		/* source is completely occulted */
		if (p >= 1.0 && d <= p - 1.0) {
			lambdad = 0.0;
			etad = 0.5;
			lambdae = 1.0;
			flux[i] = 1.0 - ((1.0 - c1 - 2.0 * c2) * lambdae + (c1 + 2.0 * c2) * (lambdad + 2.0 / 3.0) +
					 c2 * etad) / omega;
			continue;
		}
		/* source is partly occulted and occulting object crosses the limb */
		if (d >= fabs(1.0 - p) && d <= 1.0 + p) {
			kap1 = acos(ORC_MIN((1.0 - p * p + d * d) / 2.0 / d, 1.0));
			kap0 = acos(ORC_MIN((p * p + d * d - 1.0) / 2.0 / p / d, 1.0));
			lambdae = p * p * kap0 + kap1;
			lambdae = (lambdae - 0.50 * sqrt(ORC_MAX(4.0 * d * d - pow((1.0 + d * d - p * p), 2.0), 0.0))) / pi;
		}
		/* edge of the occulting star lies at the origin */
		if (d == p) {
			if (d < 0.5) {
				q = 2.0 * p;
				Kk = ellk(q, mode);
				Ek = ellec(q, mode);
				lambdad = 1.0 / 3.0 + 2.0 / 9.0 / pi * (4.0 * (2.0 * p * p - 1.0) * Ek + (1.0 - 4.0 * p * p) * Kk);
				etad = p * p / 2.0 * (p * p + 2.0 * d * d);
				/* batman reads lambdae without having set it on this branch (planet edge at the
				 * stellar centre, fully inside the disc); the Mandel-Agol value is p^2.  Measure-zero
				 * set (|d - p| < 1e-14); recorded in DESIGN.md. */
				lambdae = p * p;
				flux[i] = 1.0 - ((1.0 - c1 - 2.0 * c2) * lambdae + (c1 + 2.0 * c2) * lambdad + c2 * etad) / omega;
				continue;
			} else if (d > 0.5) {
				q = 0.5 / p;
				Kk = ellk(q, mode);
				Ek = ellec(q, mode);
				lambdad = 1.0 / 3.0 + 16.0 * p / 9.0 / pi * (2.0 * p * p - 1.0) * Ek -
					  (32.0 * pow(p, 4.0) - 20.0 * p * p + 3.0) / 9.0 / pi / p * Kk;
				etad = 1.0 / 2.0 / pi * (kap1 + p * p * (p * p + 2.0 * d * d) * kap0 -
						       (1.0 + 5.0 * p * p + d * d) / 4.0 * sqrt((1.0 - x1) * (x2 - 1.0)));
			} else {
				lambdad = 1.0 / 3.0 - 4.0 / pi / 9.0;
				etad = 3.0 / 32.0;
				flux[i] = 1.0 - ((1.0 - c1 - 2.0 * c2) * lambdae + (c1 + 2.0 * c2) * lambdad + c2 * etad) / omega;
				continue;
			}
			flux[i] = 1.0 - ((1.0 - c1 - 2.0 * c2) * lambdae + (c1 + 2.0 * c2) * lambdad + c2 * etad) / omega;
			continue;
		}
		/* occulting star partly occults the source and crosses the limb */
		if ((d > 0.5 + fabs(p - 0.5) && d < 1.0 + p) || (p > 0.5 && d > fabs(1.0 - p) && d < p)) {
			q = sqrt((1.0 - x1) / 4.0 / d / p);
			Kk = ellk(q, mode);
			Ek = ellec(q, mode);
			nn = 1.0 / x1 - 1.0;
			Pk = ellpi(nn, q, mode);
			lambdad = 1.0 / 9.0 / pi / sqrt(p * d) *
				  (((1.0 - x2) * (2.0 * x2 + x1 - 3.0) - 3.0 * x3 * (x2 - 2.0)) * Kk +
				   4.0 * p * d * (d * d + 7.0 * p * p - 4.0) * Ek - 3.0 * x3 / x1 * Pk);
			if (d < p) lambdad += 2.0 / 3.0;
			etad = 1.0 / 2.0 / pi * (kap1 + p * p * (p * p + 2.0 * d * d) * kap0 -
					       (1.0 + 5.0 * p * p + d * d) / 4.0 * sqrt((1.0 - x1) * (x2 - 1.0)));
			flux[i] = 1.0 - ((1.0 - c1 - 2.0 * c2) * lambdae + (c1 + 2.0 * c2) * lambdad + c2 * etad) / omega;
			continue;
		}
		/* occulting star transits the source */
		if (p <= 1.0 && d <= (1.0 - p)) {
			etad = p * p / 2.0 * (p * p + 2.0 * d * d);
			lambdae = p * p;
			q = sqrt((x2 - x1) / (1.0 - x1));
			Kk = ellk(q, mode);
			Ek = ellec(q, mode);
			nn = x2 / x1 - 1.0;
			Pk = ellpi(nn, q, mode);
			lambdad = 2.0 / 9.0 / pi / sqrt(1.0 - x1) *
				  ((1.0 - 5.0 * d * d + p * p + x3 * x3) * Kk + (1.0 - x1) * (d * d + 7.0 * p * p - 4.0) * Ek -
				   3.0 * x3 / x1 * Pk);
			/* edge of planet hits edge of star */
			if (fabs(p + d - 1.0) <= tol) {
				lambdad = 2.0 / 3.0 / pi * acos(1.0 - 2.0 * p) -
					  4.0 / 9.0 / pi * sqrt(p * (1.0 - p)) * (3.0 + 2.0 * p - 8.0 * p * p);
			}
			if (d < p) lambdad += 2.0 / 3.0;
		}
		flux[i] = 1.0 - ((1.0 - c1 - 2.0 * c2) * lambdae + (c1 + 2.0 * c2) * lambdad + c2 * etad) / omega;
	}
}

/* ------------------------------------------------------------------------- */
/* batman.TransitModel(params, t, supersample_factor=s, exp_time=e)            */
/* .light_curve(params) as driven from transit.py:88-106,124-127.              */
/* pars9 = (per, t0, rp, a, inc[deg], ecc, w[deg], u1, u2) in batman terms.     */
/* out_flux[N] = mean over the s supersamples.                                 */
/* ------------------------------------------------------------------------- */
void orc_light_curve(const double *t, int N, const double *pars9, int ss, double exp_time, int mode,
		     double *out_flux)
{
	double per = pars9[0], t0 = pars9[1], rp = pars9[2], a = pars9[3];
	double inc = pars9[4] * M_PI / 180., ecc = pars9[5], w = pars9[6] * M_PI / 180.;
	double u1 = pars9[7], u2 = pars9[8];
	if (ss < 1) ss = 1;
	double *tss = calloc((size_t)N * ss, sizeof(double));
	double *ds = malloc(sizeof(double) * (size_t)N * ss);
	double *fl = malloc(sizeof(double) * (size_t)N * ss);
	for (int n = 0; n < N; n++) {
		if (ss > 1) {
			/* np.linspace(-e/2, e/2, s) + t */
			double start = -exp_time / 2., stop = exp_time / 2.;
			double step = (stop - start) / (ss - 1);
			for (int i = 0; i < ss; i++) {
				double off = (i == ss - 1) ? stop : start + i * step;
				tss[(size_t)n * ss + i] = off + t[n];
			}
		} else {
			tss[n] = t[n];
		}
	}
	orc_rsky(tss, N * ss, t0, per, a, inc, ecc, w, ds);
	orc_quadratic_ld(ds, N * ss, fabs(rp), u1, u2, mode, fl);
	for (int n = 0; n < N; n++) {
		if (ss > 1) {
			double s = 0.0;
			for (int i = 0; i < ss; i++) s += fl[(size_t)n * ss + i];
			out_flux[n] = s / ss;
		} else {
			out_flux[n] = fl[n];
		}
		/* batman (>= 2.4.6, transitmodel.py light_curve) [recalled]: rp < 0 is an "inverse transit":
		 * the model is computed for |rp| and returned as 2 - lc */
		if (rp < 0.0) out_flux[n] = 2.0 - out_flux[n];
	}
	free(tss);
	free(ds);
	free(fl);
}

/* ------------------------------------------------------------------------- */
/* L1: full posterior, model.py:299-325 (+ transit-only chi^2 per SURVEY A9). */
/* ------------------------------------------------------------------------- */
typedef struct {
	int ncomp;
	const int *comp_type;
	int ndim_gp;
	int has_transit;
	const double *tr_fixed;        /* [9] */
	const unsigned char *tr_free;  /* [9] */
	int ndim; /* total */
	const int *prior_kind;
	const double *prior_a, *prior_b;
	int supersample;
	double exp_time;
	int ellip_mode;
	int inc_mode; /* 0: transit.py:103 verbatim; 1: inc = degrees(arccos(cosi)) */
	int N;
	const double *t_days, *y, *yerr; /* yerr already the value celerite gets */
} orc_model;

static int model_ndim_gp(const int *comp_type, int ncomp)
{
	int n = 0;
	for (int c = 0; c < ncomp; c++) n += comp_npars(comp_type[c]);
	return n;
}

/* parts = [lnprior, quad, logdet, lnL] */
static double log_prob_one(const orc_model *m, const double *theta, double *parts, double *x_ms,
			   double *diag, double *resid, double *flux)
{
	double lnprior = orc_lnprior(m->prior_kind, m->prior_a, m->prior_b, theta, m->ndim);
	double quad = NAN, logdet = NAN, lnl = NAN, lp;
	/* model.py:302-303 / 317-318: both partial sums finite, else -inf */
	double lp_gp = orc_lnprior(m->prior_kind, m->prior_a, m->prior_b, theta, m->ndim_gp);
	double lp_tr = orc_lnprior(m->prior_kind + m->ndim_gp, m->prior_a + m->ndim_gp, m->prior_b + m->ndim_gp,
				   theta + m->ndim_gp, m->ndim - m->ndim_gp);
	if (!(isfinite(lp_gp) && isfinite(lp_tr))) {
		lp = -INFINITY;
		goto done;
	}
	lnprior = lp_tr + lp_gp;
	const double *r = m->y;
	if (m->has_transit) {
		double pars[9];
		int k = m->ndim_gp;
		for (int i = 0; i < 9; i++) pars[i] = m->tr_free[i] ? theta[k++] : m->tr_fixed[i];
		/* transit.py:99-106: per,t0,rp,a,inc,ecc,w,u1,u2 <- pars[0..8] verbatim */
		if (m->inc_mode == 1) pars[4] = acos(pars[4]) * (180. / M_PI);
		orc_light_curve(m->t_days, m->N, pars, m->supersample, m->exp_time, m->ellip_mode, flux);
		for (int n = 0; n < m->N; n++) resid[n] = m->y[n] - (flux[n] - 1) * 1e6;
		r = resid;
	}
	if (m->ncomp > 0) {
		int st = orc_celerite_loglike(m->comp_type, m->ncomp, theta, m->N, x_ms, diag, r, &quad, &logdet);
		if (st != 0 || !isfinite(logdet)) {
			lp = -INFINITY;
			goto done;
		}
		lnl = -0.5 * (quad + logdet + m->N * log(2.0 * M_PI));
		if (!isfinite(lnl)) {
			lp = -INFINITY;
			goto done;
		}
	} else {
		/* transit only: plain chi^2 (model.py:333-338 with the prior sign fixed) */
		double s = 0.0;
		for (int n = 0; n < m->N; n++) {
			double e2 = m->yerr ? m->yerr[n] * m->yerr[n] : 1.0;
			s += -0.5 * (r[n] * r[n] / e2);
		}
		quad = -2.0 * s;
		logdet = 0.0;
		lnl = s;
	}
	lp = lnprior + lnl;
done:
	if (parts) {
		parts[0] = lnprior;
		parts[1] = quad;
		parts[2] = logdet;
		parts[3] = lnl;
	}
	return lp;
}

int orc_log_prob(int ncomp, const int *comp_type, int has_transit, const double *tr_fixed,
		 const unsigned char *tr_free, int ndim, const int *prior_kind, const double *prior_a,
		 const double *prior_b, int supersample, double exp_time, int ellip_mode, int inc_mode, int N,
		 const double *t_days, const double *y, const double *yerr, const double *theta, int W,
		 double *lp, double *parts, int nthreads)
{
	orc_model m;
	m.ncomp = ncomp;
	m.comp_type = comp_type;
	m.ndim_gp = model_ndim_gp(comp_type, ncomp);
	m.has_transit = has_transit;
	m.tr_fixed = tr_fixed;
	m.tr_free = tr_free;
	m.ndim = ndim;
	m.prior_kind = prior_kind;
	m.prior_a = prior_a;
	m.prior_b = prior_b;
	m.supersample = supersample;
	m.exp_time = exp_time;
	m.ellip_mode = ellip_mode;
	m.inc_mode = inc_mode;
	m.N = N;
	m.t_days = t_days;
	m.y = y;
	m.yerr = yerr;
	double *x_ms = malloc(sizeof(double) * (size_t)N);
	double *diag = malloc(sizeof(double) * (size_t)N);
	/* model.py:183-187: t*0.0864, yerr**2 (celerite squares it) */
	const double days_to_microsec = (24 * 3600) / 1e6;
	for (int n = 0; n < N; n++) {
		x_ms[n] = t_days[n] * days_to_microsec;
		diag[n] = yerr ? yerr[n] * yerr[n] : 1.123e-12 * 1.123e-12;
	}
	if (nthreads < 1) nthreads = 1;
#pragma omp parallel num_threads(nthreads)
	{
		double *resid = malloc(sizeof(double) * (size_t)N);
		double *flux = malloc(sizeof(double) * (size_t)N);
#pragma omp for schedule(dynamic, 1)
		for (int w = 0; w < W; w++)
			lp[w] = log_prob_one(&m, theta + (size_t)w * ndim, parts ? parts + 4 * (size_t)w : NULL, x_ms,
					     diag, resid, flux);
		free(resid);
		free(flux);
	}
	free(x_ms);
	free(diag);
	return 0;
}

/* ------------------------------------------------------------------------- */
/* Philox4x32-10 (Salmon et al. 2011, Random123).                              */
/* ------------------------------------------------------------------------- */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
	uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
	uint32_t k0 = key[0], k1 = key[1];
	for (int r = 0; r < 10; r++) {
		uint64_t p0 = (uint64_t)0xD2511F53u * c0;
		uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
		uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
		uint32_t n1 = (uint32_t)p1;
		uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
		uint32_t n3 = (uint32_t)p0;
		c0 = n0;
		c1 = n1;
		c2 = n2;
		c3 = n3;
		k0 += 0x9E3779B9u;
		k1 += 0xBB67AE85u;
	}
	out[0] = c0;
	out[1] = c1;
	out[2] = c2;
	out[3] = c3;
}

/* Stream layout shared with the device sampler (DESIGN.md "RNG streams"):
 *   key = (seed lo, seed hi); counter = (walker, step lo, step hi | purpose<<28, star)
 *   purpose 0: split key   -> u64 = w0 | w1<<32
 *   purpose 1: stretch     -> u_z = 53 bits of (w0,w1); u_acc = 53 bits of (w2,w3)
 *   purpose 2: partner     -> floor(w0 * Nc / 2^32)
 */
static void rng_words(uint64_t seed, uint32_t star, uint64_t step, uint32_t walker, uint32_t purpose,
		      uint32_t out[4])
{
	uint32_t ctr[4] = {walker, (uint32_t)step, (uint32_t)((step >> 32) & 0x0FFFFFFFu) | (purpose << 28), star};
	uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
	orc_philox4x32_10(ctr, key, out);
}

static double u53(uint32_t lo, uint32_t hi)
{
	uint64_t v = ((uint64_t)hi << 32) | lo;
	return (double)(v >> 11) * (1.0 / 9007199254740992.0);
}

/* ------------------------------------------------------------------------- */
/* S1: emcee EnsembleSampler.run_mcmc with the default StretchMove(a=2)        */
/* (RedBlueMove.propose: random half/half split, two half-steps, complement    */
/* taken from the CURRENT state).  chain[W][nsteps][ndim], logp[nsteps][W]     */
/* like sampler.chain / get_log_prob() (model.py:283-284).                     */
/* lp0 may be NULL (computed).  step0 = iteration offset (resume).             */
/* ------------------------------------------------------------------------- */
int orc_sample(int ncomp, const int *comp_type, int has_transit, const double *tr_fixed,
	       const unsigned char *tr_free, int ndim, const int *prior_kind, const double *prior_a,
	       const double *prior_b, int supersample, double exp_time, int ellip_mode, int inc_mode, int N,
	       const double *t_days, const double *y, const double *yerr, int W, const double *init,
	       long nsteps, uint64_t seed, uint32_t star, uint64_t step0, double a_stretch, double *chain,
	       double *logp, long *naccept, int nthreads)
{
	if (W % 2 || W < 2) return -1;
	const int H = W / 2;
	double *x = malloc(sizeof(double) * (size_t)W * ndim);
	double *lp = malloc(sizeof(double) * (size_t)W);
	double *q = malloc(sizeof(double) * (size_t)H * ndim);
	double *lq = malloc(sizeof(double) * (size_t)H);
	double *fac = malloc(sizeof(double) * (size_t)H);
	uint64_t *keys = malloc(sizeof(uint64_t) * (size_t)W);
	int *set = malloc(sizeof(int) * (size_t)W);
	int *act = malloc(sizeof(int) * (size_t)H), *oth = malloc(sizeof(int) * (size_t)H);
	memcpy(x, init, sizeof(double) * (size_t)W * ndim);
	orc_log_prob(ncomp, comp_type, has_transit, tr_fixed, tr_free, ndim, prior_kind, prior_a, prior_b,
		     supersample, exp_time, ellip_mode, inc_mode, N, t_days, y, yerr, x, W, lp, NULL, nthreads);
	for (int w = 0; w < W; w++)
		if (naccept) naccept[w] = 0;
	for (long s = 0; s < nsteps; s++) {
		uint64_t step = step0 + (uint64_t)s;
		uint32_t r[4];
		for (int w = 0; w < W; w++) {
			rng_words(seed, star, step, (uint32_t)w, 0, r);
			keys[w] = ((uint64_t)r[1] << 32) | r[0];
		}
		for (int w = 0; w < W; w++) {
			int rank = 0;
			for (int v = 0; v < W; v++)
				if (keys[v] < keys[w] || (keys[v] == keys[w] && v < w)) rank++;
			set[w] = rank < H ? 0 : 1;
		}
		for (int half = 0; half < 2; half++) {
			int na = 0, no = 0;
			for (int w = 0; w < W; w++) {
				if (set[w] == half)
					act[na++] = w;
				else
					oth[no++] = w;
			}
			for (int i = 0; i < H; i++) {
				int w = act[i];
				rng_words(seed, star, step, (uint32_t)w, 1, r);
				double uz = u53(r[0], r[1]);
				double zz = ((a_stretch - 1.0) * uz + 1.0);
				zz = zz * zz / a_stretch;
				fac[i] = (ndim - 1.0) * log(zz);
				uint32_t r2[4];
				rng_words(seed, star, step, (uint32_t)w, 2, r2);
				int j = oth[(int)(((uint64_t)r2[0] * (uint64_t)H) >> 32)];
				for (int d = 0; d < ndim; d++) {
					double cj = x[(size_t)j * ndim + d];
					q[(size_t)i * ndim + d] = cj - (cj - x[(size_t)w * ndim + d]) * zz;
				}
			}
			orc_log_prob(ncomp, comp_type, has_transit, tr_fixed, tr_free, ndim, prior_kind, prior_a,
				     prior_b, supersample, exp_time, ellip_mode, inc_mode, N, t_days, y, yerr, q, H, lq, NULL,
				     nthreads);
			for (int i = 0; i < H; i++) {
				int w = act[i];
				rng_words(seed, star, step, (uint32_t)w, 1, r);
				double ua = u53(r[2], r[3]);
				double lnpdiff = fac[i] + lq[i] - lp[w];
				if (lnpdiff > log(ua)) {
					memcpy(x + (size_t)w * ndim, q + (size_t)i * ndim, sizeof(double) * ndim);
					lp[w] = lq[i];
					if (naccept) naccept[w]++;
				}
			}
		}
		for (int w = 0; w < W; w++) {
			memcpy(chain + ((size_t)w * nsteps + s) * ndim, x + (size_t)w * ndim, sizeof(double) * ndim);
			logp[(size_t)s * W + w] = lp[w];
		}
	}
	free(x);
	free(lp);
	free(q);
	free(lq);
	free(fac);
	free(keys);
	free(set);
	free(act);
	free(oth);
	return 0;
}
