// Cost of one celerite step when the J x J state of a (walker, chunk) is split over a TEAM of lanes
// (one SHO term = two rows of P^ per lane, partial sums and (V~ - t^) exchanged by shuffles) against
// the one-thread-per-walker step of gp_chunk_kernel.  Both run the scaled-state step at constant
// cadence (ORDER 0) on identical synthetic inputs; final states are compared.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/team_step_probe tools/team_step_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "../gptransits_b200/csrc/kernels.cuh"

using namespace gpt;

constexpr int NT = 3, J = 6;

struct Walker {
	double a[NT], b[NT], c[NT], d[NT], asum;
};

__global__ void __launch_bounds__(128, 2) single_kernel(const Walker *wk, const double *y, int nsteps, double dt, double *out)
{
	const int w = blockIdx.y * blockDim.x + threadIdx.x;
	const Walker W = wk[w];
	FastCoef<NT> K;
	for (int k = 0; k < NT; k++) {
		K.c[k] = W.c[k];
		K.d[k] = W.d[k];
	}
	K.set_base(dt);
	GPCore<NT> S;
	for (int i = 0; i < J * (J + 1) / 2; i++) S.P[i] = 0.0;
	for (int i = 0; i < J; i++) S.g[i] = 0.0;
	S.quad = 0.0;
	S.ldm = 1.0;
	S.zhi = 0;
	S.dhi = 0x7fefffff;
	S.lde = 0;
	double UC[J], VC[J];
	for (int k = 0; k < NT; k++) {
		UC[2 * k] = W.a[k];
		UC[2 * k + 1] = -W.b[k];
		VC[2 * k] = 1.0;
		VC[2 * k + 1] = 0.0;
	}
	__shared__ double s_y[256];
	for (int i = threadIdx.x; i < 256; i += blockDim.x) s_y[i] = y[i];
	__syncthreads();
	const long long t0 = clock64();
	for (int n = 0; n < nsteps; n++) {
		double UN[J], VN[J];
		if ((n & 63) == 0) {  // re-anchor: unscale
			for (int k = 0; k < NT; k++) {
				UC[2 * k] = W.a[k];
				UC[2 * k + 1] = -W.b[k];
				VC[2 * k] = 1.0;
				VC[2 * k + 1] = 0.0;
			}
			if (n) {
				double sc[NT];
				for (int k = 0; k < NT; k++) sc[k] = exp(-W.c[k] * dt * 64);
				for (int i = 0; i < J; i++) {
					for (int j = i; j < J; j++) S.P[tri(i, j, J)] *= sc[i / 2] * sc[j / 2];
					S.g[i] *= sc[i / 2];
				}
			}
		}
		auto prep = [&]() { fast_advance<NT, 0>(K, 0.0, UC, VC, UN, VN); };
		gp_step_scaled<NT, true>(S, UC, VC, s_y[n & 255], W.asum + 2500.0, prep);
		for (int i = 0; i < J; i++) {
			UC[i] = UN[i];
			VC[i] = VN[i];
		}
	}
	const long long t1 = clock64();
	out[3 * (blockIdx.x * gridDim.y * blockDim.x + w) + 0] = S.quad;
	out[3 * (blockIdx.x * gridDim.y * blockDim.x + w) + 1] = (double)S.lde * 0.693147180559945309417 + log(S.ldm);
	out[3 * (blockIdx.x * gridDim.y * blockDim.x + w) + 2] = (double)(t1 - t0);
}

// ---- team form: TL lanes per walker, lane k owns term k (rows 2k, 2k+1 of P^, its U~, V~, g)
constexpr int TL = 4;
__device__ __forceinline__ double team_get(double v, int src) { return __shfl_sync(0xffffffffu, v, src, TL); }
__device__ __forceinline__ double team_sum(double v)
{
	v += __shfl_xor_sync(0xffffffffu, v, 1, TL);
	v += __shfl_xor_sync(0xffffffffu, v, 2, TL);
	return v;
}

template <int VARIANT>
__global__ void __launch_bounds__(128, 4) team_kernel(const Walker *wk, const double *y, int nsteps, double dt, double *out)
{
	const int tid = blockIdx.y * blockDim.x + threadIdx.x;
	const int w = tid / TL, k = tid % TL;
	const Walker W = wk[w];
	const bool real = k < NT;
	const double a = real ? W.a[k] : 0.0, b = real ? W.b[k] : 0.0, c = real ? W.c[k] : 0.0, d = real ? W.d[k] : 0.0;
	FastCoef<1> K;
	K.c[0] = c;
	K.d[0] = d;
	K.set_base(dt);
	double P0[J], P1[J], g0 = 0.0, g1 = 0.0;
	for (int j = 0; j < J; j++) P0[j] = P1[j] = 0.0;
	double quad = 0.0, ldm = 1.0;
	int lde = 0;
	double U0 = a, U1 = -b, V0 = 1.0, V1 = 0.0;
	double Ua[J];
	__shared__ double s_y[256];
	for (int i = threadIdx.x; i < 256; i += blockDim.x) s_y[i] = y[i];
	__syncthreads();
	auto gather = [&](double v0, double v1, double (&all)[J]) {
#pragma unroll
		for (int t = 0; t < NT; t++) {
			all[2 * t] = team_get(v0, t);
			all[2 * t + 1] = team_get(v1, t);
		}
	};
	gather(U0, U1, Ua);
	const long long t0 = clock64();
	for (int n = 0; n < nsteps; n++) {
		if ((n & 63) == 0) {
			U0 = a;
			U1 = -b;
			V0 = 1.0;
			V1 = 0.0;
			gather(U0, U1, Ua);
			if (n) {
				const double sk = exp(-c * dt * 64);
				double sc[NT];
#pragma unroll
				for (int t = 0; t < NT; t++) sc[t] = team_get(sk, t);
#pragma unroll
				for (int j = 0; j < J; j++) {
					P0[j] *= sk * sc[j / 2];
					P1[j] *= sk * sc[j / 2];
				}
				g0 *= sk;
				g1 *= sk;
			}
		}
		const double yv = s_y[n & 255], Avar = W.asum + 2500.0;
		// rows of t^ = P^ U~ (two accumulators per row)
		double ta = P0[0] * Ua[0], tb = P0[1] * Ua[1], tc = P1[0] * Ua[0], td = P1[1] * Ua[1];
#pragma unroll
		for (int j = 2; j < J; j += 2) {
			ta = fma(P0[j], Ua[j], ta);
			tb = fma(P0[j + 1], Ua[j + 1], tb);
			tc = fma(P1[j], Ua[j], tc);
			td = fma(P1[j + 1], Ua[j + 1], td);
		}
		const double tv0 = ta + tb, tv1 = tc + td;
		const double pu = fma(U0, tv0, U1 * tv1);
		const double pz = fma(U0, g0, U1 * g1);
		const double vt0 = V0 - tv0, vt1 = V1 - tv1;
		double vta[J];
		gather(vt0, vt1, vta);
		const double ut = team_sum(pu), zp = team_sum(pz);
		const double D = Avar - ut, z = yv - zp;
		const double invD = fast_rcp(D);
		// next step's U~, V~ (independent of the chain above)
		double UN[2], VN[2];
		{
			const double UC2[2] = {U0, U1}, VC2[2] = {V0, V1};
			fast_advance<1, 0>(K, 0.0, UC2, VC2, UN, VN);
		}
		double Un[J];
		gather(UN[0], UN[1], Un);
		const double w0 = vt0 * invD, w1 = vt1 * invD;
#pragma unroll
		for (int j = 0; j < J; j++) {
			P0[j] = fma(w0, vta[j], P0[j]);
			P1[j] = fma(w1, vta[j], P1[j]);
		}
		g0 = fma(w0, z, g0);
		g1 = fma(w1, z, g1);
		const double zi = z * invD;
		quad = fma(z, zi, quad);
		ldm *= D;
		const int hi = __double2hiint(ldm);
		const int ex = ((hi >> 20) & 0x7ff) - 1023;
		lde += ex;
		ldm = __hiloint2double(hi - (ex << 20), __double2loint(ldm));
		U0 = UN[0];
		U1 = UN[1];
		V0 = VN[0];
		V1 = VN[1];
#pragma unroll
		for (int j = 0; j < J; j++) Ua[j] = Un[j];
	}
	const long long t1 = clock64();
	if (k == 0) {
		const int gw = blockIdx.x * gridDim.y * (blockDim.x / TL) + w;
		out[3 * gw + 0] = quad;
		out[3 * gw + 1] = (double)lde * 0.693147180559945309417 + log(ldm);
		out[3 * gw + 2] = (double)(t1 - t0);
	}
}

// ---- lookahead form of the one-thread step: the product q = P U' with the matrix BEFORE the pending rank-one
// update, corrected by the scalar sigma = (v . U') / D once D is known, so that the reciprocal's latency and the
// matrix update are off the dependent chain (chain per step: sigma -> D -> 1/D).
template <int DUMMY>
__global__ void __launch_bounds__(128, 2) look_kernel(const Walker *wk, const double *y, int nsteps, double dt, double *out)
{
	constexpr int NP = J * (J + 1) / 2;
	const int w = blockIdx.y * blockDim.x + threadIdx.x;
	const Walker W = wk[w];
	FastCoef<NT> K;
	for (int k = 0; k < NT; k++) {
		K.c[k] = W.c[k];
		K.d[k] = W.d[k];
	}
	K.set_base(dt);
	double P[NP], g[J];
	for (int i = 0; i < NP; i++) P[i] = 0.0;
	for (int i = 0; i < J; i++) g[i] = 0.0;
	double quad = 0.0, ldm = 1.0;
	int lde = 0;
	__shared__ double s_y[256];
	for (int i = threadIdx.x; i < 256; i += blockDim.x) s_y[i] = y[i];
	__syncthreads();
	const double Avar = W.asum + 2500.0;
	double UC[J], VC[J];   // U~, V~ of the current step n
	double q[J];           // P_{n-1} U_n
	double vp[J], wp[J];   // v_{n-1}, w_{n-1}: the rank-one update still pending on P, g
	double zp = 0.0, invDp = 0.0, cn = 0.0, kap = 0.0, zeta = 0.0;
	auto prime = [&](int n) {
		for (int k = 0; k < NT; k++) {
			UC[2 * k] = W.a[k];
			UC[2 * k + 1] = -W.b[k];
			VC[2 * k] = 1.0;
			VC[2 * k + 1] = 0.0;
		}
		double uq = 0.0, ug = 0.0;
#pragma unroll
		for (int i = 0; i < J; i++) {
			double a_ = 0.0;
#pragma unroll
			for (int j = 0; j < J; j++) a_ = fma((i <= j) ? P[tri(i, j, J)] : P[tri(j, i, J)], UC[j], a_);
			q[i] = a_;
			uq = fma(UC[i], a_, uq);
			ug = fma(UC[i], g[i], ug);
			vp[i] = wp[i] = 0.0;
		}
		kap = Avar - uq;
		zeta = s_y[n & 255] - ug;
		zp = 0.0;
		invDp = 0.0;
		cn = 0.0;
	};
	auto flush = [&]() {  // apply the pending update
#pragma unroll
		for (int i = 0; i < J; i++) {
#pragma unroll
			for (int j = i; j < J; j++) P[tri(i, j, J)] = fma(vp[i], wp[j], P[tri(i, j, J)]);
			g[i] = fma(wp[i], zp, g[i]);
		}
	};
	const long long t0 = clock64();
	for (int n = 0; n < nsteps; n++) {
		if ((n & 63) == 0) {
			if (n) {
				flush();
				double sc[NT];
				for (int k = 0; k < NT; k++) sc[k] = exp(-W.c[k] * dt * 64);
				for (int i = 0; i < J; i++) {
					for (int j = i; j < J; j++) P[tri(i, j, J)] *= sc[i / 2] * sc[j / 2];
					g[i] *= sc[i / 2];
				}
			}
			prime(n);
		}
		// ---- dependent chain
		const double sig = cn * invDp;
		const double D = fma(-sig, cn, kap);
		const double z = fma(-sig, zp, zeta);
		double v[J];
#pragma unroll
		for (int i = 0; i < J; i++) v[i] = VC[i] - fma(vp[i], sig, q[i]);
		const double invD = fast_rcp(D);
		// ---- off the chain: pending update, next step's inputs and products
		flush();
		double UN[J], VN[J];
		fast_advance<NT, 0>(K, 0.0, UC, VC, UN, VN);
		double uq = 0.0, ug = 0.0, c1 = 0.0;
#pragma unroll
		for (int i = 0; i < J; i++) {
			double a_ = 0.0;
#pragma unroll
			for (int j = 0; j < J; j++) a_ = fma((i <= j) ? P[tri(i, j, J)] : P[tri(j, i, J)], UN[j], a_);
			q[i] = a_;
			uq = fma(UN[i], a_, uq);
			ug = fma(UN[i], g[i], ug);
			c1 = fma(v[i], UN[i], c1);
		}
		kap = Avar - uq;
		zeta = s_y[(n + 1) & 255] - ug;
		cn = c1;
#pragma unroll
		for (int i = 0; i < J; i++) {
			vp[i] = v[i];
			wp[i] = v[i] * invD;
			UC[i] = UN[i];
			VC[i] = VN[i];
		}
		zp = z;
		invDp = invD;
		const double zi = z * invD;
		quad = fma(z, zi, quad);
		ldm *= D;
		const int hi = __double2hiint(ldm);
		const int ex = ((hi >> 20) & 0x7ff) - 1023;
		lde += ex;
		ldm = __hiloint2double(hi - (ex << 20), __double2loint(ldm));
	}
	const long long t1 = clock64();
	out[3 * (blockIdx.x * gridDim.y * blockDim.x + w) + 0] = quad;
	out[3 * (blockIdx.x * gridDim.y * blockDim.x + w) + 1] = (double)lde * 0.693147180559945309417 + log(ldm);
	out[3 * (blockIdx.x * gridDim.y * blockDim.x + w) + 2] = (double)(t1 - t0);
}

int main(int argc, char **argv)
{
	const int nsteps = argc > 1 ? atoi(argv[1]) : 4096;
	const int NW = 128;
	std::vector<Walker> hw(NW);
	srand(1);
	for (auto &w : hw) {
		double as = 0;
		for (int k = 0; k < NT; k++) {
			const double amp = 100.0 + 300.0 * rand() / RAND_MAX, freq = (k == 2 ? 150.0 : 10.0 + 40.0 * k) * (0.7 + 0.6 * rand() / RAND_MAX);
			const double S0 = amp * amp / freq, w0 = 6.2831853 * freq, Q = k == 2 ? 5.0 : 0.70710678;
			const double f = sqrt(4 * Q * Q - 1);
			w.a[k] = S0 * w0 * Q * 1e-3;
			w.b[k] = w.a[k] / f;
			w.c[k] = 0.5 * w0 / Q * 1e-3;
			w.d[k] = w.c[k] * f;
			as += w.a[k];
		}
		w.asum = as + 1e4;
	}
	std::vector<double> hy(256);
	for (auto &v : hy) v = 200.0 * (rand() / (double)RAND_MAX - 0.5);
	Walker *dw;
	double *dy, *dout;
	cudaMalloc(&dw, sizeof(Walker) * NW);
	cudaMalloc(&dy, 8 * 256);
	cudaMalloc(&dout, 8 * 3 * NW * 4096);
	cudaMemcpy(dw, hw.data(), sizeof(Walker) * NW, cudaMemcpyHostToDevice);
	cudaMemcpy(dy, hy.data(), 8 * 256, cudaMemcpyHostToDevice);
	const double dt = 1.7654;
	std::vector<double> o1(3 * NW), o2(3 * NW);
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0);
	cudaEventCreate(&e1);
	auto time_it = [&](auto launch, const char *name, int chunks, int lanes_per_walker) {
		launch();
		cudaDeviceSynchronize();
		cudaEventRecord(e0);
		launch();
		cudaEventRecord(e1);
		cudaDeviceSynchronize();
		float ms;
		cudaEventElapsedTime(&ms, e0, e1);
		const cudaError_t err = cudaGetLastError();
		const double warps = (double)chunks * NW * lanes_per_walker / 32.0;
		printf("%-8s chunks %4d  warps/scheduler %.2f  %.3f ms  %.0f cycles per step (1.965 GHz)  %s\n", name, chunks, warps / 592.0, ms,
		       ms * 1e-3 * 1.965e9 / nsteps, err == cudaSuccess ? "" : cudaGetErrorString(err));
	};
	for (int chunks : {148, 296}) time_it([&] { single_kernel<<<dim3(chunks, NW / 128), 128>>>(dw, dy, nsteps, dt, dout); }, "single", chunks, 1);
	cudaMemcpy(o1.data(), dout, 8 * 3 * NW, cudaMemcpyDeviceToHost);
	std::vector<double> o3(3 * NW);
	for (int chunks : {148, 296}) time_it([&] { look_kernel<0><<<dim3(chunks, NW / 128), 128>>>(dw, dy, nsteps, dt, dout); }, "look", chunks, 1);
	cudaMemcpy(o3.data(), dout, 8 * 3 * NW, cudaMemcpyDeviceToHost);
	{
		double eq = 0, el = 0;
		for (int w = 0; w < NW; w++) {
			eq = fmax(eq, fabs(o1[3 * w] - o3[3 * w]) / fabs(o1[3 * w]));
			el = fmax(el, fabs(o1[3 * w + 1] - o3[3 * w + 1]) / fabs(o1[3 * w + 1]));
		}
		printf("lookahead vs single: max rel diff quad %.2e logdet %.2e; clock64 per step %.0f\n", eq, el, o3[2] / nsteps);
	}
	for (int chunks : {37, 74})
		time_it([&] { team_kernel<0><<<dim3(chunks, NW * TL / 128), 128>>>(dw, dy, nsteps, dt, dout); }, "team", chunks, TL);
	cudaMemcpy(o2.data(), dout, 8 * 3 * NW, cudaMemcpyDeviceToHost);
	double eq = 0, el = 0;
	for (int w = 0; w < NW; w++) {
		eq = fmax(eq, fabs(o1[3 * w] - o2[3 * w]) / fabs(o1[3 * w]));
		el = fmax(el, fabs(o1[3 * w + 1] - o2[3 * w + 1]) / fabs(o1[3 * w + 1]));
	}
	printf("team vs single: max rel diff quad %.2e logdet %.2e  (quad %.6e logdet %.6e)\n", eq, el, o1[0], o1[1]);
	printf("clock64 per step: single %.0f team %.0f\n", o1[2] / nsteps, o2[2] / nsteps);
	return 0;
}
