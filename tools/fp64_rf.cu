// Does a DFMA with three distinct register-pair operands issue every 2 cycles or every 3?
#include <cstdio>
#include <cuda_runtime.h>
__global__ void rf3(double *out, long long *cyc, int iters)
{
	double a[8], b[8], c[8];
	for (int i = 0; i < 8; i++) { a[i] = threadIdx.x * 1e-9 + i; b[i] = 1.0 - 1e-7 * (i + 1); c[i] = 1e-7 * (i + 2); }
	long long t0 = clock64();
	for (int it = 0; it < iters; it++) {
#pragma unroll
		for (int i = 0; i < 8; i++) a[i] = fma(a[i], b[i], c[i]);      // 3 distinct pairs, none shared with neighbours
	}
	long long t1 = clock64();
	double s = 0; for (int i = 0; i < 8; i++) s += a[i] + b[i] + c[i];
	out[threadIdx.x] = s; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void rf2(double *out, long long *cyc, int iters, double m)
{
	double a[8], c[8];
	for (int i = 0; i < 8; i++) { a[i] = threadIdx.x * 1e-9 + i; c[i] = 1e-7 * (i + 2); }
	long long t0 = clock64();
	for (int it = 0; it < iters; it++) {
#pragma unroll
		for (int i = 0; i < 8; i++) a[i] = fma(a[i], m, c[i]);         // one operand shared by all (reuse cache)
	}
	long long t1 = clock64();
	double s = 0; for (int i = 0; i < 8; i++) s += a[i] + c[i];
	out[threadIdx.x] = s; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void mul2(double *out, long long *cyc, int iters)
{
	double a[8], b[8];
	for (int i = 0; i < 8; i++) { a[i] = 1.0 + threadIdx.x * 1e-9 + i * 1e-3; b[i] = 1.0 - 1e-9 * (i + 1); }
	long long t0 = clock64();
	for (int it = 0; it < iters; it++) {
#pragma unroll
		for (int i = 0; i < 8; i++) a[i] = a[i] * b[i];
	}
	long long t1 = clock64();
	double s = 0; for (int i = 0; i < 8; i++) s += a[i] + b[i];
	out[threadIdx.x] = s; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
int main()
{
	double *out; long long *cyc, h;
	cudaMalloc(&out, 1 << 16); cudaMalloc(&cyc, 8);
	const int iters = 4096;
	for (int w = 1; w <= 16; w *= 2) {
		rf3<<<1, 32 * w>>>(out, cyc, iters); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
		printf("DFMA 3 distinct operands, %2d warps on one SM (%d per sub-partition): %.2f cycles per warp-instr per sub-partition\n", w, (w + 3) / 4, (double)h / iters / 8 / ((w + 3) / 4));
	}
	rf2<<<1, 32>>>(out, cyc, iters, 0.999999); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
	printf("DFMA 2 distinct + 1 shared: %.2f cycles per instr\n", (double)h / iters / 8);
	mul2<<<1, 32>>>(out, cyc, iters); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
	printf("DMUL 2 distinct: %.2f cycles per instr\n", (double)h / iters / 8);
	return 0;
}
