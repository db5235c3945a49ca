// Micro-benchmark: dependent-chain latency and single-warp throughput of FP64 ops on this GPU.
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void chain(double *out, long long *cyc, int iters, double m, double b)
{
	double a[ILP];
	for (int i = 0; i < ILP; i++) a[i] = threadIdx.x * 1e-9 + i;
	long long t0 = clock64();
	for (int it = 0; it < iters; it++) {
#pragma unroll
		for (int i = 0; i < ILP; i++) a[i] = fma(a[i], m, b);
	}
	long long t1 = clock64();
	double s = 0;
	for (int i = 0; i < ILP; i++) s += a[i];
	out[threadIdx.x + blockIdx.x * blockDim.x] = s;
	if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
__global__ void rcpchain(double *out, long long *cyc, int iters)
{
	double a = 1.5 + threadIdx.x * 1e-3;
	long long t0 = clock64();
	for (int it = 0; it < iters; it++) {
		double r;
		asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
		a = r + 1.0;
	}
	long long t1 = clock64();
	out[threadIdx.x] = a;
	if (threadIdx.x == 0) *cyc = t1 - t0;
}
int main()
{
	double *out; long long *cyc, h;
	cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 8);
	const int iters = 4096;
#define RUN(ILP, W) chain<ILP><<<1, 32 * W>>>(out, cyc, iters, 0.999999, 1e-7); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); \
	printf("DFMA ILP=%d warps/SM=%d (1 block): %.2f cycles per dependent step, %.2f cycles per DFMA warp-instr per SMSP\n", ILP, W, (double)h / iters, (double)h / iters / ILP / ((W + 3) / 4));
	RUN(1, 1) RUN(2, 1) RUN(4, 1) RUN(8, 1) RUN(12, 1) RUN(1, 4) RUN(8, 4) RUN(8, 8) RUN(4, 16)
	rcpchain<<<1, 32>>>(out, cyc, iters); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
	printf("rcp.approx.f64 + DADD dependent: %.2f cycles per pair\n", (double)h / iters);
	return 0;
}
