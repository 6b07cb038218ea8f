// Breaks the per-hop latency of the real-time path into its parts on the box it runs on:
// launch + synchronize of an empty kernel, mapped-memory round trips, and the fakert region.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>
#include "../include/zen_b200.h"

__global__ void empty_kernel() {}
__global__ void touch_mapped(const float* in, float* out, int n)
{
	for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = in[i] + 1.0f;
}
__global__ void flag_kernel(const float* in, float* out, int n, volatile int* flag, int v)
{
	for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = in[i] + 1.0f;
	__syncthreads();
	__threadfence_system();
	if (threadIdx.x == 0) *flag = v;
}

static double p50(std::vector<double>& v) { std::sort(v.begin(), v.end()); return v[v.size() / 2]; }

int main()
{
	using clk = std::chrono::high_resolution_clock;
	cudaStream_t s;
	cudaStreamCreate(&s);
	const int N = 3000;
	std::vector<double> t(N);
	for (int i = 0; i < 200; ++i) { empty_kernel<<<1, 32, 0, s>>>(); cudaStreamSynchronize(s); }
	for (int i = 0; i < N; ++i) {
		auto a = clk::now();
		empty_kernel<<<1, 32, 0, s>>>();
		cudaStreamSynchronize(s);
		t[i] = std::chrono::duration<double, std::micro>(clk::now() - a).count();
	}
	std::printf("empty kernel launch+streamSync        p50 %.2f us\n", p50(t));
	zen_io io;
	zen_io_alloc(&io, 1024);
	for (int i = 0; i < N; ++i) {
		auto a = clk::now();
		touch_mapped<<<1, 256, 0, s>>>(io.device_in, io.device_out, 1024);
		cudaStreamSynchronize(s);
		t[i] = std::chrono::duration<double, std::micro>(clk::now() - a).count();
	}
	std::printf("mapped 4KB in -> 4KB out kernel + sync  p50 %.2f us\n", p50(t));
	int* flag;
	cudaHostAlloc((void**)&flag, 64, cudaHostAllocMapped);
	int* dflag;
	cudaHostGetDevicePointer((void**)&dflag, flag, 0);
	*flag = 0;
	for (int i = 0; i < N; ++i) {
		auto a = clk::now();
		flag_kernel<<<1, 256, 0, s>>>(io.device_in, io.device_out, 1024, dflag, i + 1);
		while (*(volatile int*)flag != i + 1) {}
		t[i] = std::chrono::duration<double, std::micro>(clk::now() - a).count();
	}
	std::printf("same, completion by polling a mapped flag p50 %.2f us\n", p50(t));
	cudaStreamSynchronize(s);
	for (int hop : {256, 1024, 4096}) {
		const long n_h = 1500;
		std::vector<float> audio(n_h * hop), perc(n_h * hop);
		for (size_t i = 0; i < audio.size(); ++i) audio[i] = 0.3f * sinf(0.01f * i) + ((i % 7000) < 30 ? 0.7f : 0.0f);
		std::vector<double> us(n_h);
		for (int fused = 0; fused < 3; ++fused) {
			zen_fakert_run(44100.0f, hop, 2.5f, 0, audio.data(), n_h, 300, fused, perc.data(), us.data());
			std::printf("fakert region hop %4d %s p50 %.2f us\n", hop, fused == 2 ? "resident " : (fused ? "fused    " : "two-call "), p50(us));
		}
	}
	return 0;
}
