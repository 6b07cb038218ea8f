// What HBM gives a streaming kernel on this box for the traffic mixes of the PCM16 kernels: read-only, write-only,
// copy (1:1), read 1 : write 2 (decode: 2 B in, 4 B out) and read 2 : write 1 (encode: 4 B in, 2 B out).
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a tools/stream_bench.cu -o tools/_build/stream_bench
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { std::fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

// each thread: R 16-byte loads and W 16-byte stores per trip (U trips unrolled), grid-stride over `n16` 16-byte units of the
// larger side; STREAM: ld.cs / st.cs hints
template <int R, int W, int U, bool STREAM>
__global__ void __launch_bounds__(256) mix_kernel(const int4* __restrict__ in, int4* __restrict__ out, long trips, int4* sink)
{
	int4 acc = make_int4(0, 0, 0, 0);
	for (long t = (long)blockIdx.x * U; t < trips; t += (long)gridDim.x * U) {
		int4 v[U][R > 0 ? R : 1];
#pragma unroll
		for (int u = 0; u < U; ++u)
#pragma unroll
			for (int r = 0; r < R; ++r) {
				const int4* p = in + ((t + u) * R + r) * 256 + threadIdx.x;
				v[u][r] = (t + u < trips) ? (STREAM ? __ldcs(p) : *p) : make_int4(0, 0, 0, 0);
			}
#pragma unroll
		for (int u = 0; u < U; ++u) {
			int4 x = make_int4((int)t, u, 0, 0);
#pragma unroll
			for (int r = 0; r < R; ++r) {
				x.x ^= v[u][r].x; x.y += v[u][r].y; x.z ^= v[u][r].z; x.w += v[u][r].w;
			}
			if (W == 0) { acc.x ^= x.x; acc.y += x.y; acc.z ^= x.z; acc.w += x.w; }
#pragma unroll
			for (int w = 0; w < W; ++w) {
				int4* p = out + ((t + u) * W + w) * 256 + threadIdx.x;
				if (t + u < trips) { if (STREAM) __stcs(p, x); else *p = x; }
			}
		}
	}
	if (W == 0 && acc.x == 0x12345678 && acc.y == 42) *sink = acc;
}

template <int R, int W, int U, bool STREAM>
static int run(const char* name, int4* a, int4* b, long bytes_big, int ctas_per_sm)
{
	// the larger side moves bytes_big bytes
	const int big = R > W ? R : W;
	const long trips = bytes_big / (16L * 256 * big);
	int sms = 148;
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0); cudaEventCreate(&e1);
	float best = 1e30f;
	for (int rep = 0; rep < 6; ++rep) {
		cudaEventRecord(e0);
		mix_kernel<R, W, U, STREAM><<<sms * ctas_per_sm, 256>>>(a, b, trips, b);
		cudaEventRecord(e1);
		CK(cudaEventSynchronize(e1));
		float ms; cudaEventElapsedTime(&ms, e0, e1);
		if (rep && ms < best) best = ms;
	}
	const double total = (double)trips * 16 * 256 * (R + W);
	std::printf(" \"%s\": {\"GBps\": %.0f, \"ms\": %.3f},\n", name, total / best / 1e6, best);
	return 0;
}

int main()
{
	const long bytes = 8L << 30;
	int4 *a, *b;
	CK(cudaMalloc(&a, bytes)); CK(cudaMalloc(&b, bytes));
	CK(cudaMemset(a, 1, bytes)); CK(cudaMemset(b, 2, bytes));
	std::printf("{\n");
	run<1, 0, 4, false>("read_only", a, b, bytes, 8);
	run<0, 1, 4, false>("write_only", a, b, bytes, 8);
	run<0, 1, 4, true>("write_only_cs", a, b, bytes, 8);
	run<1, 1, 4, false>("copy_1r_1w", a, b, bytes, 8);
	run<1, 1, 4, true>("copy_1r_1w_cs", a, b, bytes, 8);
	run<1, 2, 1, false>("decode_mix_1r_2w_u1", a, b, bytes, 8);
	run<1, 2, 4, false>("decode_mix_1r_2w_u4", a, b, bytes, 8);
	run<1, 2, 4, true>("decode_mix_1r_2w_u4_cs", a, b, bytes, 8);
	run<1, 2, 4, true>("decode_mix_1r_2w_u4_cs_4cta", a, b, bytes, 4);
	run<2, 1, 2, false>("encode_mix_2r_1w_u2", a, b, bytes, 8);
	run<2, 1, 2, true>("encode_mix_2r_1w_u2_cs", a, b, bytes, 8);
	run<2, 1, 4, true>("encode_mix_2r_1w_u4_cs", a, b, bytes, 8);
	{
		cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
		cudaMemcpyAsync(b, a, bytes, cudaMemcpyDeviceToDevice);
		cudaEventRecord(e0); cudaMemcpyAsync(b, a, bytes, cudaMemcpyDeviceToDevice); cudaEventRecord(e1); cudaEventSynchronize(e1);
		float ms; cudaEventElapsedTime(&ms, e0, e1);
		std::printf(" \"cudaMemcpy_d2d\": {\"GBps\": %.0f},\n", 2.0 * bytes / ms / 1e6);
		cudaEventRecord(e0); cudaMemsetAsync(b, 0, bytes); cudaEventRecord(e1); cudaEventSynchronize(e1);
		cudaEventElapsedTime(&ms, e0, e1);
		std::printf(" \"cudaMemset\": {\"GBps\": %.0f}\n}\n", 1.0 * bytes / ms / 1e6);
	}
	return 0;
}
