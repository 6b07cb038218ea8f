// FFTC2CWrapperGPU::forward / backward (libzen/fftw.h:35-43) without cuFFT:
// one CTA runs the shared-memory Stockham FFT of fft_smem.cuh on the whole
// transform.  Unnormalised in both directions, like cufftExecC2C.
#include <algorithm>
#include <cmath>
#include <mutex>
#include <vector>

#include "fft_smem.cuh"
#include "zen_common.cuh"

using namespace zen_b200;

namespace {

template <int N, int NT, int S>
__global__ void __launch_bounds__(NT) fft_c2c_kernel(float2* __restrict__ data, const float2* __restrict__ tw)
{
	extern __shared__ __align__(16) unsigned char fft_smem_raw[];
	float2* buf = reinterpret_cast<float2*>(fft_smem_raw);
	for (int i = threadIdx.x; i < N; i += NT)
		buf[i] = data[i];
	__syncthreads();
	fft_smem<N, NT, S>(buf, tw, threadIdx.x);
	for (int i = threadIdx.x; i < N; i += NT)
		data[i] = buf[i];
}

// per-stage twiddle tables (fft_fill_twiddles), one per size and device, built on first use
struct TwiddleCache {
	std::mutex mu;
	float2* tab[8][20] = {};
	float2* scr[8] = {};
	size_t scr_n[8] = {};
	// scratch of the four-step transform, grown on demand (one per device; calls on one device are serialised by the caller's stream)
	float2* scratch(size_t n)
	{
		int dev = 0;
		if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 8)
			return nullptr;
		std::lock_guard<std::mutex> lk(mu);
		if (scr_n[dev] < n) {
			cudaFree(scr[dev]);
			scr[dev] = nullptr;
			scr_n[dev] = 0;
			if (cudaMalloc(&scr[dev], sizeof(float2) * n) != cudaSuccess)
				return nullptr;
			scr_n[dev] = n;
		}
		return scr[dev];
	}
	float2* get(int n, int order)
	{
		int dev = 0;
		if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 8)
			return nullptr;
		std::lock_guard<std::mutex> lk(mu);
		if (!tab[dev][order]) {
			const int cnt = std::max(1, fft_twiddle_count_rt(n));
			std::vector<float2> h(cnt);
			fft_fill_twiddles(n, h.data());
			float2* d = nullptr;
			if (cudaMalloc(&d, sizeof(float2) * cnt) != cudaSuccess)
				return nullptr;
			if (cudaMemcpy(d, h.data(), sizeof(float2) * cnt, cudaMemcpyHostToDevice) != cudaSuccess) {
				cudaFree(d);
				return nullptr;
			}
			tab[dev][order] = d;
		}
		return tab[dev][order];
	}
};
TwiddleCache g_tw;

template <int N>
int launch_fft(float2* d, const float2* tw, int inverse, cudaStream_t s)
{
	constexpr int NT = (N / 8) < 32 ? 32 : ((N / 8) > 512 ? 512 : (N / 8));
	size_t smem = sizeof(float2) * (size_t)fpad_size(N);
	if (inverse) {
		auto k = fft_c2c_kernel<N, NT, +1>;
		ZEN_CUDA_CHECK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
		k<<<1, NT, smem, s>>>(d, tw);
	}
	else {
		auto k = fft_c2c_kernel<N, NT, -1>;
		ZEN_CUDA_CHECK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
		k<<<1, NT, smem, s>>>(d, tw);
	}
	ZEN_CUDA_CHECK(cudaGetLastError());
	return ZEN_OK;
}


// ---- N = 32768, 65536: four-step FFT through global memory --------------------------------------------
// N = N1 * N2.  With n = n1*N2 + n2 and k = k1 + N1*k2:
//   X[k1 + N1*k2] = sum_{n2} [ w_N^(n2*k1) * sum_{n1} x[n1*N2 + n2] * w_N1^(n1*k1) ] * w_N2^(n2*k2)
// Kernel 1: one CTA per column n2 does the N1-point FFT over n1 (stride N2), multiplies by the twiddle
// w_N^(n2*k1) and stores Y[k1][n2] into a scratch buffer.  Kernel 2: one CTA per row k1 does the N2-point FFT
// over n2 and scatters X[k1 + N1*k2] back into the caller's array.
template <int N1, int N2, int NT, int S>
__global__ void __launch_bounds__(NT) fft_large_cols_kernel(const float2* __restrict__ x, float2* __restrict__ y, const float2* __restrict__ tw1)
{
	extern __shared__ __align__(16) unsigned char fft_smem_raw[];
	float2* buf = reinterpret_cast<float2*>(fft_smem_raw);
	const int n2 = blockIdx.x;
	for (int n1 = threadIdx.x; n1 < N1; n1 += NT)
		buf[n1] = x[(size_t)n1 * N2 + n2];
	__syncthreads();
	fft_smem<N1, NT, S>(buf, tw1, threadIdx.x);
	constexpr int N = N1 * N2;
	for (int k1 = threadIdx.x; k1 < N1; k1 += NT) {
		float sn, cs;
		sincospif(2.0f * (float)(((long)n2 * k1) % N) / (float)N, &sn, &cs);  // accurate to ~1 ulp on the reduced angle
		float2 w = make_float2(cs, S > 0 ? sn : -sn);
		y[(size_t)k1 * N2 + n2] = cmul(buf[k1], w);
	}
}

template <int N1, int N2, int NT, int S>
__global__ void __launch_bounds__(NT) fft_large_rows_kernel(const float2* __restrict__ y, float2* __restrict__ x, const float2* __restrict__ tw2)
{
	extern __shared__ __align__(16) unsigned char fft_smem_raw[];
	float2* buf = reinterpret_cast<float2*>(fft_smem_raw);
	const int k1 = blockIdx.x;
	for (int n2 = threadIdx.x; n2 < N2; n2 += NT)
		buf[n2] = y[(size_t)k1 * N2 + n2];
	__syncthreads();
	fft_smem<N2, NT, S>(buf, tw2, threadIdx.x);
	for (int k2 = threadIdx.x; k2 < N2; k2 += NT)
		x[(size_t)k1 + (size_t)N1 * k2] = buf[k2];
}

template <int N1, int N2>
int launch_fft_large(float2* d, int inverse, cudaStream_t s)
{
	constexpr int NT = 128;
	int o1 = 0, o2 = 0;
	while ((1 << o1) < N1) ++o1;
	while ((1 << o2) < N2) ++o2;
	const float2* tw1 = g_tw.get(N1, o1);
	const float2* tw2 = g_tw.get(N2, o2);
	float2* scratch = g_tw.scratch((size_t)N1 * N2);
	if (!tw1 || !tw2 || !scratch)
		return ZEN_ERR_CUDA;
	size_t sm1 = sizeof(float2) * (size_t)fpad_size(N1), sm2 = sizeof(float2) * (size_t)fpad_size(N2);
	if (inverse) {
		fft_large_cols_kernel<N1, N2, NT, +1><<<N2, NT, sm1, s>>>(d, scratch, tw1);
		fft_large_rows_kernel<N1, N2, NT, +1><<<N1, NT, sm2, s>>>(scratch, d, tw2);
	}
	else {
		fft_large_cols_kernel<N1, N2, NT, -1><<<N2, NT, sm1, s>>>(d, scratch, tw1);
		fft_large_rows_kernel<N1, N2, NT, -1><<<N1, NT, sm2, s>>>(scratch, d, tw2);
	}
	ZEN_CUDA_CHECK(cudaGetLastError());
	return ZEN_OK;
}

}  // namespace

// per-stage twiddle table of the n-point transform (device pointer, cached per device), for other translation units
namespace zen_b200 {
const float2* fft_twiddle_table(int n)
{
	int order = 0;
	while ((1 << order) < n)
		++order;
	return g_tw.get(n, order);
}
}  // namespace zen_b200

extern "C" int zen_fft_c2c(int nfft, float* d_inout, int inverse, void* cuda_stream)
{
	if (!d_inout || nfft < 2)
		return ZEN_ERR_ARG;
	if (!is_pow2(nfft) || nfft > 65536)
		return ZEN_ERR_UNSUPPORTED;
	// From 8192 points on the transform is spread over many CTAs (four-step: N1 column FFTs, twiddle, N2 row FFTs through
	// a scratch buffer): one CTA running all of it is bound by its own stage-to-stage latency (measured on a B200,
	// forward + backward: 8192 points 20.6 us in one CTA against cuFFT's 20.3, 16384 points 28.7 against 23.6, while the
	// four-step 32768-point transform takes 16.4 us; profiles/r02_fft_bench_vs_cufft.json).
	if (nfft == 8192)
		return launch_fft_large<64, 128>(reinterpret_cast<float2*>(d_inout), inverse, (cudaStream_t)cuda_stream);
	if (nfft == 16384)
		return launch_fft_large<128, 128>(reinterpret_cast<float2*>(d_inout), inverse, (cudaStream_t)cuda_stream);
	if (nfft == 32768)
		return launch_fft_large<128, 256>(reinterpret_cast<float2*>(d_inout), inverse, (cudaStream_t)cuda_stream);
	if (nfft == 65536)
		return launch_fft_large<256, 256>(reinterpret_cast<float2*>(d_inout), inverse, (cudaStream_t)cuda_stream);
	int order = 0;
	while ((1 << order) < nfft)
		++order;
	const float2* tw = g_tw.get(nfft, order);
	if (!tw)
		return ZEN_ERR_CUDA;
	float2* d = reinterpret_cast<float2*>(d_inout);
	cudaStream_t s = (cudaStream_t)cuda_stream;
	switch (nfft) {
	case 2: return launch_fft<2>(d, tw, inverse, s);
	case 4: return launch_fft<4>(d, tw, inverse, s);
	case 8: return launch_fft<8>(d, tw, inverse, s);
	case 16: return launch_fft<16>(d, tw, inverse, s);
	case 32: return launch_fft<32>(d, tw, inverse, s);
	case 64: return launch_fft<64>(d, tw, inverse, s);
	case 128: return launch_fft<128>(d, tw, inverse, s);
	case 256: return launch_fft<256>(d, tw, inverse, s);
	case 512: return launch_fft<512>(d, tw, inverse, s);
	case 1024: return launch_fft<1024>(d, tw, inverse, s);
	case 2048: return launch_fft<2048>(d, tw, inverse, s);
	case 4096: return launch_fft<4096>(d, tw, inverse, s);
	case 8192: return launch_fft<8192>(d, tw, inverse, s);
	case 16384: return launch_fft<16384>(d, tw, inverse, s);
	}
	return ZEN_ERR_UNSUPPORTED;
}
