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
		buf[fpad(i)] = data[i];
	__syncthreads();
	fft_smem<N, NT, S>(buf, tw, threadIdx.x);
	for (int i = threadIdx.x; i < N; i += NT)
		data[i] = buf[fpad(i)];
}

// per-stage twiddle tables (fft_fill_twiddles), one per size and device, built on first use
struct TwiddleCache {
	std::mutex mu;
	float2* tab[8][20] = {};
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

}  // namespace

extern "C" int zen_fft_c2c(int nfft, float* d_inout, int inverse, void* cuda_stream)
{
	if (!d_inout || nfft < 2)
		return ZEN_ERR_ARG;
	if (!is_pow2(nfft) || nfft > 16384)
		return ZEN_ERR_UNSUPPORTED;
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
