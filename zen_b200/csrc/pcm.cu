// The on-disk format either side of the HPR path, batched on the device:
//   PCM16 (mono or interleaved stereo) -> float32 mono, as libnyquist decodes it for zen offline / fakert
//     (vendor/libnyquist/include/libnyquist/Common.h:296-302 int16_to_float32 = (float)s / 32767.f,
//      Common.h:669-675 StereoToMono = (l + r) / 2.0f; zen/offline.h:104-117, zen/fakert.h:117-130);
//   float32 -> peak-normalised PCM16, as the command line writes its outputs
//     (zen/offline.h:180-192 / zen/fakert.h:259-268: x /= max(-min, max);
//      vendor/libnyquist/src/Common.cpp:332-337: (int16_t)lroundf(x * 32767.f), no dither).
// Every stream (row) of a batch is independent.  Pure streaming kernels: 2-4 B in + 4 B out per frame for the decode,
// 4 B in (peak pass) + 4 B in + 2 B out for the encode; IEEE division / multiplication as the reference's host code.
#include <cstdint>

#include "zen_common.cuh"

namespace {

__global__ void pcm16_decode_kernel(const int16_t* __restrict__ pcm, long pcm_stride, int channels, float* __restrict__ out,
                                    long out_stride, long n_frames)
{
	const int16_t* src = pcm + (size_t)blockIdx.y * pcm_stride;
	float* dst = out + (size_t)blockIdx.y * out_stride;
	for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n_frames; i += (long)gridDim.x * blockDim.x) {
		float v;
		if (channels == 2) {
			const float l = __fdiv_rn((float)src[2 * i], 32767.0f), r = __fdiv_rn((float)src[2 * i + 1], 32767.0f);
			v = __fdiv_rn(__fadd_rn(l, r), 2.0f);
		}
		else {
			v = __fdiv_rn((float)src[i], 32767.0f);
		}
		dst[i] = v;
	}
}

// max(-min, max) of a row == max |x| ; non-negative floats order like their bit patterns
__global__ void peak_kernel(const float* __restrict__ in, long in_stride, long n, unsigned* __restrict__ peak_bits)
{
	const float* src = in + (size_t)blockIdx.y * in_stride;
	float m = 0.0f;
	for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
		m = fmaxf(m, fabsf(src[i]));
	for (int s = 16; s > 0; s >>= 1)
		m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, s));
	__shared__ float wm[32];
	if ((threadIdx.x & 31) == 0) wm[threadIdx.x >> 5] = m;
	__syncthreads();
	if (threadIdx.x < 32) {
		m = threadIdx.x < (blockDim.x + 31) / 32 ? wm[threadIdx.x] : 0.0f;
		for (int s = 16; s > 0; s >>= 1)
			m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, s));
		if (threadIdx.x == 0) atomicMax(peak_bits + blockIdx.y, __float_as_uint(m));
	}
}

__global__ void pcm16_encode_kernel(const float* __restrict__ in, long in_stride, long n, const unsigned* __restrict__ peak_bits,
                                    int16_t* __restrict__ out, long out_stride)
{
	const float* src = in + (size_t)blockIdx.y * in_stride;
	int16_t* dst = out + (size_t)blockIdx.y * out_stride;
	const float peak = __uint_as_float(peak_bits[blockIdx.y]);
	for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
		// a silent stream has no peak to divide by (the reference would write lroundf(NaN)): it stays silent
		const float x = peak > 0.0f ? __fdiv_rn(src[i], peak) : 0.0f;
		dst[i] = (int16_t)lroundf(__fmul_rn(x, 32767.0f));
	}
}

int grid_x_for(long n, int n_streams)
{
	int sms = 148, dev = 0;
	if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
	// about eight 256-thread CTAs per SM over the whole batch, at least one per row, no more than the row needs
	long per_row = ((long)sms * 8 + n_streams - 1) / n_streams;
	long need = (n + 255) / 256;
	long gx = per_row < need ? per_row : need;
	return (int)(gx < 1 ? 1 : gx);
}

}  // namespace

extern "C" {

int zen_pcm16_decode_mono(const int16_t* d_pcm, long pcm_stride, int channels, int n_streams, long n_frames, float* d_out, long out_stride)
{
	if (!d_pcm || !d_out || (channels != 1 && channels != 2) || n_streams < 1 || n_frames < 0 || pcm_stride < n_frames * channels
	    || out_stride < n_frames || n_streams > 65535)
		return ZEN_ERR_ARG;
	if (zen_device_count() <= 0)
		return ZEN_ERR_CUDA;
	if (n_frames == 0)
		return ZEN_OK;
	dim3 grid(grid_x_for(n_frames, n_streams), n_streams);
	pcm16_decode_kernel<<<grid, 256>>>(d_pcm, pcm_stride, channels, d_out, out_stride, n_frames);
	ZEN_CUDA_CHECK(cudaGetLastError());
	ZEN_CUDA_CHECK(cudaDeviceSynchronize());
	return ZEN_OK;
}

int zen_pcm16_encode_normalized(const float* d_in, long in_stride, int n_streams, long n, int16_t* d_out, long out_stride, float* d_peaks)
{
	if (!d_in || !d_out || !d_peaks || n_streams < 1 || n < 0 || in_stride < n || out_stride < n || n_streams > 65535)
		return ZEN_ERR_ARG;
	if (zen_device_count() <= 0)
		return ZEN_ERR_CUDA;
	ZEN_CUDA_CHECK(cudaMemset(d_peaks, 0, sizeof(float) * (size_t)n_streams));
	if (n == 0)
		return ZEN_OK;
	dim3 grid(grid_x_for(n, n_streams), n_streams);
	peak_kernel<<<grid, 256>>>(d_in, in_stride, n, reinterpret_cast<unsigned*>(d_peaks));
	ZEN_CUDA_CHECK(cudaGetLastError());
	pcm16_encode_kernel<<<grid, 256>>>(d_in, in_stride, n, reinterpret_cast<const unsigned*>(d_peaks), d_out, out_stride);
	ZEN_CUDA_CHECK(cudaGetLastError());
	ZEN_CUDA_CHECK(cudaDeviceSynchronize());
	return ZEN_OK;
}

}  // extern "C"
