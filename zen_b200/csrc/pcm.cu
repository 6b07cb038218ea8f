// The on-disk format either side of the HPR path, batched on the device:
//   PCM16 (mono or interleaved stereo) -> float32 mono, as libnyquist decodes it for zen offline / fakert
//     (vendor/libnyquist/include/libnyquist/Common.h:296-302 int16_to_float32 = (float)s / 32767.f,
//      Common.h:669-675 StereoToMono = (l + r) / 2.0f; zen/offline.h:104-117, zen/fakert.h:117-130);
//   float32 -> peak-normalised PCM16, as the command line writes its outputs
//     (zen/offline.h:180-192 / zen/fakert.h:259-268: x /= max(-min, max);
//      vendor/libnyquist/src/Common.cpp:332-337: (int16_t)lroundf(x * 32767.f), no dither).
// Every stream (row) of a batch is independent.  Pure streaming kernels, bound by HBM: 2-4 B in + 4 B out per
// frame for the decode, 4 B in (peak pass) + 4 B in + 2 B out for the encode.  Each thread moves 16 bytes per
// access (8 samples of PCM16, 4 floats); rows are folded into the x dimension of the grid, so any number of
// streams fits one launch.  The results are the reference's bit for bit:
//   * s / 32767.f is evaluated as q = s*r, q' = fma(fma(-q, 32767, s), r, q) with r = RN(1/32767): correctly
//     rounded for every int16 s (checked exhaustively, tests/test_pcm.py::test_div32767_exact), three FMA-pipe
//     instructions instead of the IEEE division's slow path;
//   * x / peak keeps the IEEE division (the divisor is arbitrary).
#include <cstdint>

#include "zen_common.cuh"

namespace {

__device__ __forceinline__ float pcm_to_float(int s)
{
	const float r = 1.0f / 32767.0f;  // RN(1/32767), folded at compile time
	const float x = (float)s;
	const float q = __fmul_rn(x, r);
	const float e = __fmaf_rn(-q, 32767.0f, x);
	return __fmaf_rn(e, r, q);
}

__device__ __forceinline__ float stereo_fold(int l, int r)
{
	return __fmul_rn(__fadd_rn(pcm_to_float(l), pcm_to_float(r)), 0.5f);  // (l + r) / 2.0f, exact halving
}

// Work decomposition shared by the three kernels: an item = 8 consecutive samples of one row (16 bytes of PCM16, two
// 16-byte accesses of float32); a unit = 256 x U items of ONE row; a CTA walks units with a grid stride, keeping
// (row, chunk) counters instead of dividing per item.  A thread owns items item0 + u * 256, u < U, and issues the loads
// of all of them before it touches the first: U x 16..32 bytes in flight per thread is what lets a streaming kernel
// reach HBM speed (one access per thread and trip left the decode at half of it).
template <int U, typename Body>
__device__ __forceinline__ void walk_units(long items_per_row, long total_units, Body body)
{
	const long chunks_per_row = (items_per_row + 256 * U - 1) / (256 * U);
	long row = blockIdx.x / chunks_per_row, chunk = blockIdx.x - row * chunks_per_row;  // one division per CTA, then counters
	for (long unit = blockIdx.x; unit < total_units; unit += gridDim.x, chunk += gridDim.x) {
		while (chunk >= chunks_per_row) {
			chunk -= chunks_per_row;
			++row;
		}
		body(row, chunk * (256 * U) + threadIdx.x);
	}
}

constexpr int DEC_U = 4, ENC_U = 2, PEAK_U = 4;

template <int CH, bool VEC>
__global__ void __launch_bounds__(256) pcm16_decode_kernel(const int16_t* __restrict__ pcm, long pcm_stride, float* __restrict__ out,
                                                          long out_stride, long n_frames, long items_per_row, long total_units, int stream_out)
{
	walk_units<DEC_U>(items_per_row, total_units, [&](long row, long item0) {
		const int16_t* rsrc = pcm + (size_t)row * pcm_stride;
		float* rdst = out + (size_t)row * out_stride;
		int4 p[DEC_U][CH];
		bool whole[DEC_U];
#pragma unroll
		for (int u = 0; u < DEC_U; ++u) {
			const long i0 = (item0 + u * 256) * 8;
			whole[u] = VEC && i0 + 8 <= n_frames;
			if (whole[u]) {
#pragma unroll
				for (int h = 0; h < CH; ++h)
					p[u][h] = __ldcs(reinterpret_cast<const int4*>(rsrc + (size_t)i0 * CH) + h);
			}
		}
#pragma unroll
		for (int u = 0; u < DEC_U; ++u) {
			const long i0 = (item0 + u * 256) * 8;
			if (whole[u]) {
				float v[8];
				if (CH == 1) {
					const int w[4] = {p[u][0].x, p[u][0].y, p[u][0].z, p[u][0].w};
#pragma unroll
					for (int j = 0; j < 4; ++j) {
						v[2 * j] = pcm_to_float((int)(short)(w[j] & 0xffff));
						v[2 * j + 1] = pcm_to_float(w[j] >> 16);
					}
				}
				else {
					const int w[8] = {p[u][0].x, p[u][0].y, p[u][0].z, p[u][0].w, p[u][CH - 1].x, p[u][CH - 1].y, p[u][CH - 1].z, p[u][CH - 1].w};
#pragma unroll
					for (int j = 0; j < 8; ++j)
						v[j] = stereo_fold((int)(short)(w[j] & 0xffff), w[j] >> 16);
				}
				float4* d = reinterpret_cast<float4*>(rdst + i0);
				if (stream_out) {  // an output larger than L2 is not going to be read from it: evict-first stores (+8 %)
					__stcs(d, make_float4(v[0], v[1], v[2], v[3]));
					__stcs(d + 1, make_float4(v[4], v[5], v[6], v[7]));
				}
				else {
					d[0] = make_float4(v[0], v[1], v[2], v[3]);
					d[1] = make_float4(v[4], v[5], v[6], v[7]);
				}
			}
			else {
				for (long i = i0; i < i0 + 8 && i < n_frames; ++i)
					rdst[i] = CH == 1 ? pcm_to_float(rsrc[i]) : stereo_fold(rsrc[2 * i], rsrc[2 * i + 1]);
			}
		}
	});
}

// max(-min, max) of a row == max |x| ; non-negative floats order like their bit patterns.
// The whole CTA works on one row per unit: warp reduction, one atomic per warp.
template <bool VEC>
__global__ void __launch_bounds__(256) peak_kernel(const float* __restrict__ in, long in_stride, long n, long items_per_row,
                                                  long total_units, unsigned* __restrict__ peak_bits)
{
	walk_units<PEAK_U>(items_per_row, total_units, [&](long row, long item0) {
		const float* rsrc = in + (size_t)row * in_stride;
		float4 a[PEAK_U][2];
		bool whole[PEAK_U];
#pragma unroll
		for (int u = 0; u < PEAK_U; ++u) {
			const long i0 = (item0 + u * 256) * 8;
			whole[u] = VEC && i0 + 8 <= n;
			if (whole[u]) {
				a[u][0] = __ldg(reinterpret_cast<const float4*>(rsrc + i0));
				a[u][1] = __ldg(reinterpret_cast<const float4*>(rsrc + i0) + 1);
			}
		}
		float m = 0.0f;
#pragma unroll
		for (int u = 0; u < PEAK_U; ++u) {
			const long i0 = (item0 + u * 256) * 8;
			if (whole[u])
				m = fmaxf(m, fmaxf(fmaxf(fmaxf(fabsf(a[u][0].x), fabsf(a[u][0].y)), fmaxf(fabsf(a[u][0].z), fabsf(a[u][0].w))),
				                   fmaxf(fmaxf(fabsf(a[u][1].x), fabsf(a[u][1].y)), fmaxf(fabsf(a[u][1].z), fabsf(a[u][1].w)))));
			else
				for (long i = i0; i < i0 + 8 && i < n; ++i)
					m = fmaxf(m, fabsf(rsrc[i]));
		}
		for (int s = 16; s > 0; s >>= 1)
			m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, s));
		if ((threadIdx.x & 31) == 0 && m > 0.0f) atomicMax(peak_bits + row, __float_as_uint(m));
	});
}

__device__ __forceinline__ int pcm_from_float(float v, float peak)
{
	// a silent stream has no peak to divide by: it stays silent.  (The reference divides 0 by 0 and hands NaN to
	// lroundf; its x86-64 build converts that to PCM16 zeros as well - tests/golden/nyq_pcm.npz, silence_out.)
	const float x = peak > 0.0f ? __fdiv_rn(v, peak) : 0.0f;
	return (int)(int16_t)lroundf(__fmul_rn(x, 32767.0f));
}

// The same integer with the per-row work hoisted: rcp = RN(1 / peak) once per thread, then q = v * rcp and two
// residual corrections (e = fma(-q, peak, v) is exact; q += e * rcp) - the sequence IEEE division itself runs after
// refining its reciprocal, minus that refinement and the range check per sample.  Valid while peak and every |v| <= peak
// keep all intermediates normal or far below half a PCM16 step, which pcm_fast_range() checks per row.
// lroundf(t), t = RN(q * 32767): with u = trunc(2 t) (2 t = RN(q * 65534) exactly) it is (u + sign(u)) / 2 truncated.
__device__ __forceinline__ bool pcm_fast_range(float peak) { return peak > 1e-18f && peak < 1e18f; }
__device__ __forceinline__ int pcm_from_float_fast(float v, float peak, float rcp)
{
	float q = __fmul_rn(v, rcp);
	q = __fmaf_rn(__fmaf_rn(-q, peak, v), rcp, q);
	q = __fmaf_rn(__fmaf_rn(-q, peak, v), rcp, q);
	const int u = __float2int_rz(__fmul_rn(q, 65534.0f));
	const int w = u + ((u >> 31) | 1);
	return (w + (int)((unsigned)w >> 31)) >> 1;
}

template <bool VEC>
__global__ void __launch_bounds__(256) pcm16_encode_kernel(const float* __restrict__ in, long in_stride, long n,
                                                          const unsigned* __restrict__ peak_bits, int16_t* __restrict__ out,
                                                          long out_stride, long items_per_row, long total_units)
{
	walk_units<ENC_U>(items_per_row, total_units, [&](long row, long item0) {
		const float* rsrc = in + (size_t)row * in_stride;
		int16_t* rdst = out + (size_t)row * out_stride;
		float4 a[ENC_U][2];
		bool whole[ENC_U];
#pragma unroll
		for (int u = 0; u < ENC_U; ++u) {
			const long i0 = (item0 + u * 256) * 8;
			whole[u] = VEC && i0 + 8 <= n;
			if (whole[u]) {
				a[u][0] = __ldcs(reinterpret_cast<const float4*>(rsrc + i0));
				a[u][1] = __ldcs(reinterpret_cast<const float4*>(rsrc + i0) + 1);
			}
		}
		const float peak = __uint_as_float(__ldg(peak_bits + row));
		const bool fast = pcm_fast_range(peak);  // uniform over the CTA: one row per unit
		const float rcp = fast ? __frcp_rn(peak) : 0.0f;
#pragma unroll
		for (int u = 0; u < ENC_U; ++u) {
			const long i0 = (item0 + u * 256) * 8;
			if (whole[u]) {
				const float v[8] = {a[u][0].x, a[u][0].y, a[u][0].z, a[u][0].w, a[u][1].x, a[u][1].y, a[u][1].z, a[u][1].w};
				int w[4];
				if (fast) {
#pragma unroll
					for (int j = 0; j < 4; ++j)
						w[j] = (pcm_from_float_fast(v[2 * j], peak, rcp) & 0xffff) | (pcm_from_float_fast(v[2 * j + 1], peak, rcp) << 16);
				}
				else {
#pragma unroll 1
					for (int j = 0; j < 4; ++j)
						w[j] = (pcm_from_float(v[2 * j], peak) & 0xffff) | (pcm_from_float(v[2 * j + 1], peak) << 16);
				}
				__stcs(reinterpret_cast<int4*>(rdst + i0), make_int4(w[0], w[1], w[2], w[3]));
			}
			else {
				for (long i = i0; i < i0 + 8 && i < n; ++i)
					rdst[i] = (int16_t)pcm_from_float(rsrc[i], peak);
			}
		}
	});
}

// work units (row, chunk of 256 x U items) of a launch and the grid that walks them
long units_for(long items_per_row, int n_streams, int U) { return ((items_per_row + 256 * U - 1) / (256 * U)) * (long)n_streams; }
int grid_for(long units)
{
	int sms = 148, dev = 0;
	if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
	long cap = (long)sms * 8;  // eight 256-thread CTAs per SM, grid-stride beyond that
	long g = units < cap ? units : cap;
	return (int)(g < 1 ? 1 : g);
}

bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

}  // namespace

extern "C" {

int zen_pcm16_decode_mono_async(const int16_t* d_pcm, long pcm_stride, int channels, int n_streams, long n_frames, float* d_out,
                                long out_stride, void* cuda_stream)
{
	if (!d_pcm || !d_out || (channels != 1 && channels != 2) || n_streams < 1 || n_frames < 0 || pcm_stride < n_frames * channels
	    || out_stride < n_frames)
		return ZEN_ERR_ARG;
	if (zen_device_count() <= 0)
		return ZEN_ERR_CUDA;
	if (n_frames == 0)
		return ZEN_OK;
	cudaStream_t s = (cudaStream_t)cuda_stream;
	const long ipr = (n_frames + 7) / 8, total = units_for(ipr, n_streams, DEC_U);
	const bool vec = aligned16(d_pcm) && aligned16(d_out) && (pcm_stride % 8) == 0 && (out_stride % 4) == 0;
	const int g = grid_for(total);
	const int so = (double)n_streams * (double)n_frames * 4.0 > 96e6 ? 1 : 0;
	if (channels == 1) {
		if (vec) pcm16_decode_kernel<1, true><<<g, 256, 0, s>>>(d_pcm, pcm_stride, d_out, out_stride, n_frames, ipr, total, so);
		else pcm16_decode_kernel<1, false><<<g, 256, 0, s>>>(d_pcm, pcm_stride, d_out, out_stride, n_frames, ipr, total, so);
	}
	else {
		if (vec) pcm16_decode_kernel<2, true><<<g, 256, 0, s>>>(d_pcm, pcm_stride, d_out, out_stride, n_frames, ipr, total, so);
		else pcm16_decode_kernel<2, false><<<g, 256, 0, s>>>(d_pcm, pcm_stride, d_out, out_stride, n_frames, ipr, total, so);
	}
	ZEN_CUDA_CHECK(cudaGetLastError());
	return ZEN_OK;
}

int zen_pcm16_peaks_async(const float* d_in, long in_stride, int n_streams, long n, float* d_peaks, void* cuda_stream)
{
	if (!d_in || !d_peaks || n_streams < 1 || n < 0 || in_stride < n)
		return ZEN_ERR_ARG;
	if (zen_device_count() <= 0)
		return ZEN_ERR_CUDA;
	cudaStream_t s = (cudaStream_t)cuda_stream;
	ZEN_CUDA_CHECK(cudaMemsetAsync(d_peaks, 0, sizeof(float) * (size_t)n_streams, s));
	if (n == 0)
		return ZEN_OK;
	const long ipr = (n + 7) / 8, total = units_for(ipr, n_streams, PEAK_U);
	const bool vec = aligned16(d_in) && (in_stride % 4) == 0;
	const int g = grid_for(total);
	if (vec) peak_kernel<true><<<g, 256, 0, s>>>(d_in, in_stride, n, ipr, total, reinterpret_cast<unsigned*>(d_peaks));
	else peak_kernel<false><<<g, 256, 0, s>>>(d_in, in_stride, n, ipr, total, reinterpret_cast<unsigned*>(d_peaks));
	ZEN_CUDA_CHECK(cudaGetLastError());
	return ZEN_OK;
}

int zen_pcm16_encode_with_peaks_async(const float* d_in, long in_stride, int n_streams, long n, const float* d_peaks, int16_t* d_out,
                                      long out_stride, void* cuda_stream)
{
	if (!d_in || !d_out || !d_peaks || n_streams < 1 || n < 0 || in_stride < n || out_stride < n)
		return ZEN_ERR_ARG;
	if (zen_device_count() <= 0)
		return ZEN_ERR_CUDA;
	if (n == 0)
		return ZEN_OK;
	cudaStream_t s = (cudaStream_t)cuda_stream;
	const long ipr = (n + 7) / 8, total = units_for(ipr, n_streams, ENC_U);
	const bool vec = aligned16(d_in) && aligned16(d_out) && (in_stride % 4) == 0 && (out_stride % 8) == 0;
	const int g = grid_for(total);
	if (vec)
		pcm16_encode_kernel<true><<<g, 256, 0, s>>>(d_in, in_stride, n, reinterpret_cast<const unsigned*>(d_peaks), d_out, out_stride, ipr, total);
	else
		pcm16_encode_kernel<false><<<g, 256, 0, s>>>(d_in, in_stride, n, reinterpret_cast<const unsigned*>(d_peaks), d_out, out_stride, ipr, total);
	ZEN_CUDA_CHECK(cudaGetLastError());
	return ZEN_OK;
}

int zen_pcm16_encode_normalized_async(const float* d_in, long in_stride, int n_streams, long n, int16_t* d_out, long out_stride,
                                      float* d_peaks, void* cuda_stream)
{
	if (!d_out)
		return ZEN_ERR_ARG;
	int rc = zen_pcm16_peaks_async(d_in, in_stride, n_streams, n, d_peaks, cuda_stream);
	if (rc != ZEN_OK)
		return rc;
	return zen_pcm16_encode_with_peaks_async(d_in, in_stride, n_streams, n, d_peaks, d_out, out_stride, cuda_stream);
}

// the round-1 entry points: legacy default stream, synchronous
int zen_pcm16_decode_mono(const int16_t* d_pcm, long pcm_stride, int channels, int n_streams, long n_frames, float* d_out, long out_stride)
{
	int rc = zen_pcm16_decode_mono_async(d_pcm, pcm_stride, channels, n_streams, n_frames, d_out, out_stride, nullptr);
	if (rc != ZEN_OK)
		return rc;
	ZEN_CUDA_CHECK(cudaStreamSynchronize(nullptr));
	return ZEN_OK;
}

int zen_pcm16_encode_normalized(const float* d_in, long in_stride, int n_streams, long n, int16_t* d_out, long out_stride, float* d_peaks)
{
	int rc = zen_pcm16_encode_normalized_async(d_in, in_stride, n_streams, n, d_out, out_stride, d_peaks, nullptr);
	if (rc != ZEN_OK)
		return rc;
	ZEN_CUDA_CHECK(cudaStreamSynchronize(nullptr));
	return ZEN_OK;
}

}  // extern "C"
