// Onset detection function on device buffers: the documented consumer of HPRRealtime<GPU>'s percussive output
// (reference: demos/beat-tracking/OnsetDetection.cpp:60-131 "complex spectral difference,
// half-wave rectified", Window.h:31-40, fed by demos/beat-tracking/main.cu:92-118 one 256-sample hop at a time from
// io.host_out after copy_percussive).  Here it reads the percussive hops where the HPR kernels leave them - device
// memory - for any number of streams per launch, and leaves one ODF sample per hop for the beat tracker.
//
// What the reference computes, bugs included.  calculate_sample() shifts its 512-sample frame back by one hop and
// appends the new hop; perform_FFT() then swaps the two halves and windows them IN PLACE, in the frame itself.  The
// next shift therefore copies a windowed, swapped half - and what it copies is the half that was shifted in from the
// initial zeros: by induction the older half of the frame is zero at every call.  The transform the reference takes is
//     x_t[i] * w[i + 256], i < 256, followed by 256 zeros            (the new hop under the FALLING half of the Hann window)
// so a frame depends on its own hop only, and the ODF sample on the phases of the last three frames and the magnitudes of
// the last two.  The oracle restates the literal in-place operations (oracle/hpr_oracle.c: zo_onset_csd); this kernel
// uses the consequence: the zero half is pruned from the first FFT stage, and a CTA can start anywhere in a stream after
// re-analysing two hops.  All 512 bins enter the sum, as in the reference (the upper half mirrors the lower one).
#include "fft_smem.cuh"
#include "zen_common.cuh"

namespace zen_b200 {
const float2* fft_twiddle_table(int n);
}
using namespace zen_b200;

namespace {

constexpr int ODF_N = 512, ODF_HOP = 256, ODF_NT = 64, ODF_PER = ODF_N / ODF_NT;

__global__ void __launch_bounds__(ODF_NT) onset_csd_kernel(const float* __restrict__ audio, long stride, long n_hops, int tile_hops, int n_tiles,
                                                         int total_items, const float* __restrict__ win, const float2* __restrict__ tw,
                                                         float* __restrict__ odf, long odf_stride)
{
	__shared__ __align__(16) float2 buf[fpad_size(ODF_N)];
	__shared__ float s_part[ODF_NT / 32];
	const int tid = threadIdx.x;
	for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
		const int stream = item / n_tiles, tile = item - stream * n_tiles;
		const long h0 = (long)tile * tile_hops, h1 = min(n_hops, h0 + (long)tile_hops);
		const float* x = audio + (size_t)stream * stride;
		float pphase[ODF_PER], pphase2[ODF_PER], pmag[ODF_PER];   // prevPhase, prevPhase2, prevMagSpec of this thread's bins
#pragma unroll
		for (int j = 0; j < ODF_PER; ++j)
			pphase[j] = pphase2[j] = pmag[j] = 0.0f;
		for (long h = max(0L, h0 - 2); h < h1; ++h) {
			__syncthreads();   // the previous frame's spectrum has been read
			for (int i = tid; i < ODF_HOP; i += ODF_NT)
				buf[i] = make_float2(x[(size_t)h * ODF_HOP + i] * win[i + ODF_HOP], 0.0f);
			__syncthreads();
			fft_smem<ODF_N, ODF_NT, -1, 1, true, false, false>(buf, tw, tid);
			float sum = 0.0f;
#pragma unroll
			for (int j = 0; j < ODF_PER; ++j) {
				const float2 X = buf[tid + j * ODF_NT];
				const float phase = atan2f(X.y, X.x);
				const float mag = sqrtf(X.x * X.x + X.y * X.y);
				const float dev = phase - 2.0f * pphase[j] + pphase2[j];
				if (mag - pmag[j] > 0.0f)
					sum += sqrtf(mag * mag + pmag[j] * pmag[j] - 2.0f * mag * pmag[j] * cosf(dev));
				pphase2[j] = pphase[j];
				pphase[j] = phase;
				pmag[j] = mag;
			}
			for (int s = 16; s > 0; s >>= 1)
				sum += __shfl_xor_sync(0xffffffffu, sum, s);
			if ((tid & 31) == 0) s_part[tid >> 5] = sum;
			__syncthreads();
			if (tid == 0 && h >= h0) {
				float t = 0.0f;
#pragma unroll
				for (int w = 0; w < ODF_NT / 32; ++w)
					t += s_part[w];
				odf[(size_t)stream * odf_stride + h] = t;
			}
		}
	}
}

// Hann window as Window.h:31-40 writes it: 0.5 (1 - cos(2 PI (n / (N - 1)))), float arithmetic, PI = 3.14159265359F
// (the reference evaluates the cosine at compile time with gcem; libm's cosf agrees with it to an ulp)
const float* odf_window()
{
	static const float* tables[16] = {};
	int dev = 0;
	if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16)
		return nullptr;
	if (!tables[dev]) {
		float h[ODF_N];
		const float PI = 3.14159265359F, N = (float)(ODF_N - 1);
		for (int n = 0; n < ODF_N; ++n)
			h[n] = 0.5F * (1.0F - cosf(2.0F * PI * ((float)n / N)));
		float* d = nullptr;
		if (cudaMalloc(&d, sizeof(h)) != cudaSuccess || cudaMemcpy(d, h, sizeof(h), cudaMemcpyHostToDevice) != cudaSuccess)
			return nullptr;
		tables[dev] = d;
	}
	return tables[dev];
}

}  // namespace

extern "C" int zen_onset_csd(const float* d_audio, long stride, int n_streams, long n_hops, float* d_odf, long odf_stride, void* cuda_stream)
{
	if (!d_audio || !d_odf || n_streams < 1 || n_hops < 0 || stride < n_hops * ODF_HOP || odf_stride < n_hops)
		return ZEN_ERR_ARG;
	if (zen_device_count() <= 0)
		return ZEN_ERR_CUDA;
	if (n_hops == 0)
		return ZEN_OK;
	const float2* tw = fft_twiddle_table(ODF_N);
	const float* win = odf_window();
	if (!tw || !win)
		return ZEN_ERR_CUDA;
	int sms = 148, dev = 0;
	if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
	// tiles: enough (stream, tile) items to fill the GPU a few times over, long enough that the two re-analysed hops
	// in front of a tile stay a small share
	long tiles = ((long)sms * 64 + n_streams - 1) / n_streams;
	long tile_hops = (n_hops + tiles - 1) / (tiles < 1 ? 1 : tiles);
	if (tile_hops < 32) tile_hops = 32;
	if (tile_hops > n_hops) tile_hops = n_hops;
	const int n_tiles = (int)((n_hops + tile_hops - 1) / tile_hops);
	const long total = (long)n_tiles * n_streams;
	if (total > 0x7fffffffL)
		return ZEN_ERR_UNSUPPORTED;
	const long cap = (long)sms * 16;
	const int grid = (int)(total < cap ? total : cap);
	onset_csd_kernel<<<grid, ODF_NT, 0, (cudaStream_t)cuda_stream>>>(d_audio, stride, n_hops, (int)tile_hops, n_tiles, (int)total, win, tw, d_odf,
	                                                               odf_stride);
	ZEN_CUDA_CHECK(cudaGetLastError());
	return ZEN_OK;
}
