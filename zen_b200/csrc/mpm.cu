// McLeod pitch method on device buffers: the documented consumer of HPRRealtime<GPU>'s harmonic output
// (reference: demos/pitch-tracking/pitch.cpp:40-135, pitch_detection.h:17-93, fed by
// demos/pitch-tracking/main.cu:90-107 from io.host_out after copy_harmonic).  Here it reads the harmonic hops where
// the HPR kernels leave them - device memory - one CTA per buffer of N samples, any number of buffers per launch:
//   real_autocorrelation (pitch.cpp:40-61): zero-padded 2N-point FFT (the library's shared-memory FFT, upper half of
//     the input pruned), out[i] *= conj(out[i]) / 2N for i < N ONLY - as the reference does -, inverse FFT, real part;
//   peak_picking (pitch.cpp:63-99), parabolic interpolation and the cut-offs of MPM::pitch (pitch.cpp:101-135),
//     restated as data-parallel steps: two find-first reductions for the start position, a block scan that numbers the
//     positive lobes, one 64-bit shared-memory atomicMax per candidate (value, then lowest position) for the key
//     maximum of each lobe, one thread per lobe for the interpolation, a block max and a block min for the cut-off.
// Results: the autocorrelation within FFT rounding of the reference's, the decisions identical given those values.
#include <cfloat>

#include "fft_smem.cuh"
#include "median_select.cuh"
#include "zen_common.cuh"

namespace zen_b200 {
const float2* fft_twiddle_table(int n);
}
using namespace zen_b200;

namespace {

template <int N>
struct MpmSmem {
	static constexpr int N2 = 2 * N;
	static constexpr size_t bytes() { return sizeof(float2) * (size_t)fpad_size(N2) + sizeof(float) * (size_t)(N + 4) + sizeof(unsigned long long) * (size_t)(N / 2 + 2); }
};

template <int N, int NT>
__global__ void __launch_bounds__(NT) mpm_kernel(const float* __restrict__ audio, long stride, int n_buffers, float sample_rate,
                                                const float2* __restrict__ tw, float* __restrict__ pitch_out, float* __restrict__ nsdf_out)
{
	constexpr int N2 = 2 * N;
	constexpr int PER = (N + NT - 1) / NT;   // consecutive positions per thread in the lobe scan
	extern __shared__ __align__(16) unsigned char mpm_raw[];
	float2* buf = reinterpret_cast<float2*>(mpm_raw);
	float* r = reinterpret_cast<float*>(buf + fpad_size(N2));
	unsigned long long* best = reinterpret_cast<unsigned long long*>(r + N + 4);
	__shared__ int s_first[2];
	__shared__ int s_warp[NT / 32 + 1];
	__shared__ unsigned s_hi;       // order-preserving key of the highest amplitude
	__shared__ int s_sel, s_nest;
	const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;

	for (int bidx = blockIdx.x; bidx < n_buffers; bidx += gridDim.x) {
		const float* x = audio + (size_t)bidx * stride;
		// ---- real_autocorrelation (pitch.cpp:40-61)
		for (int i = tid; i < N; i += NT)
			buf[i] = make_float2(x[i], 0.0f);   // out_im[N .. 2N) stays zero: pruned from the first stage
		__syncthreads();
		fft_smem<N2, NT, -1, 1, true, false, false>(buf, tw, tid);
		{
			const float sr = 1.0f / (float)(N * 2), si = 0.0f;   // scale = {1 / 2N, 0}
			for (int i = tid; i < N; i += NT) {                  // i < N only (pitch.cpp:50-52)
				const float a = buf[i].x, b = buf[i].y;
				// conj(out) * scale, then out * that: std::complex products (ac - bd, ad + bc), one rounding per operation
				const float cr = __fsub_rn(__fmul_rn(a, sr), __fmul_rn(-b, si)), ci = __fadd_rn(__fmul_rn(a, si), __fmul_rn(-b, sr));
				buf[i] = make_float2(__fsub_rn(__fmul_rn(a, cr), __fmul_rn(b, ci)), __fadd_rn(__fmul_rn(a, ci), __fmul_rn(b, cr)));
			}
		}
		__syncthreads();
		fft_smem<N2, NT, +1, 1, false, true, false>(buf, tw, tid);
		for (int i = tid; i < N; i += NT) {
			const float v = buf[i].x;
			r[i] = v;
			if (nsdf_out) nsdf_out[(size_t)bidx * N + i] = v;
		}
		if (tid == 0) {
			s_first[0] = (N - 1) / 3;
			s_first[1] = N - 1;
			s_hi = 0u;
			s_sel = 0x7fffffff;
			s_nest = 0;
		}
		for (int i = tid; i < N / 2 + 2; i += NT)
			best[i] = 0ull;
		__syncthreads();
		// ---- peak_picking (pitch.cpp:63-99).  Start position: skip the leading positive run (at most a third of the
		// buffer), then the non-positive run behind it.
		for (int i = tid; i < (N - 1) / 3; i += NT)
			if (!(r[i] > 0.0f)) atomicMin(&s_first[0], i);
		__syncthreads();
		const int pos_a = s_first[0];
		for (int i = pos_a + tid; i < N - 1; i += NT)
			if (!(r[i] <= 0.0f)) atomicMin(&s_first[1], i);
		__syncthreads();
		const int pos0 = s_first[1] == 0 ? 1 : s_first[1];
		// The sequential loop examines pos0 and every later position p <= N-2 with !(r[p] <= 0); a run of non-positive
		// values closes the current lobe.  Lobe numbers: inclusive scan of "p starts a lobe".
		auto examined = [&](int p) { return p == pos0 || (p > pos0 && p <= N - 2 && !(r[p] <= 0.0f)); };
		int cnt = 0;
		unsigned startbits = 0u;
#pragma unroll
		for (int c = 0; c < PER; ++c) {
			const int p = tid * PER + c;
			const bool st = p >= pos0 && p <= N - 2 && examined(p) && (p == pos0 || !examined(p - 1));
			startbits |= (st ? 1u : 0u) << c;
			cnt += st ? 1 : 0;
		}
		int incl = cnt;
		for (int d = 1; d < 32; d <<= 1) {
			const int v = __shfl_up_sync(0xffffffffu, incl, d);
			if (lane >= d) incl += v;
		}
		if (lane == 31) s_warp[wid] = incl;
		__syncthreads();
		if (wid == 0) {
			int v = lane < NT / 32 ? s_warp[lane] : 0;
			for (int d = 1; d < 32; d <<= 1) {
				const int u = __shfl_up_sync(0xffffffffu, v, d);
				if (lane >= d) v += u;
			}
			if (lane < NT / 32) s_warp[lane] = v;   // inclusive over warps
		}
		__syncthreads();
		int lobe = incl - cnt + (wid > 0 ? s_warp[wid - 1] : 0);   // lobes started before this thread's first position
#pragma unroll
		for (int c = 0; c < PER; ++c) {
			const int p = tid * PER + c;
			if (startbits & (1u << c)) ++lobe;
			if (p >= pos0 && p <= N - 2 && examined(p)) {
				// key maximum of the lobe: highest candidate, the first one among equals (strict > in pitch.cpp:83-85)
				if (r[p] > r[p - 1] && r[p] >= r[p + 1]) {
					const unsigned long long key = ((unsigned long long)f2key(r[p]) << 32) | (unsigned long long)(0xffffffffu - (unsigned)p);
					atomicMax(&best[lobe - 1], key);
				}
			}
		}
		__syncthreads();
		const int n_lobes = s_warp[NT / 32 - 1];
		// ---- MPM::pitch (pitch.cpp:101-135): one thread per lobe
		for (int l = tid; l < n_lobes; l += NT) {
			const unsigned long long key = best[l];
			if (key == 0ull) continue;   // a lobe without a local maximum contributes nothing
			const int i = (int)(0xffffffffu - (unsigned)(key & 0xffffffffull));
			const float v = r[i];
			atomicMax(&s_hi, f2key(v));
			if (v > 0.5f) {
				// parabolic_interpolation (pitch.cpp:16-38); 1 <= i <= N - 2 here, so always the three-point branch
				const float den = __fsub_rn(__fadd_rn(r[i + 1], r[i - 1]), __fmul_rn(2.0f, r[i]));
				const float delta = __fsub_rn(r[i - 1], r[i + 1]);
				float px = (float)i, py = v;
				if (den != 0.0f) {
					px = __fadd_rn((float)i, __fdiv_rn(delta, __fmul_rn(2.0f, den)));
					py = __fsub_rn(v, __fdiv_rn(__fmul_rn(delta, delta), __fmul_rn(8.0f, den)));
				}
				atomicMax(&s_hi, f2key(py));
				atomicAdd(&s_nest, 1);
				// keep the estimate of this lobe where the lobe's key was (x in the high word, y in the low word)
				best[l] = ((unsigned long long)__float_as_uint(px) << 32) | (unsigned long long)__float_as_uint(py);
			}
			else {
				best[l] = 0ull;   // no estimate from this lobe
			}
		}
		__syncthreads();
		// first estimate, in lobe order, that reaches 0.93 of the highest amplitude
		const float highest = key2f(s_hi);
		const float cutoff = (float)(0.93 * (double)highest);
		for (int l = tid; l < n_lobes; l += NT) {
			const unsigned long long e = best[l];
			if (e == 0ull) continue;
			const float py = __uint_as_float((unsigned)(e & 0xffffffffull));
			if (py >= cutoff) atomicMin(&s_sel, l);
		}
		__syncthreads();
		if (tid == 0) {
			float result = -1.0f;
			if (s_nest > 0) {
				float period = 0.0f;
				if (s_sel != 0x7fffffff) period = __uint_as_float((unsigned)(best[s_sel] >> 32));
				const float est = __fdiv_rn(sample_rate, period);
				result = est > 80.0f ? est : -1.0f;
			}
			pitch_out[bidx] = result;
		}
		__syncthreads();
	}
}

template <int N>
int launch_mpm(const float* d_audio, long stride, int n_buffers, float fs, float* d_pitch, float* d_nsdf, cudaStream_t s)
{
	constexpr int NT = (2 * N / 16) < 64 ? 64 : ((2 * N / 16) > 512 ? 512 : (2 * N / 16));
	const float2* tw = fft_twiddle_table(2 * N);
	if (!tw)
		return ZEN_ERR_CUDA;
	auto kern = mpm_kernel<N, NT>;
	const size_t smem = MpmSmem<N>::bytes();
	ZEN_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	int sms = 148, dev = 0;
	if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
	const int grid = n_buffers < 4 * sms ? n_buffers : 4 * sms;
	kern<<<grid, NT, smem, s>>>(d_audio, stride, n_buffers, fs, tw, d_pitch, d_nsdf);
	ZEN_CUDA_CHECK(cudaGetLastError());
	return ZEN_OK;
}

}  // namespace

extern "C" int zen_mpm_pitch(int n, float sample_rate, const float* d_audio, long stride, int n_buffers, float* d_pitch, float* d_nsdf,
                             void* cuda_stream)
{
	if (!d_audio || !d_pitch || n_buffers < 1 || stride < 0)
		return ZEN_ERR_ARG;
	if (zen_device_count() <= 0)
		return ZEN_ERR_CUDA;
	cudaStream_t s = (cudaStream_t)cuda_stream;
	switch (n) {
	case 256: return launch_mpm<256>(d_audio, stride, n_buffers, sample_rate, d_pitch, d_nsdf, s);
	case 512: return launch_mpm<512>(d_audio, stride, n_buffers, sample_rate, d_pitch, d_nsdf, s);
	case 1024: return launch_mpm<1024>(d_audio, stride, n_buffers, sample_rate, d_pitch, d_nsdf, s);
	case 2048: return launch_mpm<2048>(d_audio, stride, n_buffers, sample_rate, d_pitch, d_nsdf, s);
	case 4096: return launch_mpm<4096>(d_audio, stride, n_buffers, sample_rate, d_pitch, d_nsdf, s);
	}
	return ZEN_ERR_UNSUPPORTED;
}
