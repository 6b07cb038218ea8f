// Shared-memory Stockham FFT (radix 8/4/2, autosort, in place through registers)
// used by every kernel in this library.  Replaces cuFFT as driven by
// FFTC2CWrapperGPU (libzen/fftw.h:20-49): unnormalised in both directions.
//
// Layout: complex values as float2 in shared memory, index padded by one slot
// every 16 (fpad) so the stride-R stores of the early stages spread over banks.
#pragma once
#include <cmath>
#include <cuda_runtime.h>

namespace zen_b200 {

__device__ __forceinline__ int fpad(int i) { return i + (i >> 4); }
__host__ __device__ constexpr int fpad_size(int n) { return n + (n >> 4) + 1; }

__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmul(float2 a, float2 b)
{
	return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 cconj(float2 a) { return make_float2(a.x, -a.y); }
// multiply by S*i  (S = -1 forward, +1 inverse)
template <int S>
__device__ __forceinline__ float2 mul_si(float2 a)
{
	return S > 0 ? make_float2(-a.y, a.x) : make_float2(a.y, -a.x);
}

template <int S>
__device__ __forceinline__ void dft2(float2* v)
{
	float2 t = v[0];
	v[0] = cadd(t, v[1]);
	v[1] = csub(t, v[1]);
}

template <int S>
__device__ __forceinline__ void dft4(float2* v)
{
	float2 a0 = cadd(v[0], v[2]), a1 = csub(v[0], v[2]);
	float2 a2 = cadd(v[1], v[3]), a3 = mul_si<S>(csub(v[1], v[3]));
	v[0] = cadd(a0, a2);
	v[2] = csub(a0, a2);
	v[1] = cadd(a1, a3);
	v[3] = csub(a1, a3);
}

template <int S>
__device__ __forceinline__ void dft8(float2* v)
{
	float2 e[4] = {v[0], v[2], v[4], v[6]};
	float2 o[4] = {v[1], v[3], v[5], v[7]};
	dft4<S>(e);
	dft4<S>(o);
	const float h = 0.70710678118654752440f;
	const float s = (float)S;
	float2 o1 = make_float2(h * (o[1].x - s * o[1].y), h * (s * o[1].x + o[1].y));
	float2 o2 = mul_si<S>(o[2]);
	float2 o3 = make_float2(h * (-o[3].x - s * o[3].y), h * (s * o[3].x - o[3].y));
	v[0] = cadd(e[0], o[0]);
	v[4] = csub(e[0], o[0]);
	v[1] = cadd(e[1], o1);
	v[5] = csub(e[1], o1);
	v[2] = cadd(e[2], o2);
	v[6] = csub(e[2], o2);
	v[3] = cadd(e[3], o3);
	v[7] = csub(e[3], o3);
}

template <int S, int R>
__device__ __forceinline__ void dftR(float2* v)
{
	if constexpr (R == 8)
		dft8<S>(v);
	else if constexpr (R == 4)
		dft4<S>(v);
	else
		dft2<S>(v);
}

// Per-stage twiddle tables.  Stage (NS, R) needs w^(r*k) for r = 1..R-1, k = 0..NS-1 with
// w = exp(-2*pi*i / (NS*R)).  They are stored stage after stage as tab[(r-1)*NS + k], so the 32 lanes
// of a warp (consecutive butterflies j, k = j mod NS) read consecutive addresses instead of gathering
// from one big exp(-2*pi*i*t/M) table.  Stage NS = 1 needs none.
template <int M, int NS = 1>
constexpr int fft_twiddle_count()
{
	if constexpr (NS >= M) {
		return 0;
	}
	else {
		constexpr int rem = M / NS;
		constexpr int R = (rem % 8 == 0 && rem != 16) ? 8 : ((rem % 4 == 0) ? 4 : 2);
		return (NS > 1 ? (R - 1) * NS : 0) + fft_twiddle_count<M, NS * R>();
	}
}

// host: fill the table in the order the stages consume it (double precision -> float)
template <typename F2>
inline void fft_fill_twiddles(int M, F2* out)
{
	const double two_pi = 6.283185307179586476925286766559;
	int NS = 1, pos = 0;
	while (NS < M) {
		int rem = M / NS;
		int R = (rem % 8 == 0 && rem != 16) ? 8 : ((rem % 4 == 0) ? 4 : 2);
		if (NS > 1)
			for (int r = 1; r < R; ++r)
				for (int k = 0; k < NS; ++k) {
					double a = -two_pi * (double)r * (double)k / ((double)NS * (double)R);
					out[pos].x = (float)std::cos(a);
					out[pos].y = (float)std::sin(a);
					++pos;
				}
		NS *= R;
	}
}
inline int fft_twiddle_count_rt(int M)
{
	int NS = 1, n = 0;
	while (NS < M) {
		int rem = M / NS;
		int R = (rem % 8 == 0 && rem != 16) ? 8 : ((rem % 4 == 0) ? 4 : 2);
		if (NS > 1) n += (R - 1) * NS;
		NS *= R;
	}
	return n;
}

// One Stockham decimation-in-time stage of radix R on an M-point transform
// whose already-combined sub-transforms have length NS.  tw points at this stage's table.
// ZIN : the upper half of the input is known to be zero (zero-padded frame): those loads and the butterfly
//       arithmetic that depends on them disappear (first stage only).
// HOUT: only the lower half of the output is needed: the stores of the upper half and the arithmetic feeding
//       them disappear (last stage only).
// GT  : tw may point into shared memory (the resident real-time kernel keeps its tables there): plain generic
//       loads instead of the read-only global path.
template <int M, int NT, int S, int R, int NS, bool ZIN = false, bool HOUT = false, bool GT = false>
__device__ __forceinline__ void fft_stage(float2* buf, const float2* __restrict__ tw, int tid)
{
	constexpr int NB = M / R;
	constexpr int PER = (NB + NT - 1) / NT;
	float2 v[PER][R];
#pragma unroll
	for (int b = 0; b < PER; ++b) {
		int j = tid + b * NT;
		if ((NB % NT == 0) || j < NB) {
			int k = j & (NS - 1);
			float2 w[R];
			if constexpr (NS > 1) {
#pragma unroll
				for (int r = 1; r < R; ++r)
					w[r] = GT ? tw[(r - 1) * NS + k] : __ldg(&tw[(r - 1) * NS + k]);
			}
#pragma unroll
			for (int r = 0; r < R; ++r) {
				if (ZIN && r >= R / 2)
					v[b][r] = make_float2(0.0f, 0.0f);
				else
					v[b][r] = buf[fpad(j + r * NB)];
			}
			if constexpr (NS > 1) {
#pragma unroll
				for (int r = 1; r < R; ++r) {
					if (S > 0)
						w[r].y = -w[r].y;
					v[b][r] = cmul(v[b][r], w[r]);
				}
			}
			dftR<S, R>(v[b]);
		}
	}
	__syncthreads();
#pragma unroll
	for (int b = 0; b < PER; ++b) {
		int j = tid + b * NT;
		if ((NB % NT == 0) || j < NB) {
			int k = j & (NS - 1);
			int j0 = (j - k) * R + k;
#pragma unroll
			for (int r = 0; r < R; ++r)
				if (!(HOUT && r >= R / 2))
					buf[fpad(j0 + r * NS)] = v[b][r];
		}
	}
	__syncthreads();
}

// M-point complex FFT in shared memory, all NT threads of the CTA participate.
// tw: the per-stage tables of fft_fill_twiddles(M).  The caller must have
// synchronised after filling buf.  Ends synchronised.
template <int M, int NT, int S, int NS = 1, bool ZIN = false, bool HOUT = false, bool GT = false>
__device__ __forceinline__ void fft_smem(float2* buf, const float2* __restrict__ tw, int tid)
{
	if constexpr (NS < M) {
		constexpr int rem = M / NS;
		constexpr int R = (rem % 8 == 0 && rem != 16) ? 8 : ((rem % 4 == 0) ? 4 : 2);
		fft_stage<M, NT, S, R, NS, (ZIN && NS == 1), (HOUT && NS * R == M), GT>(buf, tw, tid);
		fft_smem<M, NT, S, NS * R, ZIN, HOUT, GT>(buf, tw + (NS > 1 ? (R - 1) * NS : 0), tid);
	}
}

// ---- out-of-place (ping-pong) variant --------------------------------------
// Each stage reads src and writes dst, so one barrier per stage suffices (the in-place stage needs one between its
// loads and its stores as well).  Used where latency, not shared-memory footprint, matters: the resident real-time
// kernel.  fft_smem_pp returns the buffer that holds the result (a after an even number of stages, else b).
template <int M, int NS = 1>
constexpr int fft_stage_count()
{
	if constexpr (NS >= M) {
		return 0;
	}
	else {
		constexpr int rem = M / NS;
		constexpr int R = (rem % 8 == 0 && rem != 16) ? 8 : ((rem % 4 == 0) ? 4 : 2);
		return 1 + fft_stage_count<M, NS * R>();
	}
}

template <int M, int NT, int S, int R, int NS, bool ZIN = false, bool HOUT = false, bool GT = false>
__device__ __forceinline__ void fft_stage_pp(const float2* __restrict__ src, float2* __restrict__ dst, const float2* __restrict__ tw, int tid)
{
	constexpr int NB = M / R;
	constexpr int PER = (NB + NT - 1) / NT;
#pragma unroll
	for (int b = 0; b < PER; ++b) {
		int j = tid + b * NT;
		if ((NB % NT == 0) || j < NB) {
			float2 v[R];
			int k = j & (NS - 1);
			float2 w[R];
			if constexpr (NS > 1) {
#pragma unroll
				for (int r = 1; r < R; ++r)
					w[r] = GT ? tw[(r - 1) * NS + k] : __ldg(&tw[(r - 1) * NS + k]);
			}
#pragma unroll
			for (int r = 0; r < R; ++r) {
				if (ZIN && r >= R / 2)
					v[r] = make_float2(0.0f, 0.0f);
				else
					v[r] = src[fpad(j + r * NB)];
			}
			if constexpr (NS > 1) {
#pragma unroll
				for (int r = 1; r < R; ++r) {
					if (S > 0)
						w[r].y = -w[r].y;
					v[r] = cmul(v[r], w[r]);
				}
			}
			dftR<S, R>(v);
			int j0 = (j - k) * R + k;
#pragma unroll
			for (int r = 0; r < R; ++r)
				if (!(HOUT && r >= R / 2))
					dst[fpad(j0 + r * NS)] = v[r];
		}
	}
	__syncthreads();
}

template <int M, int NT, int S, int NS = 1, bool ZIN = false, bool HOUT = false, bool GT = false>
__device__ __forceinline__ float2* fft_smem_pp(float2* a, float2* b, const float2* __restrict__ tw, int tid)
{
	if constexpr (NS < M) {
		constexpr int rem = M / NS;
		constexpr int R = (rem % 8 == 0 && rem != 16) ? 8 : ((rem % 4 == 0) ? 4 : 2);
		fft_stage_pp<M, NT, S, R, NS, (ZIN && NS == 1), (HOUT && NS * R == M), GT>(a, b, tw, tid);
		return fft_smem_pp<M, NT, S, NS * R, ZIN, HOUT, GT>(b, a, tw + (NS > 1 ? (R - 1) * NS : 0), tid);
	}
	else {
		return a;
	}
}

}  // namespace zen_b200
