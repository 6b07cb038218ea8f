// Shared-memory Stockham FFT (radix 8/4/2, autosort) used by every kernel in this library.  Replaces cuFFT as
// driven by FFTC2CWrapperGPU (libzen/fftw.h:20-49): unnormalised in both directions.
//
// Arithmetic: complex values are float2 and every complex add / subtract / twiddle multiply is ONE or TWO packed
// fp32x2 instructions of sm_100 (add.rn.f32x2 / mul.rn.f32x2 / fma.rn.f32x2 -> FADD2 / FMUL2 / FFMA2).  The SASS
// forms take a swapped (.LO_HI), half-negated (.NP / .PN) or broadcast (.F32) operand for free, so multiplying by
// +-i, conjugating a twiddle and broadcasting a real part cost nothing: a radix-8 butterfly is 26 packed
// instructions instead of ~60 scalar ones, a twiddle multiply 2 instead of 4.
//
// Layout: float2 in shared memory.  The transform's input and output are in NATURAL order (element i at i): every
// access pattern outside the stages is unit-stride across lanes.  Between two stages the layout is chosen by the
// stage that WRITES, so that its stride-NS stores spread over the banks (the loads of every stage are unit-stride):
//   after the NS = 1 stage (lane j writes 8 j + r):            one slot of padding every 16   (LAY_F)
//   after the NS = 8 radix-8 stage (64 q + k + 8 r, k < 8):    eight slots every 64           (LAY_L2, M >= 1024)
//   after any stage with NS >= 16, and after the last stage:   natural                        (LAY_N)
// Buffers hold fpad_size(M) = M + M/8 + 1 elements.  Load / store positions are one base plus compile-time offsets.
//
// Every variant below (in place, ping-pong, first stage fed by a functor, last stage drained by a functor) runs
// the SAME butterfly code, so their results are bit-identical: the batched, the per-launch and the resident
// real-time kernels may pick whichever suits them.
#pragma once
#include <cmath>
#include <cuda_runtime.h>

namespace zen_b200 {

__host__ __device__ constexpr int fpad_size(int n) { return (n + (n >> 3) + 2) & ~1; }  // even: what follows stays 16-byte aligned

enum { LAY_N = 0, LAY_F = 1, LAY_L2 = 2 };
template <int LAY>
__device__ __forceinline__ int lay_pos(int i)
{
	if constexpr (LAY == LAY_F) return i + (i >> 4);
	else if constexpr (LAY == LAY_L2) return i + 8 * (i >> 6);
	else return i;
}
// position of element base + r * STRIDE given the position `base_p` of `base`
template <int LAY, int STRIDE>
__device__ __forceinline__ int lay_step(int base, int base_p, int r)
{
	if constexpr (LAY == LAY_N) return base_p + r * STRIDE;
	else if constexpr (LAY == LAY_F && STRIDE % 16 == 0) return base_p + r * (STRIDE + STRIDE / 16);
	else if constexpr (LAY == LAY_L2 && STRIDE % 64 == 0) return base_p + r * (STRIDE + STRIDE / 8);
	else return lay_pos<LAY>(base + r * STRIDE);
}
// layout of what the stage (NS, R) of an M-point transform writes
template <int M, int NS, int R>
constexpr int fft_layout_out()
{
	return (NS * R >= M) ? LAY_N : (NS == 1 ? LAY_F : ((NS == 8 && R == 8 && M >= 1024) ? LAY_L2 : (NS < 16 ? LAY_F : LAY_N)));
}

// ---- packed fp32x2 ----------------------------------------------------------------------------------------
typedef unsigned long long f32x2_t;
__device__ __forceinline__ f32x2_t pk(float lo, float hi)
{
	f32x2_t r;
	asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
	return r;
}
__device__ __forceinline__ f32x2_t pk(float2 a) { return pk(a.x, a.y); }
__device__ __forceinline__ float2 up(f32x2_t v)
{
	float2 r;
	asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
	return r;
}
__device__ __forceinline__ f32x2_t padd(f32x2_t a, f32x2_t b)
{
	f32x2_t d;
	asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
	return d;
}
__device__ __forceinline__ f32x2_t psub(f32x2_t a, f32x2_t b)
{
	f32x2_t d;
	asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
	return d;
}
__device__ __forceinline__ f32x2_t pmul(f32x2_t a, f32x2_t b)
{
	f32x2_t d;
	asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
	return d;
}
__device__ __forceinline__ f32x2_t pfma(f32x2_t a, f32x2_t b, f32x2_t c)
{
	f32x2_t d;
	asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
	return d;
}

__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return up(padd(pk(a), pk(b))); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return up(psub(pk(a), pk(b))); }
__device__ __forceinline__ float2 cconj(float2 a) { return make_float2(a.x, -a.y); }
// a * w:  (ax wx - ay wy, ax wy + ay wx) = ax * (wx, wy) + ay * (-wy, wx)
__device__ __forceinline__ float2 cmul(float2 a, float2 w)
{
	return up(pfma(pk(a.y, a.y), pk(-w.y, w.x), pmul(pk(a.x, a.x), pk(w.x, w.y))));
}
// a * conj(w)
__device__ __forceinline__ float2 cmulc(float2 a, float2 w)
{
	return up(pfma(pk(a.y, a.y), pk(w.y, w.x), pmul(pk(a.x, a.x), pk(w.x, -w.y))));
}
// a * s (both parts)
__device__ __forceinline__ float2 cscale(float2 a, float s) { return up(pmul(pk(a), pk(s, s))); }
// multiply by S*i  (S = -1 forward, +1 inverse): an operand modifier once it feeds a packed instruction
template <int S>
__device__ __forceinline__ float2 mul_si(float2 a)
{
	return S > 0 ? make_float2(-a.y, a.x) : make_float2(a.y, -a.x);
}
// a + s * b (both parts scaled by the same real s)
__device__ __forceinline__ float2 caxpy(float2 a, float s, float2 b) { return up(pfma(pk(b), pk(s, s), pk(a))); }

template <int S>
__device__ __forceinline__ void dft2(float2* v)
{
	const float2 t = v[0];
	v[0] = cadd(t, v[1]);
	v[1] = csub(t, v[1]);
}

template <int S>
__device__ __forceinline__ void dft4(float2* v)
{
	const float2 a0 = cadd(v[0], v[2]), a1 = csub(v[0], v[2]);
	const float2 a2 = cadd(v[1], v[3]), a3 = mul_si<S>(csub(v[1], v[3]));
	v[0] = cadd(a0, a2);
	v[2] = csub(a0, a2);
	v[1] = cadd(a1, a3);
	v[3] = csub(a1, a3);
}

// second half of the radix-8 butterfly: e = DFT4 of the even inputs, o = DFT4 of the odd ones
template <int S>
__device__ __forceinline__ void dft8_combine(float2* v, const float2* e, const float2* o)
{
	const float h = 0.70710678118654752440f;
	// o1 * w8 = h * (o1 + S i o1) ; o2 * w8^2 = S i o2 ; o3 * w8^3 = h * (S i o3 - o3)
	const float2 t1 = cadd(o[1], mul_si<S>(o[1]));
	const float2 t3 = csub(mul_si<S>(o[3]), o[3]);
	const float2 o2 = mul_si<S>(o[2]);
	v[0] = cadd(e[0], o[0]);
	v[4] = csub(e[0], o[0]);
	v[1] = caxpy(e[1], h, t1);
	v[5] = caxpy(e[1], -h, t1);
	v[2] = cadd(e[2], o2);
	v[6] = csub(e[2], o2);
	v[3] = caxpy(e[3], h, t3);
	v[7] = caxpy(e[3], -h, t3);
}

template <int S>
__device__ __forceinline__ void dft8(float2* v)
{
	float2 e[4] = {v[0], v[2], v[4], v[6]};
	float2 o[4] = {v[1], v[3], v[5], v[7]};
	dft4<S>(e);
	dft4<S>(o);
	dft8_combine<S>(v, e, o);
}

// radix-8 butterfly whose inputs 4..7 are zero (first stage of a zero-padded frame): the two inner DFT4 shrink to
// four packed operations each
template <int S>
__device__ __forceinline__ void dft8_zin(float2* v)
{
	float2 e[4], o[4];
	const float2 r2 = mul_si<S>(v[2]), r3 = mul_si<S>(v[3]);
	e[0] = cadd(v[0], v[2]);
	e[2] = csub(v[0], v[2]);
	e[1] = cadd(v[0], r2);
	e[3] = csub(v[0], r2);
	o[0] = cadd(v[1], v[3]);
	o[2] = csub(v[1], v[3]);
	o[1] = cadd(v[1], r3);
	o[3] = csub(v[1], r3);
	dft8_combine<S>(v, e, o);
}

template <int S, int R, bool ZIN = false>
__device__ __forceinline__ void dftR(float2* v)
{
	if constexpr (R == 8) {
		if constexpr (ZIN)
			dft8_zin<S>(v);
		else
			dft8<S>(v);
	}
	else if constexpr (R == 4)
		dft4<S>(v);
	else
		dft2<S>(v);
}

// radix of the stage that follows sub-transforms of length NS in an M-point transform
template <int M, int NS>
constexpr int fft_radix()
{
	constexpr int rem = M / NS;
	return (rem % 8 == 0 && rem != 16) ? 8 : ((rem % 4 == 0) ? 4 : 2);
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
		constexpr int R = fft_radix<M, NS>();
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

template <int M, int NS = 1>
constexpr int fft_stage_count()
{
	if constexpr (NS >= M) {
		return 0;
	}
	else {
		return 1 + fft_stage_count<M, NS * fft_radix<M, NS>()>();
	}
}

// shared-memory source / sink of a stage (the default I/O); `p` is the position in the layout at hand
struct SmemLoad {
	const float2* buf;
	__device__ __forceinline__ float2 operator()(int /*j*/, int /*r*/, int p) const { return buf[p]; }
};
struct SmemStore {
	float2* buf;
	__device__ __forceinline__ void operator()(int /*idx*/, int p, float2 v) const { buf[p] = v; }
};

// One Stockham decimation-in-time stage of radix R on an M-point transform whose already-combined sub-transforms
// have length NS.  tw points at this stage's table.  Element j + r*NB comes from ld(j, r, padded position), element
// j0 + r*NS goes to st(index, padded position, value).
// ZIN : the upper half of the input is known to be zero (zero-padded frame): those loads and the butterfly
//       arithmetic that depends on them disappear (first stage only).
// HOUT: only the lower half of the output is needed: the stores of the upper half and the arithmetic feeding
//       them disappear (last stage only; the packed operations are plain asm, dead ones are dropped).
// GT  : tw may point into shared memory (the resident real-time kernel keeps its tables there): plain generic
//       loads instead of the read-only global path.
// MID : barrier between the loads and the stores (needed when the stage works in place).
template <int M, int NT, int S, int R, int NS, int LIN, bool ZIN, bool HOUT, bool GT, bool MID, typename LD, typename ST>
__device__ __forceinline__ void fft_stage_io(const float2* __restrict__ tw, int tid, LD ld, ST st)
{
	constexpr int NB = M / R;
	constexpr int PER = (NB + NT - 1) / NT;
	constexpr int LOUT = fft_layout_out<M, NS, R>();
	float2 v[PER][R];
#pragma unroll
	for (int b = 0; b < PER; ++b) {
		const int j = tid + b * NT;
		if ((NB % NT == 0) || j < NB) {
			const int k = j & (NS - 1);
			const int jp = lay_pos<LIN>(j);
#pragma unroll
			for (int r = 0; r < R; ++r) {
				if (ZIN && r >= R / 2)
					v[b][r] = make_float2(0.0f, 0.0f);
				else
					v[b][r] = ld(j, r, lay_step<LIN, NB>(j, jp, r));  // element j + r*NB (r is a constant after unrolling)
			}
			if constexpr (NS > 1) {
#pragma unroll
				for (int r = 1; r < R; ++r) {
					if (ZIN && r >= R / 2)
						continue;
					const float2 w = GT ? tw[(r - 1) * NS + k] : __ldg(&tw[(r - 1) * NS + k]);
					v[b][r] = S > 0 ? cmulc(v[b][r], w) : cmul(v[b][r], w);
				}
			}
			dftR<S, R, ZIN>(v[b]);
		}
	}
	if (MID) __syncthreads();
#pragma unroll
	for (int b = 0; b < PER; ++b) {
		const int j = tid + b * NT;
		if ((NB % NT == 0) || j < NB) {
			const int k = j & (NS - 1);
			const int j0 = (j - k) * R + k;
			// position of j0 + r*NS as one base plus compile-time offsets where the geometry allows it
			int base_p;
			if constexpr (NS == 1 && R == 8 && LOUT == LAY_F)
				base_p = 8 * j + (j >> 1);                      // (8j + r) >> 4 == j >> 1
			else if constexpr (NS == 8 && R == 8 && LOUT == LAY_L2)
				base_p = 72 * (j >> 3) + k;                     // 64q + k + 8r lies in block q of 64
			else if constexpr (NS == 8 && R == 8 && LOUT == LAY_F)
				base_p = 68 * (j >> 3) + k;                     // (64q + k + 8r) >> 4 == 4q + (r >> 1)
			else
				base_p = lay_pos<LOUT>(j0);
#pragma unroll
			for (int r = 0; r < R; ++r) {
				if (HOUT && r >= R / 2)
					continue;
				int p;
				if constexpr (NS == 1 && R == 8 && LOUT == LAY_F)
					p = base_p + r;
				else if constexpr (NS == 8 && R == 8 && LOUT == LAY_L2)
					p = base_p + 8 * r;
				else if constexpr (NS == 8 && R == 8 && LOUT == LAY_F)
					p = base_p + 8 * r + (r >> 1);
				else
					p = lay_step<LOUT, NS>(j0, base_p, r);
				st(j0 + r * NS, p, v[b][r]);
			}
		}
	}
}

// M-point complex FFT in shared memory, IN PLACE, all NT threads of the CTA participate.
// tw: the per-stage tables of fft_fill_twiddles(M).  The caller must have synchronised after filling buf.
// Ends synchronised.
template <int M, int NT, int S, int NS = 1, bool ZIN = false, bool HOUT = false, bool GT = false, int LIN = LAY_N>
__device__ __forceinline__ void fft_smem(float2* buf, const float2* __restrict__ tw, int tid)
{
	if constexpr (NS < M) {
		constexpr int R = fft_radix<M, NS>();
		fft_stage_io<M, NT, S, R, NS, LIN, (ZIN && NS == 1), (HOUT && NS * R == M), GT, true>(tw, tid, SmemLoad{buf}, SmemStore{buf});
		__syncthreads();
		fft_smem<M, NT, S, NS * R, ZIN, HOUT, GT, fft_layout_out<M, NS, R>()>(buf, tw + (NS > 1 ? (R - 1) * NS : 0), tid);
	}
}

// ---- out-of-place (ping-pong) variant --------------------------------------
// Each stage reads one buffer and writes the other, so one barrier per stage suffices.  The first stage may take its
// input from a functor `first(j, r, padded position)` (element j + r * M/R) (e.g. the windowed frame straight from global memory) and
// the last stage may hand its output to a functor `last(index, padded position, value)` (e.g. the overlap-add
// into global memory) instead of shared memory.  FIRST_FN / LAST_FN select that; with both false this is the plain
// ping-pong transform from `a` (result in the returned buffer: a after an even number of stages, else b).
// With FIRST_FN the first stage writes b... see fft_pp_first_target.
struct NoFn {
};

template <int M, int NT, int S, int NS, bool ZIN, bool HOUT, bool GT, bool LAST_FN, int LIN, typename LAST>
__device__ __forceinline__ float2* fft_pp_rest(float2* src, float2* dst, const float2* __restrict__ tw, int tid, LAST last)
{
	if constexpr (NS < M) {
		constexpr int R = fft_radix<M, NS>();
		constexpr bool is_last = NS * R == M;
		if constexpr (is_last && LAST_FN) {
			fft_stage_io<M, NT, S, R, NS, LIN, (ZIN && NS == 1), HOUT, GT, false>(tw, tid, SmemLoad{src}, last);
			return nullptr;  // the result went to the functor; NOT synchronised (the caller decides)
		}
		else {
			fft_stage_io<M, NT, S, R, NS, LIN, (ZIN && NS == 1), (HOUT && is_last), GT, false>(tw, tid, SmemLoad{src}, SmemStore{dst});
			__syncthreads();
			return fft_pp_rest<M, NT, S, NS * R, ZIN, HOUT, GT, LAST_FN, fft_layout_out<M, NS, R>()>(dst, src, tw + (NS > 1 ? (R - 1) * NS : 0),
			                                                                                    tid, last);
		}
	}
	else {
		return src;
	}
}

// plain ping-pong transform: input in a, scratch b
template <int M, int NT, int S, int NS = 1, bool ZIN = false, bool HOUT = false, bool GT = false>
__device__ __forceinline__ float2* fft_smem_pp(float2* a, float2* b, const float2* __restrict__ tw, int tid)
{
	static_assert(NS == 1, "fft_smem_pp starts at the first stage");
	return fft_pp_rest<M, NT, S, NS, ZIN, HOUT, GT, false, LAY_N>(a, b, tw, tid, NoFn{});
}

// first stage fed by `first`, written to `a`; the remaining stages ping-pong between a and b.  `last` (LAST_FN)
// receives the output of the final stage.  Requires at least two stages (M >= 64).  Returns the buffer holding the
// result (nullptr with LAST_FN).  Ends synchronised unless LAST_FN.
template <int M, int NT, int S, bool ZIN, bool HOUT, bool GT, bool LAST_FN, typename FIRST, typename LAST>
__device__ __forceinline__ float2* fft_pp_fused(float2* a, float2* b, const float2* __restrict__ tw, int tid, FIRST first, LAST last)
{
	constexpr int R = fft_radix<M, 1>();
	static_assert(R < M, "fft_pp_fused needs at least two stages");
	fft_stage_io<M, NT, S, R, 1, LAY_N, ZIN, false, GT, false>(tw, tid, first, SmemStore{a});
	__syncthreads();
	return fft_pp_rest<M, NT, S, R, ZIN, HOUT, GT, LAST_FN, fft_layout_out<M, 1, R>()>(a, b, tw, tid, last);
}

// in-place transform whose LAST stage hands its output to `last(index, padded position, value)` instead of storing
// it (the other stages need their barrier between loads and stores).  Not synchronised at the end.
template <int M, int NT, int S, int NS, bool HOUT, bool GT, int LIN = LAY_N, typename LAST>
__device__ __forceinline__ void fft_inplace_last(float2* buf, const float2* __restrict__ tw, int tid, LAST last)
{
	if constexpr (NS < M) {
		constexpr int R = fft_radix<M, NS>();
		if constexpr (NS * R == M) {
			fft_stage_io<M, NT, S, R, NS, LIN, false, HOUT, GT, false>(tw, tid, SmemLoad{buf}, last);
		}
		else {
			fft_stage_io<M, NT, S, R, NS, LIN, false, false, GT, true>(tw, tid, SmemLoad{buf}, SmemStore{buf});
			__syncthreads();
			fft_inplace_last<M, NT, S, NS * R, HOUT, GT, fft_layout_out<M, NS, R>()>(buf, tw + (NS > 1 ? (R - 1) * NS : 0), tid, last);
		}
	}
}

}  // namespace zen_b200
