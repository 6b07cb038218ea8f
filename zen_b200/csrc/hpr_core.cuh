// The fused per-hop HPR step: sqrt-Hann window -> real FFT -> |X| ring ->
// time-axis + frequency-axis median (or SSE box means) of the ONE consumed row
// -> hard / soft / SSE masks -> inverse real FFT -> COLA overlap-add.
//
// It restates, for one hop, HPR<GPU>::process_next_hop + apply_median_filter /
// apply_sse_filter (libzen/hps.cu:429-652), but computes only the row the
// reference consumes (row stft_width - lag, hps.cu:501-504) instead of
// filtering the whole stft_width x nfft matrix, and works on the nfft/2+1
// independent bins of the real-input spectrum:
//   * frequency windows index the mirrored half spectrum (|X[nfft-k]| = |X[k]|);
//   * with --nocopybord the reference's masks at bins k and nfft-k differ
//     (forward-looking ROI, mfilt.h:152-158); their effect on Re(ifft) is the
//     half-spectrum mask (M[k] + M[nfft-k]) / 2, which is what we apply.
// tests/test_halfspec_model.py proves this algebra against the oracle on CPU.
#pragma once
#include "fft_smem.cuh"
#include "median_select.cuh"
#include "zen_common.cuh"

namespace zen_b200 {

constexpr int ZEN_MAX_TAPS = 128;

// rounding rule of `x / y >= c` for a float constant c (see thr_ratio_ge)
struct RatioRule {
	double m;      // rounding boundary just below the constant
	int even;      // constant has an even mantissa: a tie rounds up to it
	int trivial;   // 1: every non-negative value passes (constant <= 0); -1: nothing passes (NaN)
};

struct HprDev {
	int hop, W, lag;
	int n_taps;          // time-axis taps (0: the reference never writes the consumed row -> H = 0)
	int Lp, midp, Kp;    // frequency window (odd), its half, registers per lane of the warp-resident sliding window
	int Cp;              // register capacity of the per-thread sliding window (0: use the warp-resident one)
	int copy_bord;       // frequency windows centred+circular (1) or forward-looking (0)
	int out_flags, soft, sse;
	int decide;          // hard mask: decide M = [median >= threshold] by counting taps instead of selecting the median
	float power;         // (float)(int)beta, soft-mask exponent (hps.h:116-129)
	float beta, beta_h;  // beta, beta - eps (hps.cu:505, 540)
	RatioRule rule_p, rule_h;  // exact rounding rules of the two hard-mask tests
	float cola;
	float lh1, lp1;      // l_harm + 1, l_perc + 1 (hps.cu:599-604)
	float inv_lp;        // 1 / Lp for the box mean
	const float* window;   // nwin
	const float2* tw;      // per-stage twiddle tables of the M-point FFT (fft_fill_twiddles), M = nfft/2
	const float2* twr;     // exp(-2 pi i k / nfft), k <= M/2
	short tap_age[ZEN_MAX_TAPS];
};

// per-CTA streaming state (global memory)
struct HprState {
	float* mag_ring;   // W rows of (M+1): |X| (or |X|^2 with SSE), slot = frame index mod W
	float2* x_ring;    // xdepth rows of (M+1); may be null when xdepth == 0
	int xdepth;
	float* tail[3];    // hop floats each (H, P, R): second half of the previous frame * COLA
};

struct HprEmit {
	float* a[3];  // first destination of the emitted hop per output (H, P, R) or null
	float* b[3];  // optional second destination
};

// Table overrides of the resident real-time kernel (copies in shared memory; generic loads).
struct HprTables {
	const float* window;
	const float2* tw;
	const float2* twr;
};

// Tagged emission of the resident real-time kernel: the emitted hop is also written to mapped host memory as
// 16-byte groups {x[3g], x[3g+1], x[3g+2], tag ^ zen_group_hash(x)}.  One group is one aligned 16-byte store and
// validates ITSELF: a reader that catches it half-way (new tag word, old samples, or the reverse - observed on the
// host about once in a million groups when it reads a group the moment it lands) computes a tag that is not the one
// it waits for and simply looks again.  No completion flag behind a system-wide fence is needed (rt_call, hpr_kernels.cu).
__host__ __device__ __forceinline__ unsigned zen_group_hash(unsigned x0, unsigned x1, unsigned x2)
{
	return x0 ^ ((x1 << 11) | (x1 >> 21)) ^ ((x2 << 22) | (x2 >> 10));
}
// What a group's tag word is XORed with.  The four groups of a whole 64-byte line (twelve samples y0..y11, all inside
// the hop: groups below n_line_groups = 4 * (hop / 12)) share one LINE hash, V[l] = y[l] ^ rotl(y[4+l], 11) ^
// rotl(y[8+l], 22), and group q carries tag ^ V[q]: on the host that is three vertical vector operations on the twelve
// samples as they lie in memory (the per-group hash needs a transpose, and the shuffle port is what bounds the host's
// pack / unpack loops).  Every sample of the line is in exactly one V[l], so the line is good when its four tag words
// are.  The few groups behind the last whole line use zen_group_hash.  Must be called by all 32 lanes of the warp; the
// four groups of a line sit in four consecutive lanes.
__device__ __forceinline__ unsigned zen_group_key(unsigned x0, unsigned x1, unsigned x2, int g, int n_line_groups)
{
	auto r11 = [](unsigned v) { return (v << 11) | (v >> 21); };
	auto r22 = [](unsigned v) { return (v << 22) | (v >> 10); };
	const int q = g & 3;
	// this group holds y[3q], y[3q+1], y[3q+2]: their contributions to V[0..3]
	unsigned c0 = 0u, c1 = 0u, c2 = 0u, c3 = 0u;
	if (q == 0) {
		c0 = x0;
		c1 = x1;
		c2 = x2;
	}
	else if (q == 1) {
		c3 = x0;
		c0 = r11(x1);
		c1 = r11(x2);
	}
	else if (q == 2) {
		c2 = r11(x0);
		c3 = r11(x1);
		c0 = r22(x2);
	}
	else {
		c1 = r22(x0);
		c2 = r22(x1);
		c3 = r22(x2);
	}
#pragma unroll
	for (int m = 1; m <= 2; m <<= 1) {
		c0 ^= __shfl_xor_sync(0xffffffffu, c0, m);
		c1 ^= __shfl_xor_sync(0xffffffffu, c1, m);
		c2 ^= __shfl_xor_sync(0xffffffffu, c2, m);
		c3 ^= __shfl_xor_sync(0xffffffffu, c3, m);
	}
	const unsigned vq = q == 0 ? c0 : (q == 1 ? c1 : (q == 2 ? c2 : c3));
	return g < n_line_groups ? vq : zen_group_hash(x0, x1, x2);
}

struct HprPack {
	float* buf;     // hop floats of shared memory
	uint4* dst[3];  // per output (H, P, R) or null
	unsigned tag;
};

template <int NFFT>
struct HprSmem {
	static constexpr int M = NFFT / 2;
	float2* zbuf;  // fpad_size(M)
	float2* xbuf;  // M + 1
	float* erow;   // M + 1 + Lp + 3   (extended magnitude row, later the H row)
	float* prow;   // M + 1
	int* taps;     // ZEN_MAX_TAPS: ring-row offset of every time tap for this hop (-1: frame before the stream start)
	static __host__ __device__ size_t bytes(int Lp)
	{
		size_t z = sizeof(float2) * (size_t)fpad_size(M);
		size_t x = sizeof(float2) * (size_t)(M + 2);
		size_t e = sizeof(float) * (size_t)((M + 1 + Lp + 3 + 12 + 3) & ~3);
		size_t p = sizeof(float) * (size_t)((M + 1 + 3) & ~3);
		return z + x + e + p + sizeof(int) * ZEN_MAX_TAPS;
	}
	__device__ void carve(unsigned char* base, int Lp)
	{
		zbuf = reinterpret_cast<float2*>(base);
		xbuf = zbuf + fpad_size(M);
		erow = reinterpret_cast<float*>(xbuf + (M + 2));
		prow = erow + ((M + 1 + Lp + 3 + 12 + 3) & ~3);
		taps = reinterpret_cast<int*>(prow + ((M + 1 + 3) & ~3));
	}
};

__device__ __forceinline__ float mask_hard(float x, float y, float beta)
{
	return (x / (y + ZEN_EPS)) >= beta ? 1.0f : 0.0f;  // hps.h:100-113
}
__device__ __forceinline__ float mask_soft(float x, float y, float power)
{
	float px = powf(x, power), py = powf(y, power);    // hps.h:116-129
	return px / (px + py + ZEN_EPS);
}
__device__ __forceinline__ float mask_sse(float x, float y)
{
	return x * x / (x * x + y * y + ZEN_EPS);          // hps.h:132-140
}


// ---- real-input FFT glue, shared by every kernel (so that they agree bit for bit) -----------------------
// Bins k and M-k (0 < k < M/2) of the nfft-point real-input spectrum from the M-point transform Z of the frame
// packed as even/odd samples: A = Z[k], Zm = Z[M-k], w = exp(-2 pi i k / nfft).
__device__ __forceinline__ void rfft_split_pair(float2 A, float2 Zm, float2 w, float2& Xa, float2& Xb)
{
	const float2 B = cconj(Zm);
	const float2 Sm = cadd(A, B);
	const float2 O = cmul(w, csub(A, B));
	const float2 Or = make_float2(O.y, -O.x);          // -i O
	Xa = cscale(cadd(Sm, Or), 0.5f);
	Xb = cconj(cscale(csub(Sm, Or), 0.5f));
}
// The inverse: masked spectrum Y[k] = X[k] * ma, Y[M-k] = X[M-k] * mb (hps.h:58-66), w as above -> Z[k], Z[M-k] of the
// packed M-point inverse.  The mask products are written as the multiplicands of explicit fused multiply-adds:
// ptxas (12.9) contracts a packed mul.rn.f32x2 into a following add.rn.f32x2 whenever it sees fit, -fmad=false or
// not, and differently from one kernel to the next - leaving it no mul-then-add to find keeps every kernel that
// inlines this helper bit-identical.  (For the hard masks the products are exact, so nothing changes at all.)
__device__ __forceinline__ void rfft_pack_masked(float2 Xa, float ma, float2 Xb, float mb, float2 w, float2& Zk, float2& Zmk)
{
	const f32x2_t B = pmul(pk(Xb), pk(mb, -mb));                 // conj(X[M-k] * mb); feeds addends only
	const float2 E2 = up(pfma(pk(Xa), pk(ma, ma), B));            // Y[k] + B
	const float2 Dm = up(pfma(pk(Xa), pk(ma, ma), pk(-up(B).x, -up(B).y)));  // Y[k] - B
	const float2 O2 = mul_si<+1>(cmulc(Dm, w));                   // i conj(w) (Y[k] - B)
	Zk = cadd(E2, O2);
	Zmk = cconj(csub(E2, O2));
}
// bins 0 and M (real): Z[0] = (Y0 + YM, Y0 - YM) ; bin M/2: Z[M/2] = 2 conj(Y[M/2]).  Scalar, never contracted.
__device__ __forceinline__ float2 rfft_pack_dc(float2 X0, float m0, float2 XM, float mM)
{
	const float a = __fmul_rn(X0.x, m0), b = __fmul_rn(XM.x, mM);
	return make_float2(__fadd_rn(a, b), __fsub_rn(a, b));
}
__device__ __forceinline__ float2 rfft_pack_mid(float2 Xm, float m)
{
	return make_float2(__fmul_rn(2.0f, __fmul_rn(Xm.x, m)), __fmul_rn(-2.0f, __fmul_rn(Xm.y, m)));
}

// masks of the percussive (mp) and harmonic (mh) outputs at half-spectrum bin k
template <int NFFT>
__device__ __forceinline__ void hpr_masks(const HprDev& P, const float* prow, const float* hrow, int k, float& mp, float& mh)
{
	constexpr int M = NFFT / 2;
	const float H = hrow[k];
	const float Pf = prow[k];
	const bool want_p = (P.out_flags & ZEN_OUTPUT_PERCUSSIVE) != 0;
	const bool want_h = (P.out_flags & ZEN_OUTPUT_HARMONIC) != 0;
	mp = 0.0f;
	mh = 0.0f;
	if (P.sse) {
		if (want_p) mp = mask_sse(Pf, H);
		if (want_h) mh = mask_sse(H, Pf);
		return;
	}
	float Pb = Pf;
	bool two = false;
	if (!P.copy_bord && k != 0 && k != M) {
		two = true;
		Pb = (k > P.Lp) ? prow[k - P.Lp + 1] : 0.0f;
	}
	if (P.soft) {
		if (want_p) mp = two ? 0.5f * (mask_soft(Pf, H, P.power) + mask_soft(Pb, H, P.power)) : mask_soft(Pf, H, P.power);
		if (want_h) mh = two ? 0.5f * (mask_soft(H, Pf, P.power) + mask_soft(H, Pb, P.power)) : mask_soft(H, Pf, P.power);
	}
	else {
		if (want_p) mp = two ? 0.5f * (mask_hard(Pf, H, P.beta) + mask_hard(Pb, H, P.beta)) : mask_hard(Pf, H, P.beta);
		if (want_h) mh = two ? 0.5f * (mask_hard(H, Pf, P.beta_h) + mask_hard(H, Pb, P.beta_h)) : mask_hard(H, Pf, P.beta_h);
	}
}

// |X[k]|: the reference takes thrust::abs of the complex bin (hps.cu:492-493), which is hypotf.  CUDA's hypotf rescales
// its arguments to survive squares that leave the float range: about 25 instructions per bin, 8 % of everything the
// batched kernel executes.  Where x^2 + y^2 stays well inside the normal range - every bin of real audio - the plain
// formula with one fused product is within one ulp of the exact magnitude, like hypotf itself; a bin outside that
// range (or exactly zero) takes hypotf.  Every kernel of the library takes its magnitudes here, so they stay
// bit-identical to each other.
__device__ __forceinline__ float zen_cabs(float x, float y)
{
	const float s = __fmaf_rn(x, x, __fmul_rn(y, y));
	if (s > 1e-30f && s < 1e30f)
		return __fsqrt_rn(s);
	return hypotf(x, y);
}

// ---- hard-mask decisions without computing the median ------------------------
// The hard masks only need to know on which side of a threshold the frequency
// median P[k] lies:  Mp = [P/(H+eps) >= beta],  Mh = [H/(P+eps) >= beta-eps]
// (hps.h:100-113, hps.cu:501-505, 535-540).  Both tests are monotone in P, and
// for a monotone predicate f, f(median(x)) == median(f(x)): the mask is 1 iff at
// least mid+1 of the L window taps pass the test.  We therefore find, once per
// bin, the exact float threshold of the test (the smallest x with
// RN(x/d) >= beta, resp. the largest s with RN(h/s) >= beta_h) and then only
// COUNT taps against it: one compare (ALU pipe) and one add (FMA pipe) per tap
// instead of keeping a sorted window.  The decisions are bit-identical to
// selecting the median and applying the reference's functor.
__device__ __forceinline__ float f_next_up(float x) { return __uint_as_float(__float_as_uint(x) + 1u); }    // x >= 0, finite
__device__ __forceinline__ float f_next_down(float x) { return __uint_as_float(__float_as_uint(x) - 1u); }  // x > 0

// Exact thresholds without trial divisions.  For IEEE round-to-nearest-even division,
//     RN(x / d) >= beta   <=>   x / d > m   or   (x / d == m and beta's mantissa is even)
// where m = (pred(beta) + beta) / 2 is the rounding boundary below beta (25 significant bits, exact in
// double).  x / d >< m is decided exactly as x >< d * m, because the product of a float (24 bits) and m
// (25 bits) is exact in double.  HprDev carries m and the parity for beta and for beta - eps.
// smallest float x >= 0 with RN(x / d) >= beta   (d > 0);  +inf if none
__device__ __forceinline__ float thr_ratio_ge(float d, const RatioRule& r)
{
	if (r.trivial != 0)
		return r.trivial > 0 ? 0.0f : CUDART_INF_F;
	const double t = (double)d * r.m;               // exact
	float x0 = __double2float_ru(t);                 // smallest float >= t
	if ((double)x0 == t && !r.even)                  // exactly on the boundary and the tie rounds down
		x0 = f_next_up(x0);
	return x0;
}

// largest float s > 0 with RN(h / s) >= beta_h   (h >= 0);  -1 if none, +inf if every s passes
__device__ __forceinline__ float thr_ratio_le(float h, const RatioRule& r)
{
	if (r.trivial != 0)
		return r.trivial > 0 ? CUDART_INF_F : -1.0f;
	if (!(h > 0.0f))
		return -1.0f;
	const double hd = (double)h;
	auto pass = [&](float s) -> bool {               // h / s >= boundary, decided on exact products
		const double p = (double)s * r.m;
		return p < hd || (p == hd && r.even);
	};
	float s0 = __double2float_rd(hd / r.m);          // within one ulp of the answer
	if (!(s0 < 3.402823466e+38f))
		s0 = 3.402823466e+38f;
	if (s0 > 0.0f && pass(s0)) {
		const float up = f_next_up(s0);
		return (up <= 3.402823466e+38f && pass(up)) ? up : s0;
	}
	if (!(s0 > 0.0f))
		return (pass(1.401298464e-45f)) ? 1.401298464e-45f : -1.0f;
	const float dn = f_next_down(s0);
	return (dn > 0.0f && pass(dn)) ? dn : -1.0f;
}

constexpr int ZEN_DECIDE_U = 9;  // consecutive bins per thread in the batched kernels: odd, so the lanes' tap loads hit distinct banks

// Decisions for bins k in [k0, k0+U): taps of bin k are E[k + woff .. k + woff + L).
// Returns bit u of *dp = [P >= tau_k], bit u of *dh = [P + eps <= sig_k]  for k = k0 + u.
template <int U, bool WP, bool WH>
__device__ __forceinline__ void decide_group_t(const float* __restrict__ E, const float* __restrict__ hrow, int k0, int kmax, int woff,
                                               int L, const RatioRule& rp, const RatioRule& rh, unsigned& dp, unsigned& dh)
{
	float tau[U], sig[U], cp[U], ch[U];
#pragma unroll
	for (int u = 0; u < U; ++u) {
		const int k = min(k0 + u, kmax);
		const float H = hrow[k];
		tau[u] = WP ? thr_ratio_ge(H + ZEN_EPS, rp) : CUDART_INF_F;
		sig[u] = WH ? thr_ratio_le(H, rh) : -1.0f;
		cp[u] = 0.0f;
		ch[u] = 0.0f;
	}
	const float* base = E + k0 + woff;
	auto tapf = [&](float x, int u) {
		if (WP) cp[u] += (x >= tau[u]) ? 1.0f : 0.0f;
		if (WH) ch[u] += ((x + ZEN_EPS) <= sig[u]) ? 1.0f : 0.0f;
	};
	// head: tap j belongs to bins u <= j only
#pragma unroll
	for (int j = 0; j < U - 1; ++j) {
		const float x = base[j];
#pragma unroll
		for (int u = 0; u <= j; ++u)
			tapf(x, u);
	}
	// body: every bin of the group sees the tap (needs L >= U - 1, guaranteed by the caller)
#pragma unroll 4
	for (int j = U - 1; j < L; ++j) {
		const float x = base[j];
#pragma unroll
		for (int u = 0; u < U; ++u)
			tapf(x, u);
	}
	// tail: tap L + j belongs to bins u > j only
#pragma unroll
	for (int j = 0; j < U - 1; ++j) {
		const float x = base[L + j];
#pragma unroll
		for (int u = j + 1; u < U; ++u)
			tapf(x, u);
	}
	const float need = (float)(L / 2 + 1);
	dp = 0u;
	dh = 0u;
#pragma unroll
	for (int u = 0; u < U; ++u) {
		if (WP) dp |= (cp[u] >= need ? 1u : 0u) << u;
		if (WH) dh |= (ch[u] >= need ? 1u : 0u) << u;
	}
}

template <int U>
__device__ __forceinline__ void decide_group(const float* __restrict__ E, const float* __restrict__ hrow, int k0, int kmax, int woff,
                                             int L, const RatioRule& rp, const RatioRule& rh, bool want_p, bool want_h, unsigned& dp,
                                             unsigned& dh)
{
	dp = 0u;
	dh = 0u;
	if (want_p && want_h)
		decide_group_t<U, true, true>(E, hrow, k0, kmax, woff, L, rp, rh, dp, dh);
	else if (want_p)
		decide_group_t<U, true, false>(E, hrow, k0, kmax, woff, L, rp, rh, dp, dh);
	else if (want_h)
		decide_group_t<U, false, true>(E, hrow, k0, kmax, woff, L, rp, rh, dp, dh);
}

template <int L, typename Get>
__device__ __forceinline__ float median_fixed(Get get)
{
	float v[L];
#pragma unroll
	for (int t = 0; t < L; ++t)
		v[t] = get(t);
	return median_regs<L>(v);
}

// overlap-add of one output and emission of the hop: out = tail + Re(y[0:hop]) * COLA ; tail' = Re(y[hop:nwin]) * COLA
// (hps.h:68-80).  zb holds the inverse transform (packed: zb[n] = (y[2n], y[2n+1])).  Ends synchronised.
template <int NFFT, int NT>
__device__ __forceinline__ void hpr_ola_emit(const HprDev& P, const float2* zb, float* tail, bool fresh_tail, float* ea, float* eb,
                                             float* pbuf, uint4* pdst, unsigned ptag)
{
	constexpr int M = NFFT / 2, HOP = M / 2;
	const int tid = threadIdx.x;
	for (int n = tid; n < HOP / 2; n += NT) {
		float2 v = zb[n];
		float2 t = fresh_tail ? make_float2(0.0f, 0.0f) : reinterpret_cast<const float2*>(tail)[n];
		float2 r = make_float2(fmaf(v.x, P.cola, t.x), fmaf(v.y, P.cola, t.y));
		if (ea) __stcs(reinterpret_cast<float2*>(ea) + n, r);  // written once, never re-read by the kernel
		if (eb) __stcs(reinterpret_cast<float2*>(eb) + n, r);
		if (pdst) reinterpret_cast<float2*>(pbuf)[n] = r;
	}
	__syncthreads();
	if (pdst) {
		constexpr int NG = (HOP + 2) / 3, NLG = 4 * (HOP / 12);
		for (int g0 = 0; g0 < NG; g0 += NT) {  // (every thread makes every trip: zen_group_key shuffles)
			const int g = g0 + tid;
			uint4 v = make_uint4(0u, 0u, 0u, 0u);
			if (g < NG) {
				v.x = __float_as_uint(pbuf[3 * g]);
				v.y = 3 * g + 1 < HOP ? __float_as_uint(pbuf[3 * g + 1]) : 0u;
				v.z = 3 * g + 2 < HOP ? __float_as_uint(pbuf[3 * g + 2]) : 0u;
			}
			v.w = ptag ^ zen_group_key(v.x, v.y, v.z, g, NLG);
			if (g < NG)
				asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(pdst + g), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
				             : "memory");
		}
	}
	for (int n = HOP / 2 + tid; n < HOP; n += NT) {
		float2 v = zb[n];
		reinterpret_cast<float2*>(tail)[n - HOP / 2] = make_float2(v.x * P.cola, v.y * P.cola);
	}
	__syncthreads();
}

// One hop.  `full` == false: analysis only (fills the rings; halo iterations of a tile).
// All NT threads of the CTA must call it with identical arguments.
// cur_stash (optional): the incoming hop is also copied there while it is read (the persistent
// real-time kernel keeps it in shared memory as the next hop's `prev`).
template <int NFFT, int NT, int U = ZEN_DECIDE_U, bool GT = false>
__device__ __forceinline__ void hpr_iteration(const HprDev& P, HprSmem<NFFT>& sm, const HprState& st, const int i,
                                              const float* __restrict__ prev, const float* __restrict__ cur,
                                              bool full, bool fresh_tail, const HprEmit& em, float* cur_stash = nullptr,
                                              unsigned long long* stamps = nullptr, const float* next_hop = nullptr,
                                              const HprTables* tb = nullptr, const HprPack* pk = nullptr)
{
	const float* const t_window = GT ? tb->window : P.window;
	const float2* const t_tw = GT ? tb->tw : P.tw;
	const float2* const t_twr = GT ? tb->twr : P.twr;
	auto ldt = [](const auto* p) { return GT ? *p : __ldg(p); };
	// diagnostics (tools/rt_phases.py): SM cycle counter at the phase boundaries.  Not %globaltimer: reading it takes
	// a fraction of a microsecond, and thread 0 would hold every barrier of the hop back by that much.
	auto stamp = [&](int idx) {
		if (stamps && threadIdx.x == 0)
			stamps[idx] = (unsigned long long)clock64();
	};
	stamp(0);
	constexpr int M = NFFT / 2;   // complex FFT length; also nwin
	constexpr int HOP = M / 2;
	const int tid = threadIdx.x;
	const int W = P.W;
	const int eoff = (P.copy_bord || P.sse) ? P.midp : 0;

	// ring rows of the time-axis taps for this hop (read in step F, after several barriers)
	for (int t = tid; t < P.n_taps; t += NT) {
		int j = i - P.tap_age[t];
		sm.taps[t] = j >= 0 ? (j % W) * (M + 1) : -1;
	}

	// pull the NEXT hop towards L2 while this one is processed (the batched kernels read the input from HBM
	// exactly once; without this the first loads of every hop wait a full DRAM round trip)
	if (next_hop != nullptr && tid < (HOP * 4) / 128)
		asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(next_hop) + tid * 128));

	// ---- A. window the 2-hop frame, pack even/odd samples as one complex value (hps.cu:452-462)
	// (the zero padding n >= HOP is never stored: the first FFT stage knows it is zero)
	for (int n = tid; n < HOP; n += NT) {
		float2 x;
		if (n < HOP / 2)  // last use of that hop: streaming load, so it does not push the per-CTA scratch out of L2
			x = !prev ? make_float2(0.0f, 0.0f)
			          : (__isGlobal(prev) ? __ldcs(reinterpret_cast<const float2*>(prev) + n)  // (the resident kernel keeps it in smem)
			                              : reinterpret_cast<const float2*>(prev)[n]);
		else {
			x = reinterpret_cast<const float2*>(cur)[n - HOP / 2];
			if (cur_stash) reinterpret_cast<float2*>(cur_stash)[n - HOP / 2] = x;
		}
		float2 w = ldt(reinterpret_cast<const float2*>(t_window) + n);
		sm.zbuf[n] = make_float2(__fmul_rn(x.x, w.x), __fmul_rn(x.y, w.y));
	}
	__syncthreads();
	stamp(1);

	// ---- B. forward FFT (hps.cu:465)
	fft_smem<M, NT, -1, 1, true, false, GT>(sm.zbuf, t_tw, tid);
	stamp(2);

	// ---- C. split into the real-input spectrum X[0..M], magnitudes into the ring (hps.cu:469-472, 492-493)
	{
		float* mag_row = st.mag_ring + (size_t)(i % W) * (M + 1);
		float2* xg = st.xdepth > 0 ? st.x_ring + (size_t)(i % st.xdepth) * (M + 1) : nullptr;
		const bool direct = (P.lag == 1);
		for (int k = tid; k <= M / 2; k += NT) {
			float2 Xa, Xb;
			int ka = k, kb = M - k;
			if (k == 0) {
				float2 Z0 = sm.zbuf[0];
				Xa = make_float2(Z0.x + Z0.y, 0.0f);
				Xb = make_float2(Z0.x - Z0.y, 0.0f);
			}
			else if (k == M / 2) {
				Xa = cconj(sm.zbuf[M / 2]);
				Xb = Xa;
			}
			else {
				rfft_split_pair(sm.zbuf[k], sm.zbuf[M - k], ldt(&t_twr[k]), Xa, Xb);
			}
			float ma = zen_cabs(Xa.x, Xa.y), mb = zen_cabs(Xb.x, Xb.y);
			if (P.sse) {
				ma = powf(ma, 2.0f);  // hps.h:91-98
				mb = powf(mb, 2.0f);
			}
			mag_row[ka] = ma;
			if (kb != ka) mag_row[kb] = mb;
			if (xg) {
				xg[ka] = Xa;
				if (kb != ka) xg[kb] = Xb;
			}
			if (direct) {
				sm.xbuf[ka] = Xa;
				sm.erow[eoff + ka] = P.sse ? 1.0f / ma : ma;
				if (kb != ka) {
					sm.xbuf[kb] = Xb;
					sm.erow[eoff + kb] = P.sse ? 1.0f / mb : mb;
				}
			}
		}
	}
	__syncthreads();
	stamp(3);
	if (!full)
		return;

	// ---- D. the consumed frame: row stft_width - lag (hps.cu:501-504, 517-519)
	const int jc = i - P.lag + 1;
	if (P.lag > 1) {
		const float* mrow = st.mag_ring + (size_t)((jc >= 0 ? jc : 0) % W) * (M + 1);
		const float2* xrow = st.x_ring + (size_t)((jc >= 0 ? jc : 0) % st.xdepth) * (M + 1);
		for (int k = tid; k <= M; k += NT) {
			float m = jc >= 0 ? mrow[k] : 0.0f;
			sm.erow[eoff + k] = P.sse ? 1.0f / m : m;
			sm.xbuf[k] = jc >= 0 ? xrow[k] : make_float2(0.0f, 0.0f);
		}
		__syncthreads();
	}
	// mirrored borders: |X[-t]| = |X[t]|, |X[M+t]| = |X[M-t]|
	if (eoff > 0) {
		for (int t = tid; t < P.midp; t += NT) {
			sm.erow[eoff - 1 - t] = sm.erow[eoff + 1 + t];
			sm.erow[eoff + M + 1 + t] = sm.erow[eoff + M - 1 - t];
		}
	}
	else {
		for (int t = tid; t < P.Lp - 1; t += NT)
			sm.erow[M + 1 + t] = sm.erow[M - 1 - t];
	}
	__syncthreads();

	// time axis: H row (hps.cu:495 / 595, only the consumed row).  For the short windows the bins-per-thread
	// loop is fully unrolled so that all ring loads of a thread are in flight together.
	auto compute_h_row = [&](float* dst) {
		const int nt = P.n_taps;
		constexpr int ITER = (M + 1 + NT - 1) / NT;
		auto tap = [&](int t, int k) -> float {
			int off = sm.taps[t];
			return off >= 0 ? st.mag_ring[off + k] : 0.0f;
		};
		auto fixed = [&](auto LC) {
			constexpr int L = decltype(LC)::value;
			int offs[L];
#pragma unroll
			for (int t = 0; t < L; ++t)
				offs[t] = sm.taps[t];
#pragma unroll
			for (int it = 0; it < ITER; ++it) {
				const int k = tid + it * NT;
				if (k <= M) {
					float v[L];
#pragma unroll
					for (int t = 0; t < L; ++t)
						v[t] = offs[t] >= 0 ? st.mag_ring[offs[t] + k] : 0.0f;
					dst[k] = median_regs<L>(v);
				}
			}
		};
		if (!P.sse && (nt == 1 || nt == 3 || nt == 5 || nt == 7)) {
			switch (nt) {
			case 1: fixed(std::integral_constant<int, 1>{}); break;
			case 3: fixed(std::integral_constant<int, 3>{}); break;
			case 5: fixed(std::integral_constant<int, 5>{}); break;
			default: fixed(std::integral_constant<int, 7>{}); break;
			}
			return;
		}
		if (!P.sse && nt > 13 && nt <= 127) {
			// long time windows (hops below 256): one warp per bin sorts the taps across its lanes
			// (bitonic network, K ranks per lane) and reads the middle rank
			constexpr int NW = NT / 32;
			const int wid = tid >> 5, lane = tid & 31;
			for (int k = wid; k <= M; k += NW) {
				auto load = [&](int t) { return tap(t, k); };
				float med;
				if (nt <= 31) {
					WarpSortedWindow<1> w;
					w.init(load, nt, lane);
					med = __shfl_sync(0xffffffffu, w.median_reg(), w.med_lane);
				}
				else if (nt <= 63) {
					WarpSortedWindow<2> w;
					w.init(load, nt, lane);
					med = __shfl_sync(0xffffffffu, w.median_reg(), w.med_lane);
				}
				else {
					WarpSortedWindow<4> w;
					w.init(load, nt, lane);
					med = __shfl_sync(0xffffffffu, w.median_reg(), w.med_lane);
				}
				if (lane == 0)
					dst[k] = med;
			}
			return;
		}
		for (int k = tid; k <= M; k += NT) {
			float H;
			if (P.sse) {
				float acc = 0.0f;
				for (int t = 0; t < nt; ++t)
					acc += 1.0f / tap(t, k);
				float mean = acc / (float)nt;
				H = (1.0f / mean) * P.lh1;  // hps.cu:602-604
			}
			else {
				switch (nt) {
				case 0: H = 0.0f; break;
				case 9: H = median_fixed<9>([&](int t) { return tap(t, k); }); break;
				case 11: H = median_fixed<11>([&](int t) { return tap(t, k); }); break;
				case 13: H = median_fixed<13>([&](int t) { return tap(t, k); }); break;
				default: H = median_generic([&](int t) { return tap(t, k); }, nt); break;
				}
			}
			dst[k] = H;
		}
	};

	const bool decide = P.decide && !P.sse && !P.soft && P.Lp >= U - 1;
	if (decide) {
		// ---- E'/F'. hard mask by counting (see decide_group).  The H row goes into zbuf, which is
		// free between the split pass and the inverse-FFT build; codes go into prow:
		// bit0/bit1 = percussive decision at bins k / nfft-k, bit2/bit3 = harmonic.
		float* hrow = reinterpret_cast<float*>(sm.zbuf);
		compute_h_row(hrow);
		__syncthreads();
		stamp(4);
		unsigned* codes = reinterpret_cast<unsigned*>(sm.prow);
		const bool want_p = (P.out_flags & ZEN_OUTPUT_PERCUSSIVE) != 0;
		const bool want_h = (P.out_flags & ZEN_OUTPUT_HARMONIC) != 0;
		constexpr int NG = (M + 1 + U - 1) / U;
		for (int g = tid; g < NG; g += NT) {
			const int k0 = g * U;
			unsigned fp, fh;
			decide_group<U>(sm.erow, hrow, k0, M, 0, P.Lp, P.rule_p, P.rule_h, want_p, want_h, fp, fh);
			unsigned bp = fp, bh = fh;
			if (!P.copy_bord) {
				// value at bin nfft-k: window k-L+1 .. k for k > L, never written (P = 0) for 1 <= k <= L
				bp = 0u;
				bh = 0u;
				if (k0 + U - 1 > P.Lp)
					decide_group<U>(sm.erow, hrow, k0, M, -(P.Lp - 1), P.Lp, P.rule_p, P.rule_h, want_p, want_h, bp, bh);
			}
#pragma unroll
			for (int u = 0; u < U; ++u) {
				const int k = k0 + u;
				if (k <= M) {
					unsigned p0 = (fp >> u) & 1u, h0 = (fh >> u) & 1u, p1 = (bp >> u) & 1u, h1 = (bh >> u) & 1u;
					if (!P.copy_bord) {
						if (k == 0 || k == M) {
							p1 = p0;
							h1 = h0;
						}
						else if (k <= P.Lp) {
							const float H = hrow[k];
							p1 = (want_p && (0.0f / (H + ZEN_EPS)) >= P.beta) ? 1u : 0u;
							h1 = (want_h && (H / (0.0f + ZEN_EPS)) >= P.beta_h) ? 1u : 0u;
						}
					}
					codes[k] = p0 | (p1 << 1) | (h0 << 2) | (h1 << 3);
				}
			}
		}
		__syncthreads();
		stamp(5);
	}
	else {
		// ---- E. frequency axis: prow[s] = median / mean of erow[s .. s+Lp)
		if (!P.sse) {
			if (P.Cp > 0) {
				const int R = (M + 1 + NT - 1) / NT;
				const int s0 = tid * R;
				const int s1 = min(M + 1, s0 + R);
				thread_sliding_median_dyn(P.Cp, sm.erow, sm.prow, s0, s1, P.Lp);
			}
			else {
				constexpr int NW = NT / 32;
				const int wid = tid >> 5, lane = tid & 31;
				const int R = (M + 1 + NW - 1) / NW;
				const int s0 = wid * R;
				const int s1 = min(M + 1, s0 + R);
				warp_sliding_median_dyn<float>(P.Kp, sm.erow, sm.prow, s0, s1, P.Lp, lane);
			}
		}
		else {
			for (int k = tid; k <= M; k += NT) {
				float acc = 0.0f;
				for (int t = 0; t < P.Lp; ++t)
					acc += sm.erow[k + t];
				float mean = acc * P.inv_lp;
				sm.prow[k] = (1.0f / mean) * P.lp1;  // hps.cu:599-601
			}
		}
		__syncthreads();
		// ---- F. time axis: H row (into erow, whose magnitudes are no longer needed)
		compute_h_row(sm.erow);
		__syncthreads();
	}

	// ---- G. per output: mask, inverse real FFT, overlap-add (hps.cu:498-579, 607-651)
	// order P, H, R as in the reference; output index 0 = H, 1 = P, 2 = R
	// (no dynamically indexed local arrays: they would live in local memory, and a kernel launched per hop or a
	// resident one after its system-wide fence finds none of it in L1)
#pragma unroll 1
	for (int oi = 0; oi < 3; ++oi) {
		const int o = oi == 0 ? 1 : (oi == 1 ? 0 : 2);
		if (!(P.out_flags & (1 << o)))
			continue;
		if (o == 2 && (P.soft || P.sse))
			continue;  // residual only exists for the hard mask (hps.cu:562)
		for (int k = tid; k <= M / 2; k += NT) {
			const int kb = M - k;
			float mpa, mha, mpb, mhb;
			if (decide) {
				const unsigned ca = reinterpret_cast<const unsigned*>(sm.prow)[k];
				const unsigned cb = reinterpret_cast<const unsigned*>(sm.prow)[kb];
				mpa = 0.5f * (float)((ca & 1u) + ((ca >> 1) & 1u));
				mha = 0.5f * (float)(((ca >> 2) & 1u) + ((ca >> 3) & 1u));
				mpb = 0.5f * (float)((cb & 1u) + ((cb >> 1) & 1u));
				mhb = 0.5f * (float)(((cb >> 2) & 1u) + ((cb >> 3) & 1u));
			}
			else {
				hpr_masks<NFFT>(P, sm.prow, sm.erow, k, mpa, mha);
				hpr_masks<NFFT>(P, sm.prow, sm.erow, kb, mpb, mhb);
			}
			float ma = (o == 1) ? mpa : (o == 0 ? mha : 1.0f - (mha + mpa));  // hps.h:35-43
			float mb = (o == 1) ? mpb : (o == 0 ? mhb : 1.0f - (mhb + mpb));
			float2 Xa = sm.xbuf[k], Xb = sm.xbuf[kb];
			if (k == 0) {
				sm.zbuf[0] = rfft_pack_dc(Xa, ma, Xb, mb);
			}
			else if (k == M / 2) {
				sm.zbuf[M / 2] = rfft_pack_mid(Xa, ma);
			}
			else {
				float2 Zk, Zmk;
				rfft_pack_masked(Xa, ma, Xb, mb, ldt(&t_twr[k]), Zk, Zmk);
				sm.zbuf[k] = Zk;
				sm.zbuf[kb] = Zmk;
			}
		}
		__syncthreads();
		stamp(6);
		fft_smem<M, NT, +1, 1, false, true, GT>(sm.zbuf, t_tw, tid);
		stamp(7);
		float* const tail = o == 0 ? st.tail[0] : (o == 1 ? st.tail[1] : st.tail[2]);
		float* const ea = o == 0 ? em.a[0] : (o == 1 ? em.a[1] : em.a[2]);
		float* const eb = o == 0 ? em.b[0] : (o == 1 ? em.b[1] : em.b[2]);
		uint4* const pdst = !pk ? nullptr : (o == 0 ? pk->dst[0] : (o == 1 ? pk->dst[1] : pk->dst[2]));
		hpr_ola_emit<NFFT, NT>(P, sm.zbuf, tail, fresh_tail, ea, eb, pk ? pk->buf : nullptr, pdst, pk ? pk->tag : 0u);
		stamp(8);
	}
}

// ---- the batched fast path -------------------------------------------------------------------------------
// hpr_fast_iteration serves the default plan of HPRRealtime / the headline batch (hard mask decided by counting,
// copy-border, lag 1, at most 7 time taps) with about half the instructions of hpr_iteration.  Same arithmetic,
// same helper functions, bit-identical results; what changes is how the data moves:
//   * the window multiply is the loader of the first forward-FFT stage (straight from global memory) and the
//     overlap-add is the sink of the last inverse-FFT stage (straight to global memory): no staging passes;
//   * the FFTs ping-pong between two shared-memory buffers (one barrier per stage); the spectrum X of the frame
//     and the packed masked spectrum live in the same two buffers;
//   * one thread decides EIGHT consecutive bins: their time medians come from 16-byte loads of the |X| ring, the
//     8 + L - 1 frequency taps they share from 16-byte shared-memory loads, and the per-tap work is two compares
//     (FSET) and ONE packed add (FADD2) per pair of bins - 1.5 issue slots per (tap, bin) instead of 2;
//   * decisions travel as one bit per bin.
template <int NFFT>
struct FastSmem {
	static constexpr int M = NFFT / 2;
	static constexpr int ZN = (fpad_size(M) + 1) & ~1;
	float2* za;            // ZN
	float2* zb;            // ZN
	float* erow;           // |X| of the consumed frame with mirrored borders: bin k is element midp + k, stored at
	                       // fast_esw(midp + k) (16-byte chunks swizzled against bank conflicts); 16-byte aligned.
	                       // NOT storage of its own: between the split and the masked pack one of the two FFT buffers is
	                       // free (the split works in place), and erow lives there (set per hop by hpr_fast_iteration)
	unsigned short* codes; // entry k >> 3, bit k & 7: percussive decision of bin k; bit 8 + (k & 7): harmonic
	// (a multiple of 32 floats: the swizzle permutes chunks within blocks of eight)
	static __host__ __device__ size_t erow_floats(int Lp) { return (size_t)((M + 1 + Lp + 16 + 31) & ~31); }
	static __host__ __device__ size_t bytes(int Lp)
	{
		return 2 * sizeof(float2) * (size_t)ZN + sizeof(unsigned short) * (size_t)((M / 8 + 1 + 7) & ~7);
	}
	__device__ void carve(unsigned char* base, int Lp)
	{
		(void)Lp;
		za = reinterpret_cast<float2*>(base);
		zb = za + ZN;
		erow = nullptr;
		codes = reinterpret_cast<unsigned short*>(zb + ZN);
	}
};

// Element e of the magnitude row lives at fast_esw(e): the 16-byte chunk index c = e >> 2 has its lowest bit flipped in
// every other block of eight chunks.  The decision threads read one chunk each at a stride of TWO chunks (eight bins
// per thread): the lanes t and t + 4 of a quarter-warp then hit the same four banks in the plain layout - with the
// swizzle they are in neighbouring blocks, one of which is flipped.
__device__ __forceinline__ int fast_esw_chunk(int c) { return c ^ ((c >> 3) & 1); }
__device__ __forceinline__ int fast_esw(int e) { return e ^ ((e >> 3) & 4); }  // == (fast_esw_chunk(e >> 2) << 2) | (e & 3)

// per-CTA state of the fast path (global scratch that stays in L2)
struct FastState {
	float* mag_ring;   // W rows of ring_stride floats (multiple of 4, rows 16-byte aligned)
	int ring_stride;
	float* tail[3];    // HOP floats per output (H, P, R)
};

__host__ __device__ inline bool hpr_fast_supported(const HprDev& P)
{
	// (the magnitude row borrows one FFT buffer of fpad_size(M) float2: it must fit, which it does from nfft 512 on
	// whatever the window, and for the shorter transforms unless the window is nearly as long as the row)
	const int M = 2 * P.hop;
	const long buf_bytes = 8L * ((M + M / 8 + 2) & ~1), erow_bytes = 4L * ((M + 1 + P.Lp + 16 + 31) & ~31);
	return P.decide && !P.sse && !P.soft && P.copy_bord && P.lag == 1 && P.Lp >= 9 && (P.n_taps == 1 || P.n_taps == 3 || P.n_taps == 5 || P.n_taps == 7)
	       && erow_bytes <= buf_bytes;
}

// peak of |emitted sample| per output, kept by each thread across the hops of a tile (optional)
struct FastPeaks {
	float v[3];
};

// ring offset of time tap t for hop i (ring slot `slot` = i mod W), -1: the frame lies before the stream start
__device__ __forceinline__ int fast_ring_off(const HprDev& P, const FastState& st, int i, int slot, int t)
{
	const int age = P.tap_age[t];
	int sl = slot - age;
	sl += sl < 0 ? P.W : 0;
	return (i - age) >= 0 ? sl * st.ring_stride : -1;
}

// time medians (hps.cu:495, only the consumed row) of the eight bins k0 .. k0 + 7
template <int NTAPS>
__device__ __forceinline__ void fast_h8(const HprDev& P, const FastState& st, int i, int slot, int k0, float (&H)[8])
{
#pragma unroll
	for (int half = 0; half < 2; ++half) {
		float4 v[NTAPS];
#pragma unroll
		for (int t = 0; t < NTAPS; ++t) {
			const int off = fast_ring_off(P, st, i, slot, t);
			v[t] = off >= 0 ? *reinterpret_cast<const float4*>(st.mag_ring + off + k0 + 4 * half) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
		}
		float w[NTAPS];
#pragma unroll
		for (int t = 0; t < NTAPS; ++t) w[t] = v[t].x;
		H[4 * half + 0] = median_regs<NTAPS>(w);
#pragma unroll
		for (int t = 0; t < NTAPS; ++t) w[t] = v[t].y;
		H[4 * half + 1] = median_regs<NTAPS>(w);
#pragma unroll
		for (int t = 0; t < NTAPS; ++t) w[t] = v[t].z;
		H[4 * half + 2] = median_regs<NTAPS>(w);
#pragma unroll
		for (int t = 0; t < NTAPS; ++t) w[t] = v[t].w;
		H[4 * half + 3] = median_regs<NTAPS>(w);
	}
}
__device__ __forceinline__ float fast_h1(const HprDev& P, const FastState& st, int i, int slot, int k)
{
	float w[7];
#pragma unroll
	for (int t = 0; t < 7; ++t) {
		w[t] = 0.0f;
		if (t < P.n_taps) {
			const int off = fast_ring_off(P, st, i, slot, t);
			w[t] = off >= 0 ? st.mag_ring[off + k] : 0.0f;
		}
	}
	if (P.n_taps == 1) return w[0];
	if (P.n_taps == 3) {
		float u[3] = {w[0], w[1], w[2]};
		return median_regs<3>(u);
	}
	if (P.n_taps == 5) {
		float u[5] = {w[0], w[1], w[2], w[3], w[4]};
		return median_regs<5>(u);
	}
	return median_regs<7>(w);
}

// one frequency tap x that belongs to bins LO..HI of the U a thread decides: counts, two bins per packed register
template <int LO, int HI, int U, bool WP, bool WH>
__device__ __forceinline__ void fast_tap(float x, const float (&tau)[U], const float (&sig)[U], f32x2_t (&cp)[U / 2], f32x2_t (&ch)[U / 2])
{
	const float xe = x + ZEN_EPS;
#pragma unroll
	for (int q = 0; q < U / 2; ++q) {
		const int u0 = 2 * q, u1 = 2 * q + 1;
		const bool a0 = u0 >= LO && u0 <= HI, a1 = u1 >= LO && u1 <= HI;
		if (!a0 && !a1) continue;
		if (WP) {
			const float s0 = a0 ? ((x >= tau[u0]) ? 1.0f : 0.0f) : 0.0f;
			const float s1 = a1 ? ((x >= tau[u1]) ? 1.0f : 0.0f) : 0.0f;
			cp[q] = padd(cp[q], pk(s0, s1));
		}
		if (WH) {
			const float s0 = a0 ? ((xe <= sig[u0]) ? 1.0f : 0.0f) : 0.0f;
			const float s1 = a1 ? ((xe <= sig[u1]) ? 1.0f : 0.0f) : 0.0f;
			ch[q] = padd(ch[q], pk(s0, s1));
		}
	}
}
// taps i = 0 .. N-1 of a run: tap i belongs to bins max(0, i - FULL + 1) + ... i.e. the first FULL taps to all U bins,
// tap FULL + t to bins t + 1 .. U - 1 (the tail of the window sweep); tv holds the run
template <int I, int N, int FULL, int U, bool WP, bool WH>
__device__ __forceinline__ void fast_tail(const float* tv, const float (&tau)[U], const float (&sig)[U], f32x2_t (&cp)[U / 2], f32x2_t (&ch)[U / 2])
{
	if constexpr (I < N) {
		constexpr int LO = I < FULL ? 0 : I - FULL + 1;
		fast_tap<LO, U - 1, U, WP, WH>(tv[I], tau, sig, cp, ch);
		fast_tail<I + 1, N, FULL, U, WP, WH>(tv, tau, sig, cp, ch);
	}
}
template <int I, int U, bool WP, bool WH>
__device__ __forceinline__ void fast_head(const float* tv, const float (&tau)[U], const float (&sig)[U], f32x2_t (&cp)[U / 2], f32x2_t (&ch)[U / 2])
{
	if constexpr (I < U) {
		fast_tap<0, I, U, WP, WH>(tv[I], tau, sig, cp, ch);
		fast_head<I + 1, U, WP, WH>(tv, tau, sig, cp, ch);
	}
}

// Decisions of U consecutive bins (U = 8 in the batched kernel, 4 where latency counts) given their time medians:
// bit u = percussive, bit 8 + u = harmonic (Mp = [P/(H+eps) >= beta], Mh = [H/(P+eps) >= beta-eps], hps.h:100-113,
// decided by counting: see decide_group_t).  ldv(v) returns taps 4 v .. 4 v + 3 of the group: tap j belongs to bin u for
// u <= j < u + L (L = P.Lp odd, >= U + 1).
template <int U, bool WP, bool WH, typename LDV>
__device__ __forceinline__ unsigned fast_decide(const HprDev& P, LDV ldv, const float (&H)[U])
{
	static_assert(U == 4 || U == 8, "four or eight bins per thread");
	const int L = P.Lp;
	float tau[U], sig[U];
#pragma unroll
	for (int u = 0; u < U; ++u) {
		tau[u] = WP ? thr_ratio_ge(H[u] + ZEN_EPS, P.rule_p) : CUDART_INF_F;
		sig[u] = WH ? thr_ratio_le(H[u], P.rule_h) : -1.0f;
	}
	f32x2_t cp[U / 2], ch[U / 2];
#pragma unroll
	for (int q = 0; q < U / 2; ++q)
		cp[q] = ch[q] = pk(0.0f, 0.0f);
	// head: taps 0 .. U-1, tap j belongs to bins 0 .. j (L > U, so tap U-1 already belongs to all of them)
	{
		float tv[U];
#pragma unroll
		for (int v = 0; v < U / 4; ++v) {
			const float4 x = ldv(v);
			tv[4 * v] = x.x;
			tv[4 * v + 1] = x.y;
			tv[4 * v + 2] = x.z;
			tv[4 * v + 3] = x.w;
		}
		fast_head<0, U, WP, WH>(tv, tau, sig, cp, ch);
	}
	// body: whole vectors of taps that belong to every bin: U <= 4 v and 4 v + 3 <= L - 1
	const int v_end = (L - 4) / 4;
	int v = U / 4;
#pragma unroll 2
	for (; v <= v_end; ++v) {
		const float4 x = ldv(v);
		fast_tap<0, U - 1, U, WP, WH>(x.x, tau, sig, cp, ch);
		fast_tap<0, U - 1, U, WP, WH>(x.y, tau, sig, cp, ch);
		fast_tap<0, U - 1, U, WP, WH>(x.z, tau, sig, cp, ch);
		fast_tap<0, U - 1, U, WP, WH>(x.w, tau, sig, cp, ch);
	}
	// tail: taps 4 v .. L + U - 2, tap j belongs to bins j - L + 1 .. U - 1.  L is odd: either L = 4 v + 3 (three more
	// taps that belong to every bin, then L .. L + U - 2) or L = 4 v + 1 (one more, then the same); both fully unrolled.
	{
		constexpr int NV = (3 + U - 1 + 3) / 4;  // vectors that hold the longer of the two runs
		float tv[4 * NV];
#pragma unroll
		for (int w = 0; w < NV; ++w) {
			const float4 x = ldv(v + w);
			tv[4 * w] = x.x;
			tv[4 * w + 1] = x.y;
			tv[4 * w + 2] = x.z;
			tv[4 * w + 3] = x.w;
		}
		if ((L & 3) == 3)
			fast_tail<0, 3 + U - 1, 3, U, WP, WH>(tv, tau, sig, cp, ch);
		else
			fast_tail<0, 1 + U - 1, 1, U, WP, WH>(tv, tau, sig, cp, ch);
	}
	const float need = (float)(L / 2 + 1);
	unsigned code = 0u;
#pragma unroll
	for (int q = 0; q < U / 2; ++q) {
		const float2 c = up(cp[q]), d = up(ch[q]);
		if (WP) {
			code |= (c.x >= need ? 1u : 0u) << (2 * q);
			code |= (c.y >= need ? 1u : 0u) << (2 * q + 1);
		}
		if (WH) {
			code |= (d.x >= need ? 1u : 0u) << (8 + 2 * q);
			code |= (d.y >= need ? 1u : 0u) << (8 + 2 * q + 1);
		}
	}
	return code;
}
// dispatch on the enabled masks (a mask is only computed for an output that is enabled, hps.cu:498-567)
template <int U, typename LDV>
__device__ __forceinline__ unsigned fast_decide_flags(const HprDev& P, LDV ldv, const float (&H)[U])
{
	const bool want_p = (P.out_flags & ZEN_OUTPUT_PERCUSSIVE) != 0;
	const bool want_h = (P.out_flags & ZEN_OUTPUT_HARMONIC) != 0;
	if (want_p && want_h) return fast_decide<U, true, true>(P, ldv, H);
	if (want_p) return fast_decide<U, true, false>(P, ldv, H);
	if (want_h) return fast_decide<U, false, true>(P, ldv, H);
	return 0u;
}

// masked spectrum of output O (0 harmonic, 1 percussive, 2 residual) packed for the M-point inverse transform:
// X (natural order) * mask -> Z (natural order); hps.h:35-43, 58-66
template <int NFFT, int NT, int O>
__device__ __forceinline__ void fast_pack(const HprDev& P, const unsigned short* __restrict__ codes, const float2* __restrict__ X, float2* __restrict__ Z)
{
	constexpr int M = NFFT / 2;
	const int tid = threadIdx.x;
	auto mask = [&](int k) -> float {
		const unsigned c = (unsigned)codes[k >> 3] >> (k & 7);
		if (O == 1) return (float)(c & 1u);
		if (O == 0) return (float)((c >> 8) & 1u);
		return 1.0f - ((float)((c >> 8) & 1u) + (float)(c & 1u));
	};
	constexpr int TRIPS = (M / 2 + NT - 1) / NT;
#pragma unroll
	for (int it = 0; it < TRIPS; ++it) {
		const int k = tid + it * NT;
		if (((M / 2) % NT == 0 || k < M / 2) && (it > 0 || k > 0)) {
			float2 Zk, Zmk;
			rfft_pack_masked(X[k], mask(k), X[M - k], mask(M - k), __ldg(&P.twr[k]), Zk, Zmk);
			Z[k] = Zk;
			Z[M - k] = Zmk;
		}
	}
	if (tid == 0)
		Z[0] = rfft_pack_dc(X[0], mask(0), X[M], mask(M));
	else if (tid == 32 % NT)
		Z[M / 2] = rfft_pack_mid(X[M / 2], mask(M / 2));
}

template <int NFFT, int NT, bool PEAKS>
__device__ __forceinline__ void hpr_fast_iteration(const HprDev& P, FastSmem<NFFT>& sm, const FastState& st, const int i, const int slot,
                                                   const float* __restrict__ prev, const float* __restrict__ cur, bool full, bool fresh_tail,
                                                   const HprEmit& em, const float* next_hop, FastPeaks& peaks)
{
	constexpr int M = NFFT / 2;   // complex FFT length; also nwin
	constexpr int HC = M / 4;     // complex (packed even/odd) samples per hop
	const int tid = threadIdx.x;
	const int midp = P.midp;
	// The two FFT buffers keep their roles: the forward transform ends in `zres`, the split turns it into X IN PLACE,
	// the other buffer (`zoth`) holds the magnitude row while the masks are decided and then receives the masked
	// spectrum; the inverse stages ping-pong from there, so the last one (the overlap-add) reads `zres` - the buffer the
	// first forward stage of the next hop does NOT write.  No barrier is needed between two hops.
	float2* const a = sm.za;
	float2* const b = sm.zb;

	if (next_hop != nullptr && tid < (HC * 8) / 128)
		asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(next_hop) + tid * 128));

	// ---- A + B. sqrt-Hann window (hps.cu:452-462) fused into the forward FFT (hps.cu:465)
	const float2* const win2 = reinterpret_cast<const float2*>(P.window);
	const float2* const prev2 = reinterpret_cast<const float2*>(prev);
	const float2* const cur2 = reinterpret_cast<const float2*>(cur);
	// element j + r * NB of the packed frame, NB = M / 8: r = 0, 1 lie in the previous hop, r = 2, 3 in the current one
	// (r is a constant once the stage is unrolled, so the choice costs nothing)
	auto load_frame = [&](int j, int r, int /*p*/) -> float2 {
		constexpr int NB = M / 8;
		static_assert(HC == 2 * NB, "first stage must be radix 8");
		float2 x;
		if (r < 2)  // last use of that hop: streaming load, so it does not push the per-CTA scratch out of L2
			x = prev2 ? __ldcs(prev2 + j + r * NB) : make_float2(0.0f, 0.0f);
		else
			x = cur2[j + (r - 2) * NB];
		const float2 w = __ldg(win2 + j + r * NB);
		// scalar products: ptxas would contract a packed multiply into the butterfly's first additions (see
		// rfft_pack_masked), and the kernels that window through shared memory could not follow
		return make_float2(__fmul_rn(x.x, w.x), __fmul_rn(x.y, w.y));
	};
	float2* const zres = fft_pp_fused<M, NT, -1, true, false, false, false>(a, b, P.tw, tid, load_frame, NoFn{});
	float2* const zoth = zres == a ? b : a;

	// ---- C. split into the real-input spectrum X[0..M] (in place in zres, natural order: a thread owns its pair
	// (k, M - k)), magnitudes into the ring and into erow, which borrows zoth (hps.cu:469-472, 492-493)
	sm.erow = reinterpret_cast<float*>(zoth);
	{
		float* mag_row = st.mag_ring + (size_t)slot * st.ring_stride;
		float* const E = sm.erow;
		auto put = [&](int k, int kb, float2 Xa, float2 Xb) {
			const float ma = zen_cabs(Xa.x, Xa.y), mb = zen_cabs(Xb.x, Xb.y);
			mag_row[k] = ma;
			mag_row[kb] = mb;
			zres[k] = Xa;
			zres[kb] = Xb;
			E[fast_esw(midp + k)] = ma;
			E[fast_esw(midp + kb)] = mb;
			// mirrored borders: |X[-t]| = |X[t]|, |X[M+t]| = |X[M-t]|
			if (k >= 1 && k <= midp) {
				E[fast_esw(midp - k)] = ma;
				E[fast_esw(midp + M + k)] = mb;
			}
		};
		// pairs (k, M - k), k = 1 .. M/2 - 1, unrolled so that the addresses of one trip are constant offsets from the
		// previous one; bins 0 / M and M/2 (their own partners) are one thread's extra work
		constexpr int TRIPS = (M / 2 + NT - 1) / NT;
#pragma unroll
		for (int it = 0; it < TRIPS; ++it) {
			const int k = tid + it * NT;
			if (((M / 2) % NT == 0 || k < M / 2) && (it > 0 || k > 0)) {
				float2 Xa, Xb;
				rfft_split_pair(zres[k], zres[M - k], __ldg(&P.twr[k]), Xa, Xb);
				put(k, M - k, Xa, Xb);
			}
		}
		if (tid == 0) {
			const float2 Z0 = zres[0];
			put(0, M, make_float2(Z0.x + Z0.y, 0.0f), make_float2(Z0.x - Z0.y, 0.0f));
		}
		else if (tid == 32 % NT) {
			const float2 Xm = cconj(zres[M / 2]);
			put(M / 2, M / 2, Xm, Xm);
		}
	}
	__syncthreads();
	if (!full)
		return;

	// ---- F' + E'. time medians of eight bins per thread, thresholds, decisions by counting
	{
		// a mask is only computed for an output that is enabled; the residual mask 1 - (Mh + Mp) uses whatever the
		// other two hold, zeros included (hps.cu:498-567)
		const bool want_p = (P.out_flags & ZEN_OUTPUT_PERCUSSIVE) != 0;
		const bool want_h = (P.out_flags & ZEN_OUTPUT_HARMONIC) != 0;
		for (int g = tid; g < M / 8; g += NT) {
			const int k0 = 8 * g;
			float H[8];
			switch (P.n_taps) {
			case 1: fast_h8<1>(P, st, i, slot, k0, H); break;
			case 3: fast_h8<3>(P, st, i, slot, k0, H); break;
			case 5: fast_h8<5>(P, st, i, slot, k0, H); break;
			default: fast_h8<7>(P, st, i, slot, k0, H); break;
			}
			const float4* const erow4 = reinterpret_cast<const float4*>(sm.erow);
			const unsigned code = fast_decide_flags<8>(P, [&](int v) -> float4 { return erow4[fast_esw_chunk(2 * g + v)]; }, H);
			sm.codes[g] = (unsigned short)code;
		}
		// the Nyquist bin k = M: one warp, one tap per lane and round.  The LAST warp takes it: the scheduler favours the
		// higher warp ids, so it is the one with slack before the barrier.
		if (tid >= NT - 32) {
			const int lane = tid & 31;
			const float Hm = fast_h1(P, st, i, slot, M);
			const float tau = want_p ? thr_ratio_ge(Hm + ZEN_EPS, P.rule_p) : CUDART_INF_F;
			const float sig = want_h ? thr_ratio_le(Hm, P.rule_h) : -1.0f;
			const int L = P.Lp;
			int np = 0, nh = 0;
			for (int j0 = 0; j0 < L; j0 += 32) {
				const int j = j0 + lane;
				const float x = j < L ? sm.erow[fast_esw(M + j)] : -1.0f;   // window of bin M: elements M .. M + L - 1
				np += __popc(__ballot_sync(0xffffffffu, j < L && x >= tau));
				nh += __popc(__ballot_sync(0xffffffffu, j < L && (x + ZEN_EPS) <= sig));
			}
			const int need = L / 2 + 1;
			if (lane == 0)
				sm.codes[M / 8] = (unsigned short)((np >= need ? 1u : 0u) | ((nh >= need ? 1u : 0u) << 8));
		}
	}
	__syncthreads();

	// ---- G. per output: mask, inverse real FFT, overlap-add (hps.cu:498-579); order P, H, R as in the reference
	int last_o = -1;
#pragma unroll
	for (int oi = 0; oi < 3; ++oi) {
		const int o = oi == 0 ? 1 : (oi == 1 ? 0 : 2);
		if (P.out_flags & (1 << o)) last_o = o;
	}
#pragma unroll 1
	for (int oi = 0; oi < 3; ++oi) {
		const int o = oi == 0 ? 1 : (oi == 1 ? 0 : 2);
		if (!(P.out_flags & (1 << o)))
			continue;
		// masked spectrum packed for the M-point inverse transform, into zoth (the magnitude row is no longer needed;
		// X stays in zres)
		switch (o) {
		case 1: fast_pack<NFFT, NT, 1>(P, sm.codes, zres, zoth); break;
		case 0: fast_pack<NFFT, NT, 0>(P, sm.codes, zres, zoth); break;
		default: fast_pack<NFFT, NT, 2>(P, sm.codes, zres, zoth); break;
		}
		__syncthreads();
		// inverse FFT (hps.cu:522) whose last stage IS the overlap-add (hps.h:68-80): the thread that holds sample
		// pair n < HC of the frame also holds pair n + HC, i.e. it reads tail[n], emits, then writes the new tail[n]
		float* const tail = o == 0 ? st.tail[0] : (o == 1 ? st.tail[1] : st.tail[2]);
		float* const ea = o == 0 ? em.a[0] : (o == 1 ? em.a[1] : em.a[2]);
		float2* const tail2 = reinterpret_cast<float2*>(tail);
		float2* const ea2 = reinterpret_cast<float2*>(ea);
		const float cola = P.cola;
		float pk_local = 0.0f;
		auto ola = [&](int idx, int /*p*/, float2 v) {
			if (idx < HC) {
				const float2 t = fresh_tail ? make_float2(0.0f, 0.0f) : tail2[idx];
				const float2 r = up(pfma(pk(v), pk(cola, cola), pk(t)));
				if (ea2) {
					__stcs(ea2 + idx, r);  // written once, never re-read by the kernel
					if (PEAKS) pk_local = fmaxf(pk_local, fmaxf(fabsf(r.x), fabsf(r.y)));
				}
			}
			else {
				tail2[idx - HC] = up(pmul(pk(v), pk(cola, cola)));
			}
		};
		if (o == last_o) {
			// nobody needs X any more: ping-pong through its buffer, one barrier per stage, none at the end
			fft_pp_rest<M, NT, +1, 1, false, true, false, true, LAY_N>(zoth, zres, P.tw, tid, ola);
		}
		else {
			fft_inplace_last<M, NT, +1, 1, true, false>(zoth, P.tw, tid, ola);
			__syncthreads();  // the next output packs into zoth
		}
		if (PEAKS) {
			if (o == 0) peaks.v[0] = fmaxf(peaks.v[0], pk_local);
			else if (o == 1) peaks.v[1] = fmaxf(peaks.v[1], pk_local);
			else peaks.v[2] = fmaxf(peaks.v[2], pk_local);
		}
	}
}

// ---- one real-time hop split over a thread-block cluster ---------------------
// The resident kernel may run as a cluster of C CTAs (hpr_rt_kernel).  Every CTA takes the hop and runs the forward
// FFT redundantly (it is latency-, not throughput-bound); the per-bin work that follows - split into the real-input
// spectrum, |X|, time median, hard-mask decision by counting, masked spectrum - is divided by bin PAIRS (k, M-k):
// CTA r owns pairs [r*Q, (r+1)*Q), i.e. the bin ranges A = [a0, a1) and B = [b0, b1), and analyses a halo of
// midp + US bins around them for the frequency windows.  Each CTA keeps the |X| ring of its own bins only.  The masked
// spectrum of output o is written straight into the receive buffer of the CTA that owns the output (distributed
// shared memory); after one cluster barrier that CTA runs the inverse FFT and the overlap-add (hpr_split_synth).
// Hard mask, copy-border, lag 1 only: the configuration of HPRRealtime's default path (hps.cu:282-427).
struct HprSplit {
	int rank, C;
	int k0, k1;          // own pairs [k0, k1)
	int a0, a1, b0, b1;  // own bins
	float2* recv[3];     // per output (H, P, R): the owner's receive buffer, fpad_size(M) float2, or null when the output is off
	int owner[3];        // rank of the CTA that synthesises output o (-1: off)
};

__host__ __device__ inline void hpr_split_ranges(int M, int rank, int C, int& k0, int& k1, int& a0, int& a1, int& b0, int& b1)
{
	const int Q = (M / 2 + 1 + C - 1) / C;
	k0 = rank * Q < M / 2 + 1 ? rank * Q : M / 2 + 1;
	k1 = k0 + Q < M / 2 + 1 ? k0 + Q : M / 2 + 1;
	a0 = k0;
	a1 = k1;
	b0 = M - k1 + 1 > k1 ? M - k1 + 1 : k1;  // bin M/2 is its own partner
	b1 = M - k0 + 1;
	if (k0 >= k1) b0 = b1 = a1 = a0;
}

template <int NFFT, int NT, int US>
__device__ __forceinline__ void hpr_split_analyse(const HprDev& P, HprSmem<NFFT>& sm, const HprState& st, const int i,
                                                  const float* __restrict__ prev, const float* __restrict__ cur, float* cur_stash,
                                                  const HprTables& tb, const HprSplit& sp, float2* zpp, int recv_off, unsigned long long* stamps)
{
	(void)US;
	auto stamp = [&](int idx) {
		if (stamps && threadIdx.x == 0)
			stamps[idx] = (unsigned long long)clock64();
	};
	stamp(0);
	constexpr int M = NFFT / 2;
	constexpr int HC = M / 4;  // complex (packed even / odd) samples per hop
	const int tid = threadIdx.x;
	const int W = P.W;
	const int eoff = P.midp;
	for (int t = tid; t < P.n_taps; t += NT) {
		int j = i - P.tap_age[t];
		sm.taps[t] = j >= 0 ? (j % W) * (M + 1) : -1;
	}
	// ---- A + B. sqrt-Hann window fused into the first stage of the forward FFT (prev is in this CTA's shared memory,
	// cur in this or in the leader CTA's; every element of the hop is read by exactly one thread, which also keeps the
	// copy that becomes the next hop's `prev`).  The stages ping-pong between zbuf and zpp and END in zbuf.
	constexpr bool even_stages = fft_stage_count<M>() % 2 == 0;
	float2* const fa = even_stages ? zpp : sm.zbuf;
	float2* const fb = even_stages ? sm.zbuf : zpp;
	const float2* const prev2 = reinterpret_cast<const float2*>(prev);
	const float2* const cur2 = reinterpret_cast<const float2*>(cur);
	float2* const stash2 = reinterpret_cast<float2*>(cur_stash);
	const float2* const win2 = reinterpret_cast<const float2*>(tb.window);
	auto load_frame = [&](int j, int r, int /*p*/) -> float2 {
		constexpr int NB = M / 8;
		static_assert(HC == 2 * NB, "first stage must be radix 8");
		float2 x;
		if (r < 2)
			x = prev2[j + r * NB];
		else {
			x = cur2[j + (r - 2) * NB];
			if (stash2) stash2[j + (r - 2) * NB] = x;
		}
		const float2 w = win2[j + r * NB];
		return make_float2(__fmul_rn(x.x, w.x), __fmul_rn(x.y, w.y));
	};
	stamp(1);
	fft_pp_fused<M, NT, -1, true, false, true, false>(fa, fb, tb.tw, tid, load_frame, NoFn{});
	stamp(2);
	// ---- C. real-input spectrum and |X| of the own pairs and their halo (mirrored borders included)
	const int halo = P.midp;
	const int pl = max(0, sp.k0 - halo), ph = min(M / 2, sp.k1 - 1 + halo);
	{
		float* mag_row = st.mag_ring + (size_t)(i % W) * (M + 1);
		for (int k = pl + tid; k <= ph; k += NT) {
			float2 Xa, Xb;
			const int ka = k, kb = M - k;
			if (k == 0) {
				float2 Z0 = sm.zbuf[0];
				Xa = make_float2(Z0.x + Z0.y, 0.0f);
				Xb = make_float2(Z0.x - Z0.y, 0.0f);
			}
			else if (k == M / 2) {
				Xa = cconj(sm.zbuf[M / 2]);
				Xb = Xa;
			}
			else {
				rfft_split_pair(sm.zbuf[k], sm.zbuf[M - k], tb.twr[k], Xa, Xb);
			}
			const float ma = zen_cabs(Xa.x, Xa.y), mb = zen_cabs(Xb.x, Xb.y);
			mag_row[ka] = ma;
			sm.xbuf[ka] = Xa;
			sm.erow[eoff + ka] = ma;
			if (kb != ka) {
				mag_row[kb] = mb;
				sm.xbuf[kb] = Xb;
				sm.erow[eoff + kb] = mb;
			}
			// mirrored borders: |X[-t]| = |X[t]|, |X[M+t]| = |X[M-t]|
			if (k >= 1 && k <= P.midp) {
				sm.erow[eoff - k] = ma;
				sm.erow[eoff + M + k] = mb;
			}
		}
	}
	__syncthreads();
	stamp(3);
	stamp(4);
	// ---- F' + E'. time median (hps.cu:495, consumed row only) and hard-mask decision of the own bins: ONE bin per
	// thread.  This is the latency path: a hop is a chain of short phases, and a phase is as long as its slowest
	// thread, so the work is spread as thin as it goes (the batched kernel does the opposite: eight bins per thread,
	// taps shared, fewer instructions).  Codes as in hpr_iteration (bit0|bit1 P, bit2|bit3 H).
	{
		unsigned* codes = reinterpret_cast<unsigned*>(sm.prow);
		const int nt = P.n_taps;
		const bool want_p = (P.out_flags & ZEN_OUTPUT_PERCUSSIVE) != 0;
		const bool want_h = (P.out_flags & ZEN_OUTPUT_HARMONIC) != 0;
		const int L = P.Lp;
		const float need = (float)(L / 2 + 1);
		const float* const ring = st.mag_ring;
		auto tap = [&](int t, int k) -> float {
			const int off = sm.taps[t];
			return off >= 0 ? ring[off + k] : 0.0f;
		};
		auto hmed = [&](int k) -> float {
			switch (nt) {
			case 0: return 0.0f;
			case 1: return tap(0, k);
			case 3: return median_fixed<3>([&](int t) { return tap(t, k); });
			case 5: return median_fixed<5>([&](int t) { return tap(t, k); });
			case 7: return median_fixed<7>([&](int t) { return tap(t, k); });
			case 9: return median_fixed<9>([&](int t) { return tap(t, k); });
			case 11: return median_fixed<11>([&](int t) { return tap(t, k); });
			case 13: return median_fixed<13>([&](int t) { return tap(t, k); });
			default: return median_generic([&](int t) { return tap(t, k); }, nt);
			}
		};
		const int nA = sp.a1 - sp.a0, nB = sp.b1 - sp.b0;
		const int n_own = nA + nB;
		auto bin_of = [&](int idx) { return idx < nA ? sp.a0 + idx : sp.b0 + (idx - nA); };
		// counts of one bin, specialised on the enabled masks: two taps per packed add (FADD2), two independent chains
		auto count1 = [&](auto wp_c, auto wh_c, const float* E, float tau, float sig, float& cp, float& ch) {
			constexpr bool WP = decltype(wp_c)::value, WH = decltype(wh_c)::value;
			f32x2_t ap[2] = {pk(0.0f, 0.0f), pk(0.0f, 0.0f)}, ah[2] = {pk(0.0f, 0.0f), pk(0.0f, 0.0f)};
			int j = 0;
#pragma unroll 2
			for (; j + 3 < L; j += 4) {
#pragma unroll
				for (int c = 0; c < 2; ++c) {
					const float x0 = E[j + 2 * c], x1 = E[j + 2 * c + 1];
					if (WP) ap[c] = padd(ap[c], pk((x0 >= tau) ? 1.0f : 0.0f, (x1 >= tau) ? 1.0f : 0.0f));
					if (WH) ah[c] = padd(ah[c], pk(((x0 + ZEN_EPS) <= sig) ? 1.0f : 0.0f, ((x1 + ZEN_EPS) <= sig) ? 1.0f : 0.0f));
				}
			}
			for (; j < L; ++j) {
				const float x0 = E[j];
				if (WP) ap[0] = padd(ap[0], pk((x0 >= tau) ? 1.0f : 0.0f, 0.0f));
				if (WH) ah[0] = padd(ah[0], pk(((x0 + ZEN_EPS) <= sig) ? 1.0f : 0.0f, 0.0f));
			}
			const float2 a = up(padd(ap[0], ap[1])), h2 = up(padd(ah[0], ah[1]));
			cp = a.x + a.y;
			ch = h2.x + h2.y;
		};
		if (tid < n_own) {
			const int k = bin_of(tid);
			const float H = hmed(k);
			const float tau = want_p ? thr_ratio_ge(H + ZEN_EPS, P.rule_p) : CUDART_INF_F;
			const float sig = want_h ? thr_ratio_le(H, P.rule_h) : -1.0f;
			const float* const E = sm.erow + k;   // taps of bin k: erow[k .. k + L)
			float cp = 0.0f, ch = 0.0f;
			if (stamps && tid == 0) stamps[13] = (unsigned long long)clock64() + (unsigned long long)(tau > 1e30f);
			if (want_p && want_h) count1(std::true_type{}, std::true_type{}, E, tau, sig, cp, ch);
			else if (want_p) count1(std::true_type{}, std::false_type{}, E, tau, sig, cp, ch);
			else if (want_h) count1(std::false_type{}, std::true_type{}, E, tau, sig, cp, ch);
			const unsigned p0 = (want_p && cp >= need) ? 1u : 0u, h0 = (want_h && ch >= need) ? 1u : 0u;
			codes[k] = p0 * 3u | (h0 * 3u) << 2;
			if (stamps && tid == 0) stamps[14] = (unsigned long long)clock64() + (unsigned long long)p0;
		}
		// the few bins beyond one per thread (2049 bins do not divide by the CTAs' threads): one warp each, a tap per lane
		for (int e = 0; NT + e < n_own; ++e) {
			const int wid = tid >> 5, lane = tid & 31;
			if (wid != NT / 32 - 1 - (e % (NT / 32)))
				continue;
			const int k = bin_of(NT + e);
			const float H = hmed(k);
			const float tau = want_p ? thr_ratio_ge(H + ZEN_EPS, P.rule_p) : CUDART_INF_F;
			const float sig = want_h ? thr_ratio_le(H, P.rule_h) : -1.0f;
			int np = 0, nh = 0;
			for (int j0 = 0; j0 < L; j0 += 32) {
				const int j = j0 + lane;
				const float x = j < L ? sm.erow[k + j] : -1.0f;
				np += __popc(__ballot_sync(0xffffffffu, j < L && x >= tau));
				nh += __popc(__ballot_sync(0xffffffffu, j < L && (x + ZEN_EPS) <= sig));
			}
			if (lane == 0) {
				const unsigned p0 = (want_p && (float)np >= need) ? 1u : 0u, h0 = (want_h && (float)nh >= need) ? 1u : 0u;
				codes[k] = p0 * 3u | (h0 * 3u) << 2;
			}
		}
	}
	__syncthreads();
	stamp(5);
	// ---- G (first half). masked spectrum of the own pairs, packed for the M-point inverse FFT, into the owners' buffers
	{
		const unsigned* codes = reinterpret_cast<const unsigned*>(sm.prow);
		for (int k = sp.k0 + tid; k < sp.k1; k += NT) {
			const int kb = M - k;
			const unsigned ca = codes[k], cb = codes[kb];
			const float mpa = 0.5f * (float)((ca & 1u) + ((ca >> 1) & 1u)), mha = 0.5f * (float)(((ca >> 2) & 1u) + ((ca >> 3) & 1u));
			const float mpb = 0.5f * (float)((cb & 1u) + ((cb >> 1) & 1u)), mhb = 0.5f * (float)(((cb >> 2) & 1u) + ((cb >> 3) & 1u));
			const float2 Xa = sm.xbuf[k], Xb = sm.xbuf[kb];
			const float2 twk = tb.twr[k];
#pragma unroll
			for (int o = 0; o < 3; ++o) {
				if (!sp.recv[o])
					continue;
				float2* zb = sp.recv[o] + recv_off;
				const float ma = (o == 1) ? mpa : (o == 0 ? mha : 1.0f - (mha + mpa));  // hps.h:35-43
				const float mb = (o == 1) ? mpb : (o == 0 ? mhb : 1.0f - (mhb + mpb));
				if (k == 0) {
					zb[0] = rfft_pack_dc(Xa, ma, Xb, mb);
				}
				else if (k == M / 2) {
					zb[M / 2] = rfft_pack_mid(Xa, ma);
				}
				else {
					float2 Zk, Zmk;
					rfft_pack_masked(Xa, ma, Xb, mb, twk, Zk, Zmk);   // hps.h:58-66
					zb[k] = Zk;
					zb[kb] = Zmk;
				}
			}
		}
	}
}

// tagged emission of one hop held in pbuf (HOP floats of shared memory): see HprPack.  All NT threads must call it.
template <int NFFT, int NT>
__device__ __forceinline__ void hpr_pack_groups(const float* pbuf, uint4* pdst, unsigned ptag)
{
	constexpr int HOP = NFFT / 4;
	constexpr int NG = (HOP + 2) / 3, NLG = 4 * (HOP / 12);
	const int tid = threadIdx.x;
	for (int g0 = 0; g0 < NG; g0 += NT) {  // (every thread makes every trip: zen_group_key shuffles)
		const int g = g0 + tid;
		uint4 v = make_uint4(0u, 0u, 0u, 0u);
		if (g < NG) {
			v.x = __float_as_uint(pbuf[3 * g]);
			v.y = 3 * g + 1 < HOP ? __float_as_uint(pbuf[3 * g + 1]) : 0u;
			v.z = 3 * g + 2 < HOP ? __float_as_uint(pbuf[3 * g + 2]) : 0u;
		}
		v.w = ptag ^ zen_group_key(v.x, v.y, v.z, g, NLG);
		if (g < NG)
			asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(pdst + g), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
			             : "memory");
	}
}

// G (second half), run by the CTA that owns output o after the cluster barrier: inverse FFT of the received masked
// spectrum whose last stage is the overlap-add (hps.h:68-80: the thread that holds sample pair n < HC of the frame also
// holds pair n + HC, so it reads tail[n], emits, then writes the new tail[n]), then the tagged emission.
template <int NFFT, int NT>
__device__ __forceinline__ void hpr_split_synth(const HprDev& P, float2* zb, float2* zpp, float* tail, float* ea, float* eb, float* pbuf,
                                                uint4* pdst, unsigned ptag, const HprTables& tb, unsigned long long* stamps)
{
	constexpr int M = NFFT / 2, HC = M / 4;
	if (stamps && threadIdx.x == 0) stamps[6] = (unsigned long long)clock64();
	float2* const tail2 = reinterpret_cast<float2*>(tail);
	float2* const ea2 = reinterpret_cast<float2*>(ea);
	float2* const eb2 = reinterpret_cast<float2*>(eb);
	float2* const pb2 = reinterpret_cast<float2*>(pbuf);
	const float cola = P.cola;
	auto ola = [&](int idx, int /*p*/, float2 v) {
		if (idx < HC) {
			const float2 t = tail2[idx];
			const float2 r = up(pfma(pk(v), pk(cola, cola), pk(t)));
			if (ea2) __stcs(ea2 + idx, r);
			if (eb2) __stcs(eb2 + idx, r);
			if (pdst) pb2[idx] = r;
		}
		else {
			tail2[idx - HC] = up(pmul(pk(v), pk(cola, cola)));
		}
	};
	fft_pp_rest<M, NT, +1, 1, false, true, true, true, LAY_N>(zb, zpp, tb.tw, threadIdx.x, ola);
	if (stamps && threadIdx.x == 0) stamps[7] = (unsigned long long)clock64();
	__syncthreads();
	if (pdst) hpr_pack_groups<NFFT, NT>(pbuf, pdst, ptag);
	__syncthreads();
	if (stamps && threadIdx.x == 0) stamps[8] = (unsigned long long)clock64();
}

}  // namespace zen_b200
