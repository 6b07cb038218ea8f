// Median selection primitives.
//
// Frequency axis (long windows, L up to 255): one warp keeps the current window
// SORTED in registers, K consecutive ranks per lane ("blocked" layout, rank
// p = lane*K + r), and slides it one bin at a time.  Removing the outgoing
// value o and inserting the incoming value v is a shift of the ranks that lie
// between them, which in a sorted array reduces to one min (or max) against
// the neighbouring rank per register plus a single warp shuffle for the lane
// boundary:
//     v >= o :  S'[p] = (o <= S[p] <= v) ? min(S[p+1], v) : S[p]
//     v <  o :  S'[p] = (v <= S[p] <= o) ? max(S[p-1], v) : S[p]
// Equal values are bit-identical, so which copy is removed does not matter: the
// output is always one of the input bit patterns, i.e. what NPP's median
// returns (nppiFilterMedian_32f_C1R as driven by libzen/mfilt.h:233-267).
// The window is padded below with -inf so that the median rank is register 0
// of a fixed lane, and above with +inf up to 32*K ranks.
//
// Time axis (short windows): per-thread sorting network over registers.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <type_traits>

namespace zen_b200 {

// --- float total order via monotone integer keys (only needed by the standalone
// filter, where inputs may be negative; the HPR path filters magnitudes >= 0) ---
__device__ __forceinline__ unsigned f2key(float f)
{
	unsigned u = __float_as_uint(f);
	return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key2f(unsigned k)
{
	unsigned u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
	return __uint_as_float(u);
}

// value traits: float for magnitudes, unsigned keys for arbitrary floats
template <typename T> struct SortVal;
template <> struct SortVal<float> {
	static __device__ __forceinline__ float lo() { return -CUDART_INF_F; }
	static __device__ __forceinline__ float hi() { return CUDART_INF_F; }
	static __device__ __forceinline__ float mn(float a, float b) { return fminf(a, b); }
	static __device__ __forceinline__ float mx(float a, float b) { return fmaxf(a, b); }
};
template <> struct SortVal<unsigned> {
	static __device__ __forceinline__ unsigned lo() { return 0u; }
	static __device__ __forceinline__ unsigned hi() { return 0xffffffffu; }
	static __device__ __forceinline__ unsigned mn(unsigned a, unsigned b) { return min(a, b); }
	static __device__ __forceinline__ unsigned mx(unsigned a, unsigned b) { return max(a, b); }
};

template <int K, typename T = float>
struct WarpSortedWindow {
	using V = SortVal<T>;
	T S[K];
	int pad_lo;    // number of -inf ranks below the window
	int med_lane;  // lane whose register 0 holds the median rank

	// Load window src[0..L) (L odd, L + pad_lo <= 32*K) and sort it.
	template <typename Load>
	__device__ __forceinline__ void init(Load load, int L, int lane)
	{
		const int mid = L >> 1;
		pad_lo = (K - (mid % K)) % K;
		med_lane = (mid + pad_lo) / K;
#pragma unroll
		for (int r = 0; r < K; ++r) {
			int p = lane * K + r - pad_lo;
			S[r] = (p < 0) ? V::lo() : (p < L ? load(p) : V::hi());
		}
		// bitonic sort over n = 32*K ranks, ascending
#pragma unroll
		for (int k = 2; k <= 32 * K; k <<= 1) {
#pragma unroll
			for (int j = k >> 1; j > 0; j >>= 1) {
				if (j >= K) {
					const int lane_mask = j / K;
#pragma unroll
					for (int r = 0; r < K; ++r) {
						int p = lane * K + r;
						T other = __shfl_xor_sync(0xffffffffu, S[r], lane_mask);
						bool up = ((p & k) == 0);        // ascending block
						bool lower = ((p & j) == 0);     // I am the lower index of the pair
						T mn = V::mn(S[r], other), mx = V::mx(S[r], other);
						S[r] = (up == lower) ? mn : mx;
					}
				}
				else {
#pragma unroll
					for (int r = 0; r < K; ++r) {
						if ((r & j) == 0) {
							int p = lane * K + r;
							bool up = ((p & k) == 0);
							T a = S[r], b = S[r | j];
							T mn = V::mn(a, b), mx = V::mx(a, b);
							S[r] = up ? mn : mx;
							S[r | j] = up ? mx : mn;
						}
					}
				}
			}
		}
	}

	__device__ __forceinline__ T median_reg() const { return S[0]; }

	// remove one copy of o (which must be in the window), insert v
	__device__ __forceinline__ void slide(T o, T v, int lane)
	{
		if (v >= o) {
			T edge = __shfl_down_sync(0xffffffffu, S[0], 1);
			if (lane == 31)
				edge = V::hi();
#pragma unroll
			for (int r = 0; r < K; ++r) {
				T nxt = (r + 1 < K) ? S[r + 1] : edge;
				bool in = (S[r] >= o) && (S[r] <= v);
				S[r] = in ? V::mn(nxt, v) : S[r];
			}
		}
		else {
			T edge = __shfl_up_sync(0xffffffffu, S[K - 1], 1);
			if (lane == 0)
				edge = V::lo();
#pragma unroll
			for (int r = K - 1; r >= 0; --r) {
				T prv = (r > 0) ? S[r - 1] : edge;
				bool in = (S[r] <= o) && (S[r] >= v);
				S[r] = in ? V::mx(prv, v) : S[r];
			}
		}
	}
};

// out[s] = median(E[s .. s+L)) for s in [s0, s1), one warp, E and out in shared memory
template <int K, typename T = float>
__device__ __forceinline__ void warp_sliding_median(const T* __restrict__ E, T* __restrict__ out,
                                                    int s0, int s1, int L, int lane)
{
	if (s0 >= s1)
		return;
	WarpSortedWindow<K, T> w;
	const T* base = E + s0;
	w.init([&](int p) { return base[p]; }, L, lane);
	const int ml = w.med_lane;
	if (lane == ml)
		out[s0] = w.median_reg();
	for (int s = s0 + 1; s < s1; ++s) {
		T o = E[s - 1];
		T v = E[s + L - 1];
		w.slide(o, v, lane);
		if (lane == ml)
			out[s] = w.median_reg();
	}
}

// runtime dispatch on the number of registers per lane
template <typename T>
__device__ __forceinline__ void warp_sliding_median_dyn(int K, const T* E, T* out, int s0, int s1, int L, int lane)
{
	switch (K) {
	case 1: warp_sliding_median<1, T>(E, out, s0, s1, L, lane); break;
	case 2: warp_sliding_median<2, T>(E, out, s0, s1, L, lane); break;
	case 4: warp_sliding_median<4, T>(E, out, s0, s1, L, lane); break;
	default: warp_sliding_median<8, T>(E, out, s0, s1, L, lane); break;
	}
}

// smallest K in {1,2,4,8} whose 32*K ranks hold L values plus the low padding; 0 if none
static inline int sliding_K_for(int L)
{
	for (int K = 1; K <= 8; K <<= 1) {
		int mid = L / 2;
		int pad = (K - (mid % K)) % K;
		if (L + pad <= 32 * K)
			return K;
	}
	return 0;
}

// ---- windows up to 63 taps: one sorted window PER THREAD, in registers ----
// Each thread owns a run of consecutive outputs.  Its window lives in C
// registers (C compile-time, L <= C-1 run-time), padded below with -inf so the
// median is always S[C/2] and above with +inf.  The first window is sorted
// with Batcher's odd-even merge network; every further output removes the
// outgoing value and inserts the incoming one with 4 ALU ops per register and
// no data-dependent indexing:
//     T[p]  = S[p] < o ? S[p] : S[p+1]            (drop one copy of o)
//     S'[p] = max(T[p-1], min(T[p], v))           (insert v)
// All C updates of a slide are independent (ILP), nothing is shuffled and the
// only memory traffic is two shared-memory loads and one store per output.
template <int I, int End, int Step, typename F>
__device__ __forceinline__ void static_for(F&& f)
{
	if constexpr (I < End) {
		f(std::integral_constant<int, I>{});
		static_for<I + Step, End, Step>(f);
	}
}

template <int I, int J, int C>
__device__ __forceinline__ void ce_static(float (&v)[C])
{
	if constexpr (J < C) {  // ranks >= C are virtual +inf: the exchange would be a no-op
		float a = v[I], b = v[J];
		v[I] = fminf(a, b);
		v[J] = fmaxf(a, b);
	}
}

// Batcher odd-even merge of the subsequence LO, LO+R, LO+2R, ... (N elements span)
template <int LO, int N, int R, int C>
__device__ __forceinline__ void oe_merge(float (&v)[C])
{
	constexpr int step = R * 2;
	if constexpr (step < N) {
		oe_merge<LO, N, step, C>(v);
		oe_merge<LO + R, N, step, C>(v);
		static_for<LO + R, LO + N - R, step>([&](auto i) { ce_static<decltype(i)::value, decltype(i)::value + R, C>(v); });
	}
	else {
		ce_static<LO, LO + R, C>(v);
	}
}

template <int LO, int N, int C>
__device__ __forceinline__ void oe_sort(float (&v)[C])
{
	if constexpr (N > 1 && LO < C) {
		oe_sort<LO, N / 2, C>(v);
		oe_sort<LO + N / 2, N / 2, C>(v);
		oe_merge<LO, N, 1, C>(v);
	}
}

constexpr int next_pow2(int n)
{
	int p = 1;
	while (p < n) p <<= 1;
	return p;
}

template <int C>
__device__ __forceinline__ void thread_window_slide(float (&S)[C], float o, float v)
{
	float prevT = -CUDART_INF_F;
#pragma unroll
	for (int p = 0; p < C; ++p) {
		float nxt = (p + 1 < C) ? S[p + 1] : CUDART_INF_F;
		float t = (S[p] < o) ? S[p] : nxt;
		S[p] = fmaxf(prevT, fminf(t, v));
		prevT = t;
	}
}

// out[s] = median(E[s .. s+L)) for s in [s0, s1); E, out in shared memory; L odd, L <= C-1
template <int C>
__device__ __forceinline__ void thread_sliding_median(const float* __restrict__ E, float* __restrict__ out, int s0, int s1, int L)
{
	if (s0 >= s1)
		return;
	const int pad_lo = C / 2 - (L >> 1);
	float S[C];
#pragma unroll
	for (int p = 0; p < C; ++p) {
		int idx = p - pad_lo;
		S[p] = idx < 0 ? -CUDART_INF_F : (idx < L ? E[s0 + idx] : CUDART_INF_F);
	}
	oe_sort<0, next_pow2(C), C>(S);
	out[s0] = S[C / 2];
	const float* po = E + s0;
	const float* pv = E + s0 + L;
	float* pw = out + s0 + 1;
	for (int n = s1 - s0 - 1; n > 0; --n) {
		float o = *po++;
		float v = *pv++;
		thread_window_slide<C>(S, o, v);
		*pw++ = S[C / 2];
	}
}

// Same sliding window, taps fetched through `get(j)` (j-th sample of the run's extended line) and results
// handed to `put(q, value)`: used by the standalone filter kernels, whose taps live in global or shared
// memory with arbitrary strides.  Output q is the median of get(q) .. get(q+L-1).
template <int C, typename Get, typename Put>
__device__ __forceinline__ void thread_sliding_run(Get get, Put put, int n_out, int L)
{
	if (n_out <= 0)
		return;
	const int pad_lo = C / 2 - (L >> 1);
	float S[C];
#pragma unroll
	for (int p = 0; p < C; ++p) {
		int idx = p - pad_lo;
		S[p] = idx < 0 ? -CUDART_INF_F : (idx < L ? get(idx) : CUDART_INF_F);
	}
	oe_sort<0, next_pow2(C), C>(S);
	put(0, S[C / 2]);
	for (int q = 1; q < n_out; ++q) {
		float o = get(q - 1);
		float v = get(q + L - 1);
		thread_window_slide<C>(S, o, v);
		put(q, S[C / 2]);
	}
}

__device__ __forceinline__ void thread_sliding_median_dyn(int C, const float* E, float* out, int s0, int s1, int L)
{
	switch (C) {
	case 8: thread_sliding_median<8>(E, out, s0, s1, L); break;
	case 16: thread_sliding_median<16>(E, out, s0, s1, L); break;
	case 24: thread_sliding_median<24>(E, out, s0, s1, L); break;
	case 32: thread_sliding_median<32>(E, out, s0, s1, L); break;
	default: thread_sliding_median<48>(E, out, s0, s1, L); break;
	}
}

// smallest register capacity for an L-tap window (L <= C-1); 0 if L > 47 (the
// warp-resident window takes over: 64+ registers per thread would cost occupancy)
static inline int thread_window_capacity(int L)
{
	const int caps[5] = {8, 16, 24, 32, 48};
	for (int i = 0; i < 5; ++i)
		if (L <= caps[i] - 1)
			return caps[i];
	return 0;
}

// ---- short windows: sort L registers, return the middle one ----
template <int L>
__device__ __forceinline__ float median_regs(float (&v)[L])
{
	// odd-even transposition sort: L passes, fully unrolled
#pragma unroll
	for (int pass = 0; pass < L; ++pass) {
#pragma unroll
		for (int i = (pass & 1); i + 1 < L; i += 2) {
			float a = v[i], b = v[i + 1];
			v[i] = fminf(a, b);
			v[i + 1] = fmaxf(a, b);
		}
	}
	return v[L / 2];
}

// generic selection for any L: rank counting, values fetched through `get`
template <typename Get>
__device__ __forceinline__ float median_generic(Get get, int L)
{
	const int mid = L >> 1;
	float result = 0.0f;
	for (int i = 0; i < L; ++i) {
		float x = get(i);
		int less = 0, equal = 0;
		for (int j = 0; j < L; ++j) {
			float y = get(j);
			less += (y < x);
			equal += (y == x);
		}
		if (less <= mid && mid < less + equal)
			result = x;
	}
	return result;
}

}  // namespace zen_b200
