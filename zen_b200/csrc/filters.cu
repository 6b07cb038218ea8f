// Standalone median / box filters over a time x freq matrix:
// MedianFilterGPU::filter (libzen/mfilt.h:227-267) and BoxFilterGPU::filter
// (libzen/box.h:182-213) without NPP, without the padded scratch copies
// (nppiCopyWrapBorder + nppiCopy): the wrap border is index arithmetic.
//
// Window rules (pinned bit-exactly against NPP on a B200, tests/golden/npp_median.npz):
//   copy_bord      : out[i] = median(src[(i - mid + t) mod dim]), every cell written
//   TimeCausal     : rows r in [L, T)          taps r-L .. r-1
//   TimeAnticausal : rows r in [mid, mid+T-L)  taps r-mid .. r+mid
//   Frequency      : cols c in [0, F-L)        taps c .. c+L-1
// with L = filter_len made odd, mid = L/2; other cells of dst are left untouched.
#include "median_select.cuh"
#include "zen_common.cuh"

using namespace zen_b200;

namespace {

struct AxisGeom {
	int T, F;
	int axis;     // 0: filter along time (rows), 1: along frequency (columns)
	int first;    // first written index along the axis
	int n_out;    // number of written indices along the axis
	int tap_off;  // tap t of output a reads index a + tap_off + t
	int wrap;     // taps wrap modulo the axis length
	int L;
};

// rows (or outputs) beyond the 65535 limit of grid.y fold into grid.z: rows_grid(n) launches, grid_row() reads back
__device__ __forceinline__ long grid_row() { return (long)blockIdx.y + (long)blockIdx.z * gridDim.y; }
static inline dim3 rows_grid(unsigned gx, long n)
{
	const unsigned gy = (unsigned)(n < 32768 ? (n < 1 ? 1 : n) : 32768);
	return dim3(gx, gy, (unsigned)((n + gy - 1) / gy));
}

__device__ __forceinline__ int wrap_idx(int i, int dim)
{
	i %= dim;
	return i < 0 ? i + dim : i;
}

// ---- windows up to 47 taps: one sorted window per thread, sliding along the filtered axis ----
// (thread_sliding_run in median_select.cuh: 4 ALU ops per register and output, C = capacity >= L + 1).
// Time axis: a thread owns one column and a run of consecutive output rows (8..64, shorter when the matrix is
// too small to fill the GPU otherwise); lanes are adjacent
// columns, so every load is coalesced and the tap that leaves is re-read from L1/L2.
// Frequency axis: a CTA stages one row segment in shared memory with coalesced loads; a thread owns RUN_F
// consecutive columns (odd, so the lanes' strided shared-memory accesses hit distinct banks); results go
// back through shared memory for a coalesced store.
constexpr int RUN_F = 17;
constexpr int SM_NT = 128;

template <int C>
__global__ void __launch_bounds__(256) median_run_time_kernel(const float* __restrict__ src, float* __restrict__ dst, AxisGeom g,
                                                              int run_t)
{
	const int c = blockIdx.x * blockDim.x + threadIdx.x;
	const int q0 = blockIdx.y * run_t;
	if (c >= g.F)
		return;
	const int n_out = min(run_t, g.n_out - q0);
	const int base = g.first + q0 + g.tap_off;
	auto get = [&](int j) -> float {
		int i = base + j;
		if (g.wrap) i = wrap_idx(i, g.T);
		return __ldg(src + (size_t)i * g.F + c);
	};
	auto put = [&](int q, float v) { dst[(size_t)(g.first + q0 + q) * g.F + c] = v; };
	thread_sliding_run<C>(get, put, n_out, g.L);
}

template <int C>
__global__ void __launch_bounds__(SM_NT) median_run_freq_kernel(const float* __restrict__ src, float* __restrict__ dst, AxisGeom g)
{
	extern __shared__ float run_smem[];
	constexpr int CHUNK = SM_NT * RUN_F;
	float* E = run_smem;                    // CHUNK + L - 1
	float* O = run_smem + CHUNK + g.L + 3;  // CHUNK
	const long r = grid_row();
	if (r >= g.T)
		return;
	const int q0 = blockIdx.x * CHUNK;
	const int nq = min(CHUNK, g.n_out - q0);
	const float* row = src + (size_t)r * g.F;
	for (int t = threadIdx.x; t < nq + g.L - 1; t += SM_NT) {
		int i = g.first + q0 + g.tap_off + t;
		if (g.wrap) i = wrap_idx(i, g.F);
		E[t] = __ldg(row + i);
	}
	__syncthreads();
	const int s0 = threadIdx.x * RUN_F;
	const int n_out = min(RUN_F, nq - s0);
	auto get = [&](int j) -> float { return E[s0 + j]; };
	auto put = [&](int q, float v) { O[s0 + q] = v; };
	thread_sliding_run<C>(get, put, n_out, g.L);
	__syncthreads();
	for (int t = threadIdx.x; t < nq; t += SM_NT)
		dst[(size_t)r * g.F + g.first + q0 + t] = O[t];
}

// ---- windows up to 15 taps: selection networks shared by two neighbouring outputs ----
// Outputs q and q + 1 of an L-tap filter (L = 2 m + 1) share 2 m taps.  If lo / hi are the taps of rank m - 1 / m among
// those 2 m, the median of the shared taps plus ONE more tap x is min(max(x, lo), hi): x itself when it falls between
// them, the nearer bound otherwise.  lo and hi come out of Batcher's odd-even merge sorting network over the shared taps
// with every comparator that does not feed rank m - 1 or m removed by the compiler (everything lives in registers with
// constant indices): about 13 min/max pairs per output at L = 11 against 48 ALU operations for the sliding sorted
// window, so these kernels are bound by memory, not by the FMNMX rate.  A median is still a selection: the output is the
// bit pattern of one of the taps.
template <int N>
struct OemNetwork {
	int n = 0;
	unsigned char a[N * N + 4] = {}, b[N * N + 4] = {};
	constexpr OemNetwork()
	{
		for (int p = 1; p < N; p *= 2)
			for (int k = p; k >= 1; k /= 2)
				for (int j = k % p; j + k < N; j += 2 * k)
					for (int i = 0; i < k && i + j + k < N; ++i)
						if ((i + j) / (2 * p) == (i + j + k) / (2 * p)) {
							a[n] = (unsigned char)(i + j);
							b[n] = (unsigned char)(i + j + k);
							++n;
						}
	}
};

// the two middle taps of the 2 m = L - 1 shared ones w[1] .. w[L - 1]
template <int L>
__device__ __forceinline__ void pair_bounds(const float (&w)[L + 1], float& lo, float& hi)
{
	constexpr int N = L - 1;
	constexpr OemNetwork<N> net;
	float c[N];
#pragma unroll
	for (int i = 0; i < N; ++i)
		c[i] = w[i + 1];
#pragma unroll
	for (int i = 0; i < net.n; ++i) {
		const float x = c[net.a[i]], y = c[net.b[i]];
		c[net.a[i]] = fminf(x, y);
		c[net.b[i]] = fmaxf(x, y);
	}
	lo = c[N / 2 - 1];
	hi = c[N / 2];
}

// What a thread does with the L + 1 taps w[0 .. L] of an output pair: a = f(w[0 .. L - 1]), b = f(w[1 .. L]).
template <int L>
struct MedianPairOp {
	static __device__ __forceinline__ void apply(const float (&w)[L + 1], float& a, float& b)
	{
		float lo, hi;
		pair_bounds<L>(w, lo, hi);
		a = fminf(fmaxf(w[0], lo), hi);
		b = fminf(fmaxf(w[L], lo), hi);
	}
};

// x / L for a small integer L, correctly rounded like the IEEE division it replaces: q = x r with r = RN(1 / L), then
// two residual corrections (the residual fma(-q, L, x) is exact); inf and NaN sums pass through untouched.
template <int L>
__device__ __forceinline__ float div_by_len(float x)
{
	constexpr float r = 1.0f / (float)L;
	float q = x * r;
	if (fabsf(q) < CUDART_INF_F && fabsf(x) > 1e-30f) {
		q = __fmaf_rn(__fmaf_rn(-q, (float)L, x), r, q);
		q = __fmaf_rn(__fmaf_rn(-q, (float)L, x), r, q);
	}
	else if (fabsf(q) < CUDART_INF_F)
		q = __fdiv_rn(x, (float)L);  // results near the subnormal range: the plain division
	return q;
}

// Box filter: the two windows share the sum of their L - 1 common taps; nothing is ever subtracted (see box_freq_kernel)
template <int L>
struct BoxPairOp {
	static __device__ __forceinline__ void apply(const float (&w)[L + 1], float& a, float& b)
	{
		float s = w[1];
#pragma unroll
		for (int i = 2; i < L; ++i)
			s += w[i];
		a = div_by_len<L>(w[0] + s);
		b = div_by_len<L>(s + w[L]);
	}
};

// Time axis: a thread owns one column and a run of consecutive output rows; lanes are adjacent columns, so every load
// and store is coalesced.  The window of L + 1 rows slides through registers two rows per step; the rows of the next
// step are fetched before the network of this one runs.
// EDGE: the run touches the last row of the matrix (wrap to row 0 with copy_bord, stop there without)
template <int L, bool EDGE, class OP>
__device__ __forceinline__ void pair_time_run(const float* __restrict__ p, float* __restrict__ out, int n_out, int row, int T, size_t stride, bool wrap)
{
	const size_t rewind = (size_t)(T - 1) * stride;
	auto next = [&]() -> float {
		const float v = __ldg(p);
		if (!EDGE)
			p += stride;
		else if (row < T - 1) {
			p += stride;
			++row;
		}
		else if (wrap) {
			p -= rewind;
			row = 0;
		}
		return v;
	};
	float w[L + 1];
#pragma unroll
	for (int i = 0; i < L - 1; ++i)
		w[i] = next();
	float n0 = next(), n1 = next();
#pragma unroll((L + 1) / 2)
	for (int q = 0; q < n_out; q += 2) {
		w[L - 1] = n0;
		w[L] = n1;
		n0 = next();
		n1 = next();
		float a, b;
		OP::apply(w, a, b);
		out[0] = a;
		if (q + 1 < n_out) out[stride] = b;
		out += 2 * stride;
#pragma unroll
		for (int i = 0; i < L - 1; ++i)
			w[i] = w[i + 2];
	}
}

template <int L, class OP>
__global__ void __launch_bounds__(256) pair_time_kernel(const float* __restrict__ src, float* __restrict__ dst, AxisGeom g, int run_t)
{
	const int c = blockIdx.x * blockDim.x + threadIdx.x;
	const long q0 = grid_row() * run_t;
	if (c >= g.F || q0 >= g.n_out)
		return;
	const int n_out = (int)min((long)run_t, g.n_out - q0);
	// tap j of this run is row (first + q0 + tap_off + j), wrapped; a pointer walks the rows without a multiplication
	// or a division per load.  The walk reads up to three rows past the run's last tap (the next pair is fetched
	// ahead; an odd run has half a pair too many): a run that would leave the matrix that way takes the EDGE path,
	// where the pointer wraps to row 0 (copy_bord) or stays on the last row (those taps only feed outputs that are
	// not stored).
	int row = (int)(g.first + q0 + g.tap_off);
	if (g.wrap) row = wrap_idx(row, g.T);
	const float* p = src + (size_t)row * g.F + c;
	float* out = dst + (size_t)(g.first + q0) * g.F + c;
	if ((long)row + n_out + L + 3 <= g.T)
		pair_time_run<L, false, OP>(p, out, n_out, row, g.T, (size_t)g.F, false);
	else
		pair_time_run<L, true, OP>(p, out, n_out, row, g.T, (size_t)g.F, g.wrap != 0);
}

// Frequency axis: a CTA stages one row segment (+ L - 1 taps of halo, wrap by index arithmetic) in shared memory with
// coalesced loads; a thread takes output pairs (2 p, 2 p + 1), p = thread, thread + 256, ...: its L + 1 taps are
// consecutive floats at an even offset (8-byte shared loads, no bank conflicts), its two outputs one 8-byte store.
constexpr int PAIR_CHUNK = 4096;
constexpr int PAIR_NT = 256;

template <int L, class OP>
__global__ void __launch_bounds__(PAIR_NT) pair_freq_kernel(const float* __restrict__ src, float* __restrict__ dst, AxisGeom g)
{
	__shared__ __align__(16) float E[PAIR_CHUNK + 16];
	const long r = grid_row();
	if (r >= g.T)
		return;
	const int q0 = blockIdx.x * PAIR_CHUNK;
	const int nq = min(PAIR_CHUNK, g.n_out - q0);
	const float* row = src + (size_t)r * g.F;
	const int i0 = g.first + q0 + g.tap_off;
	const int n_used = nq + L - 1;
	// staging: every thread issues all its loads before it stores the first (17 independent 4-byte loads in flight per
	// thread is what keeps this kernel on the memory roofline); only the first and the last chunk of a row wrap
	constexpr int NLD = (PAIR_CHUNK + 16 + PAIR_NT - 1) / PAIR_NT;
	float v[NLD];
	if (i0 >= 0 && i0 + n_used <= g.F) {
		const float* p0 = row + i0 + threadIdx.x;
#pragma unroll
		for (int k = 0; k < NLD; ++k)
			v[k] = threadIdx.x + k * PAIR_NT < n_used ? __ldg(p0 + k * PAIR_NT) : 0.0f;
	}
	else {
#pragma unroll
		for (int k = 0; k < NLD; ++k) {
			const int t = threadIdx.x + k * PAIR_NT;
			int i = i0 + t;
			if (g.wrap) i = wrap_idx(i, g.F);
			v[k] = t < n_used ? __ldg(row + i) : 0.0f;
		}
	}
#pragma unroll
	for (int k = 0; k < NLD; ++k)
		if (threadIdx.x + k * PAIR_NT < PAIR_CHUNK + 16) E[threadIdx.x + k * PAIR_NT] = v[k];
	__syncthreads();
	float* orow = dst + (size_t)r * g.F + g.first + q0;
	const bool vec = ((reinterpret_cast<uintptr_t>(orow) & 7u) == 0u);
	for (int p = threadIdx.x; 2 * p < nq; p += PAIR_NT) {
		float w[L + 1];
#pragma unroll
		for (int i = 0; i < (L + 1) / 2; ++i) {
			const float2 v = *reinterpret_cast<const float2*>(E + 2 * p + 2 * i);
			w[2 * i] = v.x;
			w[2 * i + 1] = v.y;
		}
		float a, b;
		OP::apply(w, a, b);
		if (2 * p + 1 < nq) {
			if (vec)
				*reinterpret_cast<float2*>(orow + 2 * p) = make_float2(a, b);
			else {
				orow[2 * p] = a;
				orow[2 * p + 1] = b;
			}
		}
		else
			orow[2 * p] = a;
	}
}

__global__ void __launch_bounds__(256) copy_axis_kernel(const float* __restrict__ src, float* __restrict__ dst, AxisGeom g)
{
	// L == 1: the median of one tap is the tap
	const int c = blockIdx.x * blockDim.x + threadIdx.x;
	const long r = grid_row();
	if (c >= g.F || r >= g.T)
		return;
	if (g.axis == 1) {
		if (c < g.first || c >= g.first + g.n_out)
			return;
		int i = c + g.tap_off;
		if (g.wrap) i = wrap_idx(i, g.F);
		dst[(size_t)r * g.F + c] = __ldg(src + (size_t)r * g.F + i);
	}
	else {
		if (r < g.first || r >= g.first + g.n_out)
			return;
		int i = r + g.tap_off;
		if (g.wrap) i = wrap_idx(i, g.T);
		dst[(size_t)r * g.F + c] = __ldg(src + (size_t)i * g.F + c);
	}
}

// ---- long windows along frequency: warp-resident sorted window that slides ----
// One CTA = one row x one chunk of outputs; the chunk (+L-1 halo) is staged in
// shared memory as order-preserving unsigned keys.
constexpr int SL_CHUNK = 2048;
constexpr int SL_NT = 256;

__global__ void __launch_bounds__(SL_NT) median_slide_freq_kernel(const float* __restrict__ src, float* __restrict__ dst, AxisGeom g, int K)
{
	extern __shared__ unsigned sl_smem[];
	unsigned* E = sl_smem;                       // SL_CHUNK + L - 1
	unsigned* O = sl_smem + SL_CHUNK + g.L + 3;  // SL_CHUNK
	const long r = grid_row();
	if (r >= g.T)
		return;
	const int q0 = blockIdx.x * SL_CHUNK;
	const int nq = min(SL_CHUNK, g.n_out - q0);
	const float* row = src + (size_t)r * g.F;
	for (int t = threadIdx.x; t < nq + g.L - 1; t += SL_NT) {
		int i = g.first + q0 + g.tap_off + t;
		if (g.wrap) i = wrap_idx(i, g.F);
		E[t] = f2key(__ldg(row + i));
	}
	__syncthreads();
	constexpr int NW = SL_NT / 32;
	const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int R = (nq + NW - 1) / NW;
	const int s0 = wid * R, s1 = min(nq, s0 + R);
	warp_sliding_median_dyn<unsigned>(K, E, O, s0, s1, g.L, lane);
	__syncthreads();
	for (int t = threadIdx.x; t < nq; t += SL_NT)
		dst[(size_t)r * g.F + g.first + q0 + t] = key2f(O[t]);
}

// ---- anything else (long windows along time): rank counting per output ----
__global__ void __launch_bounds__(256) median_generic_kernel(const float* __restrict__ src, float* __restrict__ dst, AxisGeom g)
{
	const int other = blockIdx.x * blockDim.x + threadIdx.x;  // index along the non-filtered axis
	const long q = grid_row();
	const int n_other = g.axis == 0 ? g.F : g.T;
	if (other >= n_other || q >= g.n_out)
		return;
	const int a = g.first + (int)q;
	const int dim = g.axis == 0 ? g.T : g.F;
	auto get = [&](int t) -> float {
		int i = a + g.tap_off + t;
		if (g.wrap) i = wrap_idx(i, dim);
		return g.axis == 0 ? __ldg(src + (size_t)i * g.F + other) : __ldg(src + (size_t)other * g.F + i);
	};
	float m = median_generic(get, g.L);
	if (g.axis == 0)
		dst[(size_t)a * g.F + other] = m;
	else
		dst[(size_t)other * g.F + a] = m;
}

// ---- box filter: wrap-padded moving average (box.h:194-213) ----
// A sliding sum that NEVER subtracts: the SSE path feeds it 1 / |X|^2, which is +inf on silent bins (hps.cu:591-592),
// and "add the tap that enters, subtract the one that leaves" would turn inf - inf into NaN for the rest of the row.
//
// Frequency axis: a CTA stages one row segment (+ L - 1 taps of halo, wrap by index arithmetic) in shared memory and
// cuts it into blocks of L taps.  P[p] = sum of the block's taps up to p, S[p] = sum from p to the block's end (one
// thread per block, 2 L additions); a window of L taps covers the tail of one block and the head of the next, so
// out[q] = S[q] + P[q + L - 1] (or P alone when the window is a whole block): about three additions per output
// whatever L is, every tap read from HBM once.  An inf inside the window makes the sum inf, one outside cannot touch it.
constexpr int BOX_CHUNK = 2048;
constexpr int BOX_NT = 256;

__global__ void __launch_bounds__(BOX_NT) box_freq_kernel(const float* __restrict__ src, float* __restrict__ dst, AxisGeom g)
{
	extern __shared__ float box_smem[];
	const int L = g.L;
	const int n_in = BOX_CHUNK + L - 1;
	float* E = box_smem;                         // n_in
	float* P = E + ((n_in + 3) & ~3);            // n_in
	float* S = P + ((n_in + 3) & ~3);            // n_in
	const long r = grid_row();
	if (r >= g.T)
		return;
	const int q0 = blockIdx.x * BOX_CHUNK;
	const int nq = min(BOX_CHUNK, g.F - q0);
	const int n_used = nq + L - 1;
	const float* row = src + (size_t)r * g.F;
	for (int t = threadIdx.x; t < n_used; t += BOX_NT)
		E[t] = __ldg(row + wrap_idx(q0 + g.tap_off + t, g.F));
	__syncthreads();
	const int n_blocks = (n_used + L - 1) / L;
	for (int b = threadIdx.x; b < n_blocks; b += BOX_NT) {
		const int lo = b * L, hi = min(n_used, lo + L);
		float acc = 0.0f;
		for (int p = lo; p < hi; ++p) {
			acc += E[p];
			P[p] = acc;
		}
		acc = 0.0f;
		for (int p = hi - 1; p >= lo; --p) {
			acc += E[p];
			S[p] = acc;
		}
	}
	__syncthreads();
	const float inv = (float)L;
	for (int q = threadIdx.x; q < nq; q += BOX_NT) {
		const int o = q % L;
		const float sum = o == 0 ? P[q + L - 1] : S[q] + P[q + L - 1];
		dst[(size_t)r * g.F + q0 + q] = sum / inv;
	}
}

// Time axis: a thread owns one column and walks down a strip of rows with the last L taps in registers (L <= 15): one
// load per output, the window re-summed every step (no subtraction, see above).  Longer windows read their L taps.
template <int C>
__global__ void __launch_bounds__(256) box_time_kernel(const float* __restrict__ src, float* __restrict__ dst, AxisGeom g, int run_t)
{
	const int c = blockIdx.x * blockDim.x + threadIdx.x;
	const int r0 = blockIdx.y * run_t;
	if (c >= g.F)
		return;
	const int r1 = min(g.T, r0 + run_t);
	const int L = g.L;
	float w[C];
#pragma unroll
	for (int t = 0; t < C; ++t)
		w[t] = 0.0f;
	// taps of output r: rows r + tap_off + t, t < L; w[C - L .. C) holds them, oldest first
#pragma unroll
	for (int t = 0; t < C - 1; ++t) {
		const int tt = t - (C - L);   // tap index of slot t for the first output, minus the one loaded in the loop
		w[t + 1] = tt >= 0 && tt < L - 1 ? __ldg(src + (size_t)wrap_idx(r0 + g.tap_off + tt, g.T) * g.F + c) : 0.0f;
	}
	for (int r = r0; r < r1; ++r) {
#pragma unroll
		for (int t = 0; t < C - 1; ++t)
			w[t] = w[t + 1];
		w[C - 1] = __ldg(src + (size_t)wrap_idx(r + g.tap_off + L - 1, g.T) * g.F + c);
		float acc = 0.0f;
#pragma unroll
		for (int t = 0; t < C; ++t)
			if (t >= C - L) acc += w[t];
		dst[(size_t)r * g.F + c] = acc / (float)L;
	}
}

__global__ void __launch_bounds__(256) box_time_direct_kernel(const float* __restrict__ src, float* __restrict__ dst, AxisGeom g)
{
	const int c = blockIdx.x * blockDim.x + threadIdx.x;
	const long r = grid_row();
	if (c >= g.F || r >= g.T)
		return;
	float acc = 0.0f;
	for (int t = 0; t < g.L; ++t)
		acc += __ldg(src + (size_t)wrap_idx((int)r + g.tap_off + t, g.T) * g.F + c);
	dst[(size_t)r * g.F + c] = acc / (float)g.L;
}

int make_geom(AxisGeom& g, int T, int F, int filter_len, int dir, int copy_bord)
{
	if (T < 1 || F < 1 || filter_len < 1 || dir < 0 || dir > 2)
		return ZEN_ERR_ARG;
	// mfilt.h:80-87: checked before the length is made odd
	if (((dir == ZEN_TIME_CAUSAL || dir == ZEN_TIME_ANTICAUSAL) && filter_len > T) || (dir == ZEN_FREQUENCY && filter_len > F))
		return ZEN_ERR_GEOMETRY;
	const int L = odd_len(filter_len), mid = L / 2;
	g.T = T;
	g.F = F;
	g.L = L;
	g.axis = dir == ZEN_FREQUENCY ? 1 : 0;
	const int dim = g.axis ? F : T;
	if (copy_bord) {
		g.first = 0; g.n_out = dim; g.tap_off = -mid; g.wrap = 1;
	}
	else if (dir == ZEN_TIME_CAUSAL) {
		g.first = L; g.n_out = T - L; g.tap_off = -L; g.wrap = 0;
	}
	else if (dir == ZEN_TIME_ANTICAUSAL) {
		g.first = mid; g.n_out = T - L; g.tap_off = -mid; g.wrap = 0;
	}
	else {
		g.first = 0; g.n_out = F - L; g.tap_off = 0; g.wrap = 0;
	}
	return ZEN_OK;
}

template <int L, class OP>
void launch_pair(const float* src, float* dst, const AxisGeom& g, cudaStream_t s)
{
	if (g.axis == 1) {
		dim3 grid = rows_grid((g.n_out + PAIR_CHUNK - 1) / PAIR_CHUNK, g.T);
		pair_freq_kernel<L, OP><<<grid, PAIR_NT, 0, s>>>(src, dst, g);
	}
	else {
		// run length (even): long enough to amortise the L - 1 rows of halo, short enough to fill the GPU
		const int col_blocks = (g.F + 255) / 256;
		int run_t = 128;
		while (run_t > 8 && (long)col_blocks * ((g.n_out + run_t - 1) / run_t) < 1200)
			run_t >>= 1;
		dim3 grid = rows_grid(col_blocks, (g.n_out + run_t - 1) / run_t);
		pair_time_kernel<L, OP><<<grid, 256, 0, s>>>(src, dst, g, run_t);
	}
}

template <int C>
void launch_run(const float* src, float* dst, const AxisGeom& g, cudaStream_t s)
{
	if (g.axis == 1) {
		constexpr int CHUNK = SM_NT * RUN_F;
		dim3 grid = rows_grid((g.n_out + CHUNK - 1) / CHUNK, g.T);
		size_t smem = sizeof(float) * (size_t)(2 * CHUNK + g.L + 8);
		median_run_freq_kernel<C><<<grid, SM_NT, smem, s>>>(src, dst, g);
	}
	else {
		// run length: long enough to amortise the initial sort, short enough for >= ~4 CTAs per SM
		const int col_blocks = (g.F + 255) / 256;
		int run_t = 64;
		while (run_t > 8 && (long)col_blocks * ((g.n_out + run_t - 1) / run_t) < 600)
			run_t >>= 1;
		if ((g.n_out + run_t - 1) / run_t > 65535)
			run_t = (g.n_out + 65534) / 65535;
		dim3 grid(col_blocks, (g.n_out + run_t - 1) / run_t);
		median_run_time_kernel<C><<<grid, 256, 0, s>>>(src, dst, g, run_t);
	}
}

}  // namespace

extern "C" {

int zen_median_filter(int time, int freq, int filter_len, int direction, int copy_bord,
                      const float* d_src, float* d_dst, void* cuda_stream)
{
	if (!d_src || !d_dst)
		return ZEN_ERR_ARG;
	AxisGeom g;
	int rc = make_geom(g, time, freq, filter_len, direction, copy_bord);
	if (rc != ZEN_OK)
		return rc;
	if (g.n_out <= 0)
		return ZEN_OK;
	cudaStream_t s = (cudaStream_t)cuda_stream;
	if (g.L == 1) {
		dim3 grid = rows_grid((g.F + 255) / 256, g.T);
		copy_axis_kernel<<<grid, 256, 0, s>>>(d_src, d_dst, g);
	}
	else if (g.L <= 47) {
		switch (g.L) {
		case 3: launch_pair<3, MedianPairOp<3>>(d_src, d_dst, g, s); break;
		case 5: launch_pair<5, MedianPairOp<5>>(d_src, d_dst, g, s); break;
		case 7: launch_pair<7, MedianPairOp<7>>(d_src, d_dst, g, s); break;
		case 9: launch_pair<9, MedianPairOp<9>>(d_src, d_dst, g, s); break;
		case 11: launch_pair<11, MedianPairOp<11>>(d_src, d_dst, g, s); break;
		case 13: launch_pair<13, MedianPairOp<13>>(d_src, d_dst, g, s); break;
		case 15: launch_pair<15, MedianPairOp<15>>(d_src, d_dst, g, s); break;
		default:
			if (g.L <= 23) launch_run<24>(d_src, d_dst, g, s);
			else if (g.L <= 31) launch_run<32>(d_src, d_dst, g, s);
			else launch_run<48>(d_src, d_dst, g, s);
		}
	}
	else {
		const int K = g.axis == 1 ? sliding_K_for(g.L) : 0;
		if (K > 0) {
			dim3 grid = rows_grid((g.n_out + SL_CHUNK - 1) / SL_CHUNK, g.T);
			size_t smem = sizeof(unsigned) * (size_t)(2 * SL_CHUNK + g.L + 8);
			median_slide_freq_kernel<<<grid, SL_NT, smem, s>>>(d_src, d_dst, g, K);
		}
		else {
			const int n_other = g.axis == 0 ? g.F : g.T;
			dim3 grid = rows_grid((n_other + 255) / 256, g.n_out);
			median_generic_kernel<<<grid, 256, 0, s>>>(d_src, d_dst, g);
		}
	}
	ZEN_CUDA_CHECK(cudaGetLastError());
	return ZEN_OK;
}

int zen_box_filter(int time, int freq, int filter_len, int direction, const float* d_src, float* d_dst, void* cuda_stream)
{
	if (!d_src || !d_dst)
		return ZEN_ERR_ARG;
	AxisGeom g;
	int rc = make_geom(g, time, freq, filter_len, direction, 1);  // BoxFilterGPU always wrap-pads
	if (rc != ZEN_OK)
		return rc;
	cudaStream_t s = (cudaStream_t)cuda_stream;
	if (g.L >= 3 && g.L <= 15) {
		// short windows: the output-pair kernels of the median filter with a sum in place of the selection
		switch (g.L) {
		case 3: launch_pair<3, BoxPairOp<3>>(d_src, d_dst, g, s); break;
		case 5: launch_pair<5, BoxPairOp<5>>(d_src, d_dst, g, s); break;
		case 7: launch_pair<7, BoxPairOp<7>>(d_src, d_dst, g, s); break;
		case 9: launch_pair<9, BoxPairOp<9>>(d_src, d_dst, g, s); break;
		case 11: launch_pair<11, BoxPairOp<11>>(d_src, d_dst, g, s); break;
		case 13: launch_pair<13, BoxPairOp<13>>(d_src, d_dst, g, s); break;
		default: launch_pair<15, BoxPairOp<15>>(d_src, d_dst, g, s); break;
		}
	}
	else if (g.axis == 1) {
		const size_t smem = sizeof(float) * 3 * (size_t)((BOX_CHUNK + g.L - 1 + 3) & ~3);
		if (smem > 200 * 1024)
			return ZEN_ERR_UNSUPPORTED;
		ZEN_CUDA_CHECK(cudaFuncSetAttribute(box_freq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
		dim3 grid = rows_grid((freq + BOX_CHUNK - 1) / BOX_CHUNK, time);
		box_freq_kernel<<<grid, BOX_NT, smem, s>>>(d_src, d_dst, g);
	}
	else if (g.L <= 15) {
		const int col_blocks = (freq + 255) / 256;
		int run_t = 64;
		while (run_t > 8 && (long)col_blocks * ((time + run_t - 1) / run_t) < 600)
			run_t >>= 1;
		if ((time + run_t - 1) / run_t > 65535)
			run_t = (time + 65534) / 65535;
		dim3 grid(col_blocks, (time + run_t - 1) / run_t);
		if (g.L <= 7) box_time_kernel<8><<<grid, 256, 0, s>>>(d_src, d_dst, g, run_t);
		else box_time_kernel<16><<<grid, 256, 0, s>>>(d_src, d_dst, g, run_t);
	}
	else {
		dim3 grid = rows_grid((freq + 255) / 256, time);
		box_time_direct_kernel<<<grid, 256, 0, s>>>(d_src, d_dst, g);
	}
	ZEN_CUDA_CHECK(cudaGetLastError());
	return ZEN_OK;
}

}  // extern "C"
