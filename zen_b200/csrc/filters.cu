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

__device__ __forceinline__ int wrap_idx(int i, int dim)
{
	i %= dim;
	return i < 0 ? i + dim : i;
}

// ---- short windows (L <= 15): sorting network in registers ----
// Frequency axis: one output per thread, taps are L consecutive floats that
// neighbouring threads share through L1.  Time axis: RT consecutive rows per
// thread so L + RT - 1 loads serve RT outputs.
template <int L>
__global__ void __launch_bounds__(256) median_small_freq_kernel(const float* __restrict__ src, float* __restrict__ dst, AxisGeom g)
{
	const int r = blockIdx.y;
	const int q = blockIdx.x * blockDim.x + threadIdx.x;
	if (q >= g.n_out)
		return;
	const int c = g.first + q;
	const float* row = src + (size_t)r * g.F;
	float v[L];
#pragma unroll
	for (int t = 0; t < L; ++t) {
		int i = c + g.tap_off + t;
		if (g.wrap) i = wrap_idx(i, g.F);
		v[t] = __ldg(row + i);
	}
	dst[(size_t)r * g.F + c] = median_regs<L>(v);
}

template <int L, int RT>
__global__ void __launch_bounds__(256) median_small_time_kernel(const float* __restrict__ src, float* __restrict__ dst, AxisGeom g)
{
	const int c = blockIdx.x * blockDim.x + threadIdx.x;
	const int q0 = blockIdx.y * RT;
	if (c >= g.F)
		return;
	float w[L + RT - 1];
#pragma unroll
	for (int t = 0; t < L + RT - 1; ++t) {
		int i = g.first + q0 + g.tap_off + t;
		if (g.wrap) i = wrap_idx(i, g.T);
		// rows past the last output of this block are loaded only if they exist
		bool ok = g.wrap || (i >= 0 && i < g.T);
		w[t] = ok ? __ldg(src + (size_t)i * g.F + c) : 0.0f;
	}
#pragma unroll
	for (int u = 0; u < RT; ++u) {
		if (q0 + u < g.n_out) {
			float v[L];
#pragma unroll
			for (int t = 0; t < L; ++t)
				v[t] = w[u + t];
			dst[(size_t)(g.first + q0 + u) * g.F + c] = median_regs<L>(v);
		}
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
	const int r = blockIdx.y;
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
	const int q = blockIdx.y;
	const int n_other = g.axis == 0 ? g.F : g.T;
	if (other >= n_other)
		return;
	const int a = g.first + q;
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
__global__ void __launch_bounds__(256) box_kernel(const float* __restrict__ src, float* __restrict__ dst, AxisGeom g)
{
	const int c = blockIdx.x * blockDim.x + threadIdx.x;
	const int r = blockIdx.y;
	if (c >= g.F)
		return;
	float acc = 0.0f;
	if (g.axis == 1) {
		const float* row = src + (size_t)r * g.F;
		for (int t = 0; t < g.L; ++t)
			acc += __ldg(row + wrap_idx(c + g.tap_off + t, g.F));
	}
	else {
		for (int t = 0; t < g.L; ++t)
			acc += __ldg(src + (size_t)wrap_idx(r + g.tap_off + t, g.T) * g.F + c);
	}
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

template <int L>
void launch_small(const float* src, float* dst, const AxisGeom& g, cudaStream_t s)
{
	if (g.axis == 1) {
		dim3 grid((g.n_out + 255) / 256, g.T);
		median_small_freq_kernel<L><<<grid, 256, 0, s>>>(src, dst, g);
	}
	else {
		constexpr int RT = 4;
		dim3 grid((g.F + 255) / 256, (g.n_out + RT - 1) / RT);
		median_small_time_kernel<L, RT><<<grid, 256, 0, s>>>(src, dst, g);
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
	switch (g.L) {
	case 1: launch_small<1>(d_src, d_dst, g, s); break;
	case 3: launch_small<3>(d_src, d_dst, g, s); break;
	case 5: launch_small<5>(d_src, d_dst, g, s); break;
	case 7: launch_small<7>(d_src, d_dst, g, s); break;
	case 9: launch_small<9>(d_src, d_dst, g, s); break;
	case 11: launch_small<11>(d_src, d_dst, g, s); break;
	case 13: launch_small<13>(d_src, d_dst, g, s); break;
	case 15: launch_small<15>(d_src, d_dst, g, s); break;
	default: {
		const int K = g.axis == 1 ? sliding_K_for(g.L) : 0;
		if (K > 0) {
			dim3 grid((g.n_out + SL_CHUNK - 1) / SL_CHUNK, g.T);
			size_t smem = sizeof(unsigned) * (size_t)(2 * SL_CHUNK + g.L + 8);
			median_slide_freq_kernel<<<grid, SL_NT, smem, s>>>(d_src, d_dst, g, K);
		}
		else {
			const int n_other = g.axis == 0 ? g.F : g.T;
			dim3 grid((n_other + 255) / 256, g.n_out);
			median_generic_kernel<<<grid, 256, 0, s>>>(d_src, d_dst, g);
		}
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
	dim3 grid((freq + 255) / 256, time);
	box_kernel<<<grid, 256, 0, (cudaStream_t)cuda_stream>>>(d_src, d_dst, g);
	ZEN_CUDA_CHECK(cudaGetLastError());
	return ZEN_OK;
}

}  // extern "C"
