// Kernels of the fused HPR step and their launchers, instantiated once per FFT
// size in hpr_inst_<NFFT>.cu so the sizes compile in parallel.
#pragma once
#include "hpr_core.cuh"

namespace zen_b200 {

template <int NFFT>
constexpr int nt_for()
{
	return (NFFT / 16) < 64 ? 64 : ((NFFT / 16) > 512 ? 512 : (NFFT / 16));
}
// resident CTAs per SM the register allocation is tuned for
template <int NT>
constexpr int min_blocks_for()
{
	return NT >= 512 ? 1 : 768 / NT;
}

struct TileArgs {
	HprDev dev;
	const float* in;
	long in_stride;
	float *out_h, *out_p, *out_r;
	long out_stride;
	int n_streams;
	long n_hops;
	int tile_hops;
	float* scratch;
	size_t scratch_per_cta;
	cudaStream_t stream;
};

struct HopArgs {
	HprDev dev;
	HprState st;
	long iter;
	float* input;
	const float* in_hop;
	float* ola[3];
	float* ext[3];
	cudaStream_t stream;
};

template <int NFFT> int launch_tile_impl(const TileArgs& a);
template <int NFFT> int launch_hop_impl(const HopArgs& a);

}  // namespace zen_b200

#ifdef ZEN_HPR_INSTANTIATE
using namespace zen_b200;

// ----------------------------------------------------------------- kernels ---

// One CTA = one stream x one tile of consecutive hops.  The CTA walks its hops
// in order; the W-1 hops before the tile are analysed only (ring fill) and the
// hop just before the tile is synthesised without being emitted (its second
// half is the first overlap-add tail of the tile).  Frames are independent
// given that halo, so tiles need no communication.
template <int NFFT, int NT>
__global__ void __launch_bounds__(NT, min_blocks_for<NT>()) hpr_tile_kernel(const __grid_constant__ HprDev P,
                                                      const float* __restrict__ in, long in_stride,
                                                      float* out_h, float* out_p, float* out_r, long out_stride,
                                                      long n_hops, int tile_hops,
                                                      float* scratch, size_t scratch_per_cta)
{
	constexpr int M = NFFT / 2, HOP = M / 2;
	extern __shared__ __align__(16) unsigned char smem_raw[];
	HprSmem<NFFT> sm;
	sm.carve(smem_raw, P.Lp);

	const int tile = blockIdx.x, stream = blockIdx.y;
	const size_t cta = (size_t)stream * gridDim.x + tile;
	float* sc = scratch + cta * scratch_per_cta;
	HprState st;
	st.mag_ring = sc;
	sc += (size_t)P.W * (M + 1) + ((P.W * (M + 1)) & 1);
	st.xdepth = P.lag > 1 ? P.lag : 0;
	st.x_ring = reinterpret_cast<float2*>(sc);
	sc += 2 * (size_t)st.xdepth * (M + 1);
	for (int o = 0; o < 3; ++o)
		st.tail[o] = sc + (size_t)o * HOP;

	const float* sin = in + (size_t)stream * in_stride;
	const long e0 = (long)tile * tile_hops;
	const long e1 = min(n_hops, e0 + (long)tile_hops);
	const long i_begin = max(0L, e0 - P.W);
	const long i_full = max(0L, e0 - 1);
	for (long i = i_begin; i < e1; ++i) {
		const float* cur = sin + (size_t)i * HOP;
		const float* prev = i > 0 ? cur - HOP : nullptr;
		HprEmit em;
		const bool emit = i >= e0;
		em.a[0] = (emit && out_h) ? out_h + (size_t)stream * out_stride + (size_t)i * HOP : nullptr;
		em.a[1] = (emit && out_p) ? out_p + (size_t)stream * out_stride + (size_t)i * HOP : nullptr;
		em.a[2] = (emit && out_r) ? out_r + (size_t)stream * out_stride + (size_t)i * HOP : nullptr;
		em.b[0] = em.b[1] = em.b[2] = nullptr;
		hpr_iteration<NFFT, NT>(P, sm, st, (int)i, prev, cur, i >= i_full, i == i_full, em);
	}
}

// One hop of one persistent stream (HPR<GPU>::process_next_hop): state lives in
// the zen_hpr object between launches.  ola[o] is the reference's *_out vector:
// [0:hop] the emitted hop, [hop:nwin] the overlap-add tail.
template <int NFFT, int NT>
__global__ void __launch_bounds__(NT, min_blocks_for<NT>()) hpr_hop_kernel(const __grid_constant__ HprDev P, HprState st, long i,
                                                     float* input /* nwin: previous hop | current hop */,
                                                     const float* __restrict__ in_hop,
                                                     float* ola_h, float* ola_p, float* ola_r,
                                                     float* ext_h, float* ext_p, float* ext_r)
{
	constexpr int M = NFFT / 2, HOP = M / 2;
	extern __shared__ __align__(16) unsigned char smem_raw[];
	HprSmem<NFFT> sm;
	sm.carve(smem_raw, P.Lp);
	// input = input[hop:] ++ in_hop   (hps.cu:452-453)
	for (int n = threadIdx.x; n < HOP; n += NT) {
		input[n] = input[HOP + n];
	}
	__syncthreads();
	for (int n = threadIdx.x; n < HOP; n += NT)
		input[HOP + n] = in_hop[n];
	__syncthreads();
	float* ola[3] = {ola_h, ola_p, ola_r};
	float* ext[3] = {ext_h, ext_p, ext_r};
	HprEmit em;
	for (int o = 0; o < 3; ++o) {
		st.tail[o] = ola[o] + HOP;
		em.a[o] = (P.out_flags & (1 << o)) ? ola[o] : nullptr;
		em.b[o] = (P.out_flags & (1 << o)) ? ext[o] : nullptr;
	}
	hpr_iteration<NFFT, NT>(P, sm, st, (int)i, input, input + HOP, true, false, em);
	// outputs that the masks never reach still advance like the reference's
	// rotate-and-zero (hps.cu:435-449): residual with soft mask / SSE
	if ((P.out_flags & ZEN_OUTPUT_RESIDUAL) && (P.soft || P.sse)) {
		for (int n = threadIdx.x; n < HOP; n += NT) {
			float t = ola_r[HOP + n];
			ola_r[n] = t;
			ola_r[HOP + n] = 0.0f;
			if (ext_r) ext_r[n] = t;
		}
	}
}


namespace zen_b200 {

template <int NFFT>
int launch_tile_impl(const TileArgs& a)
{
	constexpr int NT = nt_for<NFFT>();
	auto kern = hpr_tile_kernel<NFFT, NT>;
	size_t smem = HprSmem<NFFT>::bytes(a.dev.Lp);
	ZEN_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	int n_tiles = (int)((a.n_hops + a.tile_hops - 1) / a.tile_hops);
	dim3 grid(n_tiles, a.n_streams);
	kern<<<grid, NT, smem, a.stream>>>(a.dev, a.in, a.in_stride, a.out_h, a.out_p, a.out_r, a.out_stride, a.n_hops, a.tile_hops,
	                                   a.scratch, a.scratch_per_cta);
	ZEN_CUDA_CHECK(cudaGetLastError());
	return ZEN_OK;
}

template <int NFFT>
int launch_hop_impl(const HopArgs& a)
{
	constexpr int NT = nt_for<NFFT>();
	auto kern = hpr_hop_kernel<NFFT, NT>;
	size_t smem = HprSmem<NFFT>::bytes(a.dev.Lp);
	static thread_local size_t configured = 0;
	if (configured < smem) {
		ZEN_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
		configured = smem;
	}
	kern<<<1, NT, smem, a.stream>>>(a.dev, a.st, a.iter, a.input, a.in_hop, a.ola[0], a.ola[1], a.ola[2], a.ext[0], a.ext[1], a.ext[2]);
	ZEN_CUDA_CHECK(cudaGetLastError());
	return ZEN_OK;
}

template int launch_tile_impl<ZEN_HPR_INSTANTIATE>(const TileArgs&);
template int launch_hop_impl<ZEN_HPR_INSTANTIATE>(const HopArgs&);

}  // namespace zen_b200
#endif
