// Kernels of the fused HPR step and their launchers, instantiated once per FFT
// size in hpr_inst_<NFFT>.cu so the sizes compile in parallel.
#pragma once
#include <algorithm>

#include "hpr_core.cuh"

namespace zen_b200 {

template <int NFFT>
constexpr int nt_for()
{
	return (NFFT / 16) < 64 ? 64 : ((NFFT / 16) > 512 ? 512 : (NFFT / 16));
}
// resident CTAs per SM the register allocation is tuned for
#ifndef ZEN_TILE_THREADS_PER_SM
#define ZEN_TILE_THREADS_PER_SM 1024
#endif
template <int NT, int NFFT = 0>
constexpr int min_blocks_for()
{
	// 512-thread CTAs: two fit at nfft 8192 (101 KB of shared memory each), one at 16384 (200 KB)
	return NT >= 512 ? (NFFT == 8192 ? 2 : 1) : ZEN_TILE_THREADS_PER_SM / NT;
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
	float* scratch;          // resident_ctas * scratch_per_cta floats
	size_t scratch_per_cta;
	int* work_counter;       // device int, zeroed by the launcher
	int resident_ctas;       // grid size (<= SMs * occupancy)
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

// Control block of the persistent real-time kernel, in mapped pinned host memory.
// Host -> device: fill the arguments, then publish seq_in (the doorbell).
// Device -> host: seq_out = the sequence number just completed.
struct RtCtrl {
	volatile unsigned seq_in;
	volatile unsigned op;          // RT_OP_*
	const float* in;               // device-visible pointer to the incoming hop
	float* out[3];                 // device-visible destinations of the emitted hop (H, P, R) or null
	int which;                     // RT_OP_COPY: which output
	unsigned pad0[16];
	volatile unsigned seq_out;     // own cache line
	volatile unsigned alive;       // 1 while the kernel is resident
	volatile unsigned exit_reason; // 1 stop requested, 2 idle time-out
	unsigned pad1[13];
	unsigned long long stamps[16]; // globaltimer at the phase boundaries of the last hop (diagnostics)
};
enum { RT_OP_PROCESS = 1, RT_OP_COPY = 2, RT_OP_STOP = 3, RT_OP_NEW_ARGS = 0x100 };

struct RtArgs {
	HprDev dev;
	RtCtrl* ctrl;          // device alias of the control block
	float* mag_ring;       // global state, loaded at start and written back at exit
	float* input;          // nwin
	float* ola[3];         // nwin each
	int* iter;             // frame counter, read at start, written back at exit
	unsigned seq0;         // last sequence number already served
	unsigned long long idle_ns;
	int state_in_smem;     // ring + tails resident in shared memory
	cudaStream_t stream;
};

template <int NFFT>
constexpr int nt_rt_for()
{
	// measured on a B200 (tools/rt_phases.py): 512 threads are the sweet spot at nfft 4096 (1024 threads spill in
	// the FFT stages); the large transforms want 1024, the small ones nfft/4
	return NFFT == 4096 ? 512 : ((NFFT / 4) < 128 ? 128 : ((NFFT / 4) > 1024 ? 1024 : (NFFT / 4)));
}
template <int NFFT>
constexpr int rt_u_for()
{
	return NFFT == 4096 ? 5 : 3;
}

template <int NFFT> int launch_tile_impl(const TileArgs& a);
template <int NFFT> int tile_resident_ctas(const HprDev& d);
template <int NFFT> int launch_hop_impl(const HopArgs& a);
template <int NFFT> int launch_rt_impl(const RtArgs& a);
template <int NFFT> size_t rt_smem_bytes(const HprDev& d, int state_in_smem);

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
__global__ void __launch_bounds__(NT, min_blocks_for<NT, NFFT>()) hpr_tile_kernel(const __grid_constant__ HprDev P,
                                                                            const float* __restrict__ in, long in_stride,
                                                                            float* out_h, float* out_p, float* out_r, long out_stride,
                                                                            long n_hops, int tile_hops, int n_tiles, int total_items,
                                                                            int* work_counter, float* scratch, size_t scratch_per_cta)
{
	constexpr int M = NFFT / 2, HOP = M / 2;
	extern __shared__ __align__(16) unsigned char smem_raw[];
	HprSmem<NFFT> sm;
	sm.carve(smem_raw, P.Lp);
	__shared__ int s_item;

	// scratch belongs to the RESIDENT CTA, not to the work item: a few tens of MB that stay in L2
	float* sc = scratch + (size_t)blockIdx.x * scratch_per_cta;
	HprState st;
	st.mag_ring = sc;
	sc += (size_t)P.W * (M + 1) + ((P.W * (M + 1)) & 1);
	st.xdepth = P.lag > 1 ? P.lag : 0;
	st.x_ring = reinterpret_cast<float2*>(sc);
	sc += 2 * (size_t)st.xdepth * (M + 1);
	for (int o = 0; o < 3; ++o)
		st.tail[o] = sc + (size_t)o * HOP;

	// work items (stream, tile) are handed out dynamically, so no CTA idles while others finish a long tail
	for (;;) {
		__syncthreads();
		if (threadIdx.x == 0)
			s_item = atomicAdd(work_counter, 1);
		__syncthreads();
		const int item = s_item;
		if (item >= total_items)
			break;
		const int stream = item / n_tiles, tile = item - stream * n_tiles;
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
			hpr_iteration<NFFT, NT>(P, sm, st, (int)i, prev, cur, i >= i_full, i == i_full, em, nullptr, nullptr,
			                        i + 1 < e1 ? cur + HOP : nullptr);
		}
	}
}

// One hop of one persistent stream (HPR<GPU>::process_next_hop): state lives in
// the zen_hpr object between launches.  ola[o] is the reference's *_out vector:
// [0:hop] the emitted hop, [hop:nwin] the overlap-add tail.
template <int NFFT, int NT>
__global__ void __launch_bounds__(NT, 1) hpr_hop_kernel(const __grid_constant__ HprDev P, HprState st, long i,
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
	hpr_iteration<NFFT, NT, rt_u_for<NFFT>()>(P, sm, st, (int)i, input, input + HOP, true, false, em);
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



// Persistent real-time kernel: one resident CTA serves one stream hop after hop.
// The host writes the hop into mapped memory and rings a doorbell; the CTA polls
// it, runs the fused step with ring / overlap-add tails / previous hop held in
// SHARED memory, writes the outputs straight to mapped host memory and publishes
// the completion flag.  No launch and no stream synchronisation per hop (those
// alone cost ~9 us on this box).  The kernel leaves on RT_OP_STOP or after
// idle_ns without a doorbell, writing its state back to global memory so the
// per-launch kernels (or a later session) continue seamlessly.
__device__ __forceinline__ unsigned long long rt_globaltimer()
{
	unsigned long long t;
	asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
	return t;
}

template <int NFFT, int NT>
__global__ void __launch_bounds__(NT, 1) hpr_rt_kernel(const __grid_constant__ HprDev P, RtCtrl* ctrl, float* g_ring, float* g_input,
                                                       float* g_ola_h, float* g_ola_p, float* g_ola_r, int* g_iter, unsigned seq0,
                                                       unsigned long long idle_ns, int state_in_smem)
{
	constexpr int M = NFFT / 2, HOP = M / 2;
	extern __shared__ __align__(16) unsigned char smem_raw[];
	HprSmem<NFFT> sm;
	sm.carve(smem_raw, P.Lp);
	float* extra = reinterpret_cast<float*>(smem_raw + ((HprSmem<NFFT>::bytes(P.Lp) + 15) & ~(size_t)15));
	// double buffer: previous hop / incoming hop (in global memory, i.e. the `input` vector itself, when smem is short)
	float* hopbuf[2] = {state_in_smem ? extra : g_input, state_in_smem ? extra + HOP : g_input + HOP};
	if (state_in_smem) extra += 2 * HOP;
	const int tid = threadIdx.x;
	const int ring_n = P.W * (M + 1);
	float* g_ola[3] = {g_ola_h, g_ola_p, g_ola_r};
	HprState st;
	st.x_ring = nullptr;
	st.xdepth = 0;
	if (state_in_smem) {
		st.mag_ring = extra;
		extra += (ring_n + 3) & ~3;
		for (int o = 0; o < 3; ++o)
			st.tail[o] = extra + o * HOP;
		for (int n = tid; n < ring_n; n += NT)
			st.mag_ring[n] = g_ring[n];
		for (int o = 0; o < 3; ++o)
			for (int n = tid; n < HOP; n += NT)
				st.tail[o][n] = g_ola[o][HOP + n];
	}
	else {
		st.mag_ring = g_ring;
		for (int o = 0; o < 3; ++o)
			st.tail[o] = g_ola[o] + HOP;
	}
	if (state_in_smem) {
		for (int n = tid; n < HOP; n += NT) {
			hopbuf[0][n] = g_input[n];          // older hop (only kept so `input` can be written back)
			hopbuf[1][n] = g_input[HOP + n];    // the previous hop
		}
	}
	int prev_idx = 1;
	int iter = *g_iter;
	unsigned seq = seq0;
	__shared__ unsigned s_op, s_seq;
	__shared__ unsigned long long s_stamps[16];
	__shared__ const float* s_in;
	__shared__ float* s_out[3];
	__shared__ int s_which;
	__syncthreads();

	for (;;) {
		if (tid == 0) {
			const unsigned long long t0 = rt_globaltimer();
			unsigned now = seq, opw = 0;
			unsigned spins = 0;
			for (;;) {
				// doorbell and op code share one 8-byte word: one PCIe read per poll
				asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(now), "=r"(opw) : "l"(&ctrl->seq_in) : "memory");
				if (now != seq)
					break;
				if ((++spins & 63u) == 0 && rt_globaltimer() - t0 > idle_ns)
					break;
			}
			if (now == seq) {
				s_op = 0xffffffffu;  // idle time-out
			}
			else {
				if (opw & RT_OP_NEW_ARGS) {
					// pointers changed since the last hop: fetch them (one more round trip), else reuse the cached ones
					__threadfence_system();
					s_in = ctrl->in;
					s_out[0] = ctrl->out[0];
					s_out[1] = ctrl->out[1];
					s_out[2] = ctrl->out[2];
					s_which = ctrl->which;
				}
				s_op = opw & 0xffu;
			}
			s_seq = now;
		}
		__syncthreads();
		const unsigned op = s_op;
		if (op == 0xffffffffu || op == RT_OP_STOP) {
			if (tid == 0) ctrl->exit_reason = (op == RT_OP_STOP) ? 1u : 2u;
			if (op == RT_OP_STOP) seq = s_seq;
			break;
		}
		if (op == RT_OP_PROCESS) {
			HprEmit em;
			for (int o = 0; o < 3; ++o) {
				const bool on = (P.out_flags & (1 << o)) != 0;
				em.a[o] = on ? g_ola[o] : nullptr;   // keep the public *_out vectors current (hps.h:195-197)
				em.b[o] = on ? s_out[o] : nullptr;
			}
			float* stash = hopbuf[prev_idx ^ 1];
			if (tid == 0) {
				s_stamps[9] = rt_globaltimer();
				s_stamps[10] = (unsigned long long)clock64();
			}
			hpr_iteration<NFFT, NT, rt_u_for<NFFT>()>(P, sm, st, iter, hopbuf[prev_idx], s_in, true, false, em, stash, s_stamps);
			if ((P.out_flags & ZEN_OUTPUT_RESIDUAL) && (P.soft || P.sse) && s_out[2])
				for (int n = tid; n < HOP; n += NT)
					s_out[2][n] = 0.0f;
			prev_idx ^= 1;
			++iter;
			if (tid == 0) s_stamps[11] = (unsigned long long)clock64();
		}
		else if (op == RT_OP_COPY) {
			const float* src = g_ola[s_which];
			float* dst = s_out[0];
			for (int n = tid; n < HOP; n += NT)
				dst[n] = src[n];
		}
		__syncthreads();
		seq = s_seq;
		if (tid == 0) {
			__threadfence_system();
			ctrl->seq_out = seq;
		}
		if (tid < 16 && op == RT_OP_PROCESS)
			ctrl->stamps[tid] = s_stamps[tid];  // diagnostics, after the completion flag
	}

	// write the state back
	if (state_in_smem) {
		for (int n = tid; n < ring_n; n += NT)
			g_ring[n] = st.mag_ring[n];
		for (int o = 0; o < 3; ++o)
			for (int n = tid; n < HOP; n += NT)
				g_ola[o][HOP + n] = st.tail[o][n];
	}
	if (state_in_smem || prev_idx == 0) {
		// `input` must read [older hop | newest hop]; when it served as the double buffer itself, swap the halves if needed
		for (int n = tid; n < HOP; n += NT) {
			float older = hopbuf[prev_idx ^ 1][n], newest = hopbuf[prev_idx][n];
			g_input[n] = older;
			g_input[HOP + n] = newest;
		}
	}
	if (tid == 0) *g_iter = iter;
	__syncthreads();
	if (tid == 0) {
		__threadfence_system();
		ctrl->seq_out = seq;
		ctrl->alive = 0u;
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
	ZEN_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
	const int n_tiles = (int)((a.n_hops + a.tile_hops - 1) / a.tile_hops);
	const long total = (long)n_tiles * a.n_streams;
	if (total > 0x7fffffffL)
		return ZEN_ERR_UNSUPPORTED;
	ZEN_CUDA_CHECK(cudaMemsetAsync(a.work_counter, 0, sizeof(int), a.stream));
	const int grid = (int)std::min<long>(total, a.resident_ctas);
	kern<<<grid, NT, smem, a.stream>>>(a.dev, a.in, a.in_stride, a.out_h, a.out_p, a.out_r, a.out_stride, a.n_hops, a.tile_hops,
	                                   n_tiles, (int)total, a.work_counter, a.scratch, a.scratch_per_cta);
	ZEN_CUDA_CHECK(cudaGetLastError());
	return ZEN_OK;
}

// resident CTAs per SM of the tile kernel (occupancy query), times the SM count
template <int NFFT>
int tile_resident_ctas(const HprDev& d)
{
	constexpr int NT = nt_for<NFFT>();
	auto kern = hpr_tile_kernel<NFFT, NT>;
	size_t smem = HprSmem<NFFT>::bytes(d.Lp);
	if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess
	    || cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared) != cudaSuccess)
		return 0;
	int per_sm = 0, dev = 0, sms = 0;
	if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NT, smem) != cudaSuccess || cudaGetDevice(&dev) != cudaSuccess
	    || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
		return 0;
	return per_sm * sms;
}

template <int NFFT>
int launch_hop_impl(const HopArgs& a)
{
	constexpr int NT = nt_rt_for<NFFT>();  // one CTA per hop: latency, not throughput, decides the thread count
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

template <int NFFT>
size_t rt_smem_bytes(const HprDev& d, int state_in_smem)
{
	constexpr int M = NFFT / 2, HOP = M / 2;
	size_t b = (HprSmem<NFFT>::bytes(d.Lp) + 15) & ~(size_t)15;
	if (state_in_smem)
		b += sizeof(float) * (2 * (size_t)HOP + (((size_t)d.W * (M + 1) + 3) & ~(size_t)3) + 3 * (size_t)HOP);
	return b;
}

template <int NFFT>
int launch_rt_impl(const RtArgs& a)
{
	constexpr int NT = nt_rt_for<NFFT>();
	auto kern = hpr_rt_kernel<NFFT, NT>;
	size_t smem = rt_smem_bytes<NFFT>(a.dev, a.state_in_smem);
	ZEN_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	kern<<<1, NT, smem, a.stream>>>(a.dev, a.ctrl, a.mag_ring, a.input, a.ola[0], a.ola[1], a.ola[2], a.iter, a.seq0, a.idle_ns,
	                                a.state_in_smem);
	ZEN_CUDA_CHECK(cudaGetLastError());
	return ZEN_OK;
}

template int launch_tile_impl<ZEN_HPR_INSTANTIATE>(const TileArgs&);
template int tile_resident_ctas<ZEN_HPR_INSTANTIATE>(const HprDev&);
template int launch_hop_impl<ZEN_HPR_INSTANTIATE>(const HopArgs&);
template int launch_rt_impl<ZEN_HPR_INSTANTIATE>(const RtArgs&);
template size_t rt_smem_bytes<ZEN_HPR_INSTANTIATE>(const HprDev&, int);

}  // namespace zen_b200
#endif
