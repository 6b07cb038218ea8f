// Kernels of the fused HPR step and their launchers, instantiated once per FFT
// size in hpr_inst_<NFFT>.cu so the sizes compile in parallel.
#pragma once
#include <cstdlib>
#include <algorithm>

#include <cooperative_groups.h>

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

// The fast kernel keeps 37 KB of shared memory per CTA at nfft 4096 (the magnitude row borrows an FFT buffer), and
// proportionally less for the shorter transforms, so a fifth "quarter SM" of threads fits if the registers allow it:
// ZEN_FAST_THREADS_PER_SM 1280 = 48 registers per thread (1024 = 64; 1536 = 40 spills and is slower, measured at nfft 4096).
#ifndef ZEN_FAST_THREADS_PER_SM
#define ZEN_FAST_THREADS_PER_SM 1280
#endif
template <int NT, int NFFT>
constexpr int min_blocks_fast()
{
	return NT <= 256 ? ZEN_FAST_THREADS_PER_SM / NT : min_blocks_for<NT, NFFT>();
}

struct TileArgs {
	HprDev dev;
	const float* in;
	long in_stride;
	float *out_h, *out_p, *out_r;
	long out_stride;
	int n_streams;
	long n_hops;
	int tile_hops;           // tile length (of the closing tiles when the first ones are longer: big_hops / n_big below)
	float* scratch;          // resident_ctas * scratch_per_cta floats
	size_t scratch_per_cta;
	int* work_counter;       // device int, zeroed by the launcher
	int resident_ctas;       // grid size (<= SMs * occupancy)
	cudaStream_t stream;
	unsigned* peaks[3] = {nullptr, nullptr, nullptr};  // optional per-stream max |out| (fast path only), zeroed by the caller
	int force_general = 0;   // debugging / A-B: run hpr_iteration even where the fast path applies
	int big_hops = 0;        // length of the first n_big tiles of every stream (n_big == 0: all tiles are tile_hops long)
	int n_big = 0;
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
//
// Host -> device.  A request is published by TAGGING the staging buffer `stage_in`: ceil(hop/3) groups of
// 16 bytes {x[3g], x[3g+1], x[3g+2], tag ^ zen_group_hash(x)}, tag = (sequence number << 8) | op bits.  The CTA's threads poll
// their own groups, so with RT_F_PUSH_IN the doorbell and the samples arrive in the SAME PCIe round trip
// (a doorbell word followed by a read of the hop costs two).  A group is written with one aligned 16-byte
// store and read with one 16-byte load, and it validates itself: its tag word carries a hash of its samples
// (zen_group_key), so a group caught half-written is never taken for a new one.  Arguments that changed since
// the previous request (RT_F_NEW_ARGS) are written to this block BEFORE the tags.
// Device -> host.  With RT_F_TAG_OUT every emitted hop is written as tagged groups to stage_out[o] and the
// host unpacks it once the last group is in and every group checks out; seq_out (a release store) completes
// everything else.
struct RtCtrl {
	volatile unsigned seq_in;      // unused (the tags are the doorbell)
	volatile unsigned op;
	const float* in;               // device-visible pointer to the incoming hop (when it is not pushed)
	float* out[3];                 // device-visible destinations of the emitted hop (H, P, R) or null (when not tagged)
	int which;                     // RT_OP_COPY: which output
	unsigned pad0[16];
	volatile unsigned seq_out;     // own cache line
	volatile unsigned alive;       // 1 while the kernel is resident
	volatile unsigned exit_reason; // 1 stop requested, 2 idle time-out
	unsigned pad1[13];
	unsigned long long stamps[16]; // RT_F_STAMPS: SM cycle counter at the phase boundaries of the last hop, [9]/[12] globaltimer
	unsigned long long stamps_r1[16]; // the same as seen by CTA 1 of the cluster ([9]: globaltimer when it saw the command)
};
enum { RT_OP_PROCESS = 1, RT_OP_COPY = 2, RT_OP_STOP = 3, RT_OP_EXIT = 4 /* device-internal: idle time-out */, RT_OP_MASK = 0x0f, RT_F_NEW_ARGS = 0x10, RT_F_PUSH_IN = 0x20, RT_F_TAG_OUT = 0x40, RT_F_STAMPS = 0x80 };

struct RtArgs {
	HprDev dev;
	RtCtrl* ctrl;          // device alias of the control block
	const uint4* stage_in; // device alias of the tagged request / input staging buffer
	uint4* stage_out[3];   // device aliases of the tagged output staging buffers
	float* mag_ring;       // global state, loaded at start and written back at exit
	float* input;          // nwin
	float* ola[3];         // nwin each
	int* iter;             // frame counter, read at start, written back at exit
	unsigned seq0;         // last sequence number already served
	unsigned long long idle_ns;
	int state_in_smem;     // bit 0: ring + tails + previous hop resident in shared memory; bit 1: window / twiddle tables too
	int fenced;            // cluster hand-off of the pushed hop behind fence.acq_rel.cluster on both sides (ZEN_B200_RT_FENCED, see hpr_rt_kernel)
	int cluster;           // CTAs of the thread-block cluster that serves the stream (1: a single CTA)
	cudaStream_t stream;
};

template <int NFFT>
constexpr int nt_rt_for()
{
	// measured on a B200 (tools/rt_phases.py): 512 threads are the sweet spot at nfft 4096 (1024 threads spill in
	// the FFT stages); the large transforms want 1024, the small ones nfft/4
	return NFFT == 4096 ? 512 : ((NFFT / 4) < 128 ? 128 : ((NFFT / 4) > 1024 ? 1024 : (NFFT / 4)));
}
// the cluster-split hop leaves little per-bin work to each CTA: one thread per radix-8 butterfly
template <int NFFT>
constexpr int nt_rt_split_for()
{
	return (NFFT / 8) < 128 ? 128 : ((NFFT / 8) > 512 ? 512 : (NFFT / 8));
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
template <int NFFT> size_t rt_smem_bytes(const HprDev& d, int smem_flags, int cluster);

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
                                                                            long n_hops, int tile_hops, int big_hops, int n_big, int n_tiles, int total_items,
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
		// n_tiles > 0: tile-major order (the short last tiles of the streams close the queue); < 0: stream-major
		const int nt_ = n_tiles < 0 ? -n_tiles : n_tiles, n_str = total_items / nt_;
		const int tile = n_tiles < 0 ? item % nt_ : item / n_str, stream = n_tiles < 0 ? item / nt_ : item - tile * n_str;
		const float* sin = in + (size_t)stream * in_stride;
		// the first n_big tiles of a stream are big_hops long, the rest (the closing ones) tile_hops
		const long e0 = tile < n_big ? (long)tile * big_hops : (long)n_big * big_hops + (long)(tile - n_big) * tile_hops;
		const long e1 = tile < n_big ? e0 + big_hops : min(n_hops, e0 + (long)tile_hops);
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

// The same walk over (stream, tile) work items on the fast path (hpr_fast_iteration: hard mask, copy-border, lag 1).
// Per-CTA scratch: W ring rows of RS = M + 4 floats (16-byte aligned rows) followed by the three overlap-add tails.
// peaks (optional): per stream and output, max |emitted sample| as the bit pattern of a non-negative float
// (atomicMax): the divisor of the command line's peak normalisation (zen/offline.h:180-192), for free.
template <int NFFT>
constexpr int fast_ring_stride()
{
	return NFFT / 2 + 4;
}
template <int NFFT, int NT, bool PEAKS>
__global__ void __launch_bounds__(NT, min_blocks_fast<NT, NFFT>()) hpr_tile_fast_kernel(const __grid_constant__ HprDev P,
                                                                                 const float* __restrict__ in, long in_stride,
                                                                                 float* out_h, float* out_p, float* out_r, long out_stride,
                                                                                 long n_hops, int tile_hops, int big_hops, int n_big, int n_tiles, int total_items,
                                                                                 int* work_counter, float* scratch, size_t scratch_per_cta,
                                                                                 unsigned* peaks_h, unsigned* peaks_p, unsigned* peaks_r)
{
	constexpr int M = NFFT / 2, HOP = M / 2;
	extern __shared__ __align__(16) unsigned char smem_raw[];
	FastSmem<NFFT> sm;
	sm.carve(smem_raw, P.Lp);
	__shared__ int s_item;

	float* sc = scratch + (size_t)blockIdx.x * scratch_per_cta;
	FastState st;
	st.mag_ring = sc;
	st.ring_stride = fast_ring_stride<NFFT>();
	sc += (size_t)P.W * fast_ring_stride<NFFT>();
	for (int o = 0; o < 3; ++o)
		st.tail[o] = sc + (size_t)o * HOP;

	for (;;) {
		__syncthreads();
		if (threadIdx.x == 0)
			s_item = atomicAdd(work_counter, 1);
		__syncthreads();
		const int item = s_item;
		if (item >= total_items)
			break;
		// n_tiles > 0: tile-major order (the short last tiles of the streams close the queue); < 0: stream-major
		const int nt_ = n_tiles < 0 ? -n_tiles : n_tiles, n_str = total_items / nt_;
		const int tile = n_tiles < 0 ? item % nt_ : item / n_str, stream = n_tiles < 0 ? item / nt_ : item - tile * n_str;
		const float* sin = in + (size_t)stream * in_stride;
		// the first n_big tiles of a stream are big_hops long, the rest (the closing ones) tile_hops
		const long e0 = tile < n_big ? (long)tile * big_hops : (long)n_big * big_hops + (long)(tile - n_big) * tile_hops;
		const long e1 = tile < n_big ? e0 + big_hops : min(n_hops, e0 + (long)tile_hops);
		const long i_begin = max(0L, e0 - P.W);
		const long i_full = max(0L, e0 - 1);
		int slot = (int)(i_begin % P.W);
		FastPeaks pk;
		pk.v[0] = pk.v[1] = pk.v[2] = 0.0f;
		for (long i = i_begin; i < e1; ++i) {
			const float* cur = sin + (size_t)i * HOP;
			const float* prev = i > 0 ? cur - HOP : nullptr;
			HprEmit em;
			const bool emit = i >= e0;
			em.a[0] = (emit && out_h) ? out_h + (size_t)stream * out_stride + (size_t)i * HOP : nullptr;
			em.a[1] = (emit && out_p) ? out_p + (size_t)stream * out_stride + (size_t)i * HOP : nullptr;
			em.a[2] = (emit && out_r) ? out_r + (size_t)stream * out_stride + (size_t)i * HOP : nullptr;
			em.b[0] = em.b[1] = em.b[2] = nullptr;
			hpr_fast_iteration<NFFT, NT, PEAKS>(P, sm, st, (int)i, slot, prev, cur, i >= i_full, i == i_full, em,
			                                    i + 1 < e1 ? cur + HOP : nullptr, pk);
			slot = slot + 1 == P.W ? 0 : slot + 1;
		}
		if (PEAKS) {
			unsigned* dst[3] = {peaks_h, peaks_p, peaks_r};
#pragma unroll
			for (int o = 0; o < 3; ++o) {
				if (!dst[o]) continue;
				float m = o == 0 ? pk.v[0] : (o == 1 ? pk.v[1] : pk.v[2]);
				for (int sft = 16; sft > 0; sft >>= 1)
					m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, sft));
				if ((threadIdx.x & 31) == 0 && m > 0.0f) atomicMax(dst[o] + stream, __float_as_uint(m));
			}
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
// The host tags the staging buffer (see RtCtrl); the CTA polls it, runs the fused
// step with the |X| ring, the overlap-add tails, the previous hop AND the window /
// twiddle tables held in SHARED memory, writes the outputs straight to mapped host
// memory (tagged groups) and publishes the completion flag.  No launch and no
// stream synchronisation per hop (those alone cost ~9 us on this box).  The kernel
// leaves on RT_OP_STOP or after idle_ns without a request, writing its state back
// to global memory so the per-launch kernels (or a later session) continue seamlessly.
__device__ __forceinline__ unsigned long long rt_globaltimer()
{
	unsigned long long t;
	asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
	return t;
}

template <int NFFT>
constexpr int rt_table_floats()
{
	// window (nwin) + per-stage twiddles + split twiddles, as floats
	return NFFT / 2 + 2 * fft_twiddle_count<NFFT / 2>() + 2 * (NFFT / 4 + 1);
}
// bins decided per thread in the cluster-split hop (hpr_split_analyse)
template <int NFFT, int NT>
constexpr int rt_us_for()
{
	return 4;
}

// Everything the resident kernel needs from one hop to the next lives in SHARED memory, not in registers or local
// memory: the completion flag's system-wide fence invalidates L1 (CCTL.IVALL), so a spilled pointer or a dynamically
// indexed local array costs an L2 round trip at every use in the next hop.
template <int NFFT>
struct RtShared {
	HprState st;
	HprTables tb;
	HprEmit em;           // per request
	HprPack pk;           // per request
	HprSplit sp;
	float* hopbuf[2];
	unsigned zrecv_off;   // cluster mode: byte offset (in the dynamic shared memory) of the receive buffer of the masked spectrum
	const float* in;      // cached request arguments
	float* out[3];
	int which;
	unsigned op;
	unsigned cmd;         // cluster mode: (request number << 8) | op bits, written by the leader CTA through DSMEM
	unsigned timeout[2];
	unsigned long long stamps[16];
};

// SPLIT: the kernel runs as a thread-block cluster and serves every hop with hpr_split_analyse / hpr_split_synth
// (only that path is compiled in: the two variants stay small enough to keep their state in registers).
template <int NFFT, int NT, bool SPLIT>
__global__ void __launch_bounds__(NT, 1) hpr_rt_kernel(const __grid_constant__ HprDev P, RtCtrl* ctrl, const uint4* stage_in,
                                                       uint4* stage_out_h, uint4* stage_out_p, uint4* stage_out_r, float* g_ring,
                                                       float* g_input, float* g_ola_h, float* g_ola_p, float* g_ola_r, int* g_iter,
                                                       unsigned seq0, unsigned long long idle_ns, int smem_flags)
{
	namespace cg = cooperative_groups;
	constexpr int M = NFFT / 2, HOP = M / 2;
	constexpr int NG = (HOP + 2) / 3;        // tagged groups per hop
	constexpr int PER = (NG + NT - 1) / NT;  // groups polled by one thread
	constexpr int NLG = 4 * (HOP / 12);      // groups that belong to a whole 64-byte line (zen_group_key)
	constexpr int US = rt_us_for<NFFT, NT>();
	(void)US;
	extern __shared__ __align__(16) unsigned char smem_raw[];
	__shared__ RtShared<NFFT> S;
	// the working buffers are addressed off smem_raw directly (the compiler then knows they are shared memory:
	// 32-bit addresses, LDS/STS), everything else through S
	HprSmem<NFFT> sm;
	sm.carve(smem_raw, P.Lp);
	const int tid = threadIdx.x;
	// the cluster spans the whole grid: CTA rank == blockIdx.x.  C == 1 is an ordinary launch.
	const int C = SPLIT ? (int)cg::this_cluster().dim_blocks().x : 1;
	const int rank = SPLIT ? (int)cg::this_cluster().block_rank() : 0;
	const bool leader = rank == 0;
	const int state_in_smem = smem_flags & 1;
	const bool fenced = (smem_flags & 4) != 0;  // every hand-off of the hop to the other CTAs behind a cluster-scope fence
	(void)fenced;
	const int ring_n = P.W * (M + 1);

	if (tid == 0) {
		float* extra = reinterpret_cast<float*>(smem_raw + ((HprSmem<NFFT>::bytes(P.Lp) + 15) & ~(size_t)15));
		S.pk.buf = extra;  // HOP floats: the emitted hop before it is packed into tagged groups
		extra += HOP;
		// double buffer: previous hop / incoming hop (in global memory, i.e. the `input` vector itself, when smem is short)
		S.hopbuf[0] = state_in_smem ? extra : g_input;
		S.hopbuf[1] = state_in_smem ? extra + HOP : g_input + HOP;
		if (state_in_smem) {
			extra += 2 * HOP;
			S.st.mag_ring = extra;
			extra += (ring_n + 3) & ~3;
			for (int o = 0; o < 3; ++o)
				S.st.tail[o] = extra + o * HOP;
			extra += 3 * HOP;
		}
		else {
			S.st.mag_ring = g_ring;
			S.st.tail[0] = g_ola_h + HOP;
			S.st.tail[1] = g_ola_p + HOP;
			S.st.tail[2] = g_ola_r + HOP;
		}
		S.st.x_ring = nullptr;
		S.st.xdepth = 0;
		S.tb.window = P.window;
		S.tb.tw = P.tw;
		S.tb.twr = P.twr;
		if (smem_flags & 2) {
			constexpr int NTW = fft_twiddle_count<M>();
			float* w_s = extra;
			float2* tw_s = reinterpret_cast<float2*>(w_s + M);
			S.tb.window = w_s;
			S.tb.tw = tw_s;
			S.tb.twr = tw_s + NTW;
			extra += rt_table_floats<NFFT>();
		}
		S.zrecv_off = (unsigned)(reinterpret_cast<unsigned char*>(extra) - smem_raw);  // only carved (and only used) in cluster mode
		// cluster split: own pairs / bins, owners of the outputs in emission order P, H, R
		S.sp.rank = rank;
		S.sp.C = C;
		hpr_split_ranges(M, rank, C, S.sp.k0, S.sp.k1, S.sp.a0, S.sp.a1, S.sp.b0, S.sp.b1);
		int n_on = 0;
		for (int oi = 0; oi < 3; ++oi) {
			const int o = oi == 0 ? 1 : (oi == 1 ? 0 : 2);
			const bool on = (P.out_flags & (1 << o)) != 0 && !(o == 2 && (P.soft || P.sse));
			S.sp.owner[o] = on ? (n_on++ % C) : -1;
			S.sp.recv[o] = nullptr;
			if (SPLIT && on) S.sp.recv[o] = cg::this_cluster().map_shared_rank(reinterpret_cast<float2*>(smem_raw + S.zrecv_off), S.sp.owner[o]);
		}
		S.in = nullptr;
		S.out[0] = S.out[1] = S.out[2] = nullptr;
		S.which = 0;
		S.timeout[0] = S.timeout[1] = 0u;
		S.cmd = seq0 << 8;
	}
	__syncthreads();
	if (state_in_smem) {
		float* ring = S.st.mag_ring;
		for (int n = tid; n < ring_n; n += NT)
			ring[n] = g_ring[n];
		for (int n = tid; n < HOP; n += NT) {
			S.st.tail[0][n] = g_ola_h[HOP + n];
			S.st.tail[1][n] = g_ola_p[HOP + n];
			S.st.tail[2][n] = g_ola_r[HOP + n];
			S.hopbuf[0][n] = g_input[n];          // older hop (only kept so `input` can be written back)
			S.hopbuf[1][n] = g_input[HOP + n];    // the previous hop
		}
	}
	if (smem_flags & 2) {
		// tables: every load of the per-hop path then has shared-memory latency
		constexpr int NTW = fft_twiddle_count<M>();
		float* w_s = const_cast<float*>(S.tb.window);
		float2* tw_s = const_cast<float2*>(S.tb.tw);
		float2* twr_s = const_cast<float2*>(S.tb.twr);
		for (int n = tid; n < M; n += NT)
			w_s[n] = P.window[n];
		for (int n = tid; n < NTW; n += NT)
			tw_s[n] = P.tw[n];
		for (int n = tid; n <= M / 2; n += NT)
			twr_s[n] = P.twr[n];
	}
	int prev_idx = 1;
	int iter = *g_iter;
	unsigned seq = seq0;
	unsigned exit_reason = 0;
	__syncthreads();
	if (SPLIT) cg::this_cluster().sync();  // every CTA's shared memory is set up before the leader may write into it

	// thread 0: everything the CTA needs to know about request `seqn` goes into S before the barrier that starts it
	auto fill_request = [&](unsigned opw_, unsigned seqn) {
		const bool tagged = (opw_ & RT_F_TAG_OUT) != 0;
		S.op = opw_;
		S.pk.tag = (seqn << 8) | opw_;
		S.em.a[0] = (P.out_flags & 1) ? g_ola_h : nullptr;   // keep the public *_out vectors current (hps.h:195-197)
		S.em.a[1] = (P.out_flags & 2) ? g_ola_p : nullptr;
		S.em.a[2] = (P.out_flags & 4) ? g_ola_r : nullptr;
		S.em.b[0] = ((P.out_flags & 1) && !tagged) ? S.out[0] : nullptr;
		S.em.b[1] = ((P.out_flags & 2) && !tagged) ? S.out[1] : nullptr;
		S.em.b[2] = ((P.out_flags & 4) && !tagged) ? S.out[2] : nullptr;
		S.pk.dst[0] = ((P.out_flags & 1) && tagged) ? stage_out_h : nullptr;
		S.pk.dst[1] = ((P.out_flags & 2) && tagged) ? stage_out_p : nullptr;
		S.pk.dst[2] = ((P.out_flags & 4) && tagged) ? stage_out_r : nullptr;
	};

	for (;;) {
		unsigned opw = 0u;
		if (leader) {
			// ---- wait for the next request: every thread polls its own tagged group(s).  In a cluster only the leader
			// polls (PCIe read requests are the scarce resource: four CTAs polling cost 4-5 us per hop); it forwards the
			// hop and the command to the other CTAs through distributed shared memory.
			const unsigned want = (seq + 1u) << 8;
			uint4 v[PER];
			unsigned w_raw[PER];
			unsigned pending = 0u;
#pragma unroll
			for (int b = 0; b < PER; ++b) {
				v[b] = make_uint4(0u, 0u, 0u, ~want);
				w_raw[b] = ~want;
				if (tid + b * NT < NG) pending |= 1u << b;
			}
			bool got = pending == 0u;
			unsigned round = 0;
			bool all = false;
			unsigned long long t0 = 0;
			if (tid == 0) t0 = rt_globaltimer();
			for (;;) {
#pragma unroll
				for (int b = 0; b < PER; ++b) {
					if (pending & (1u << b))
						asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];"
						             : "=r"(v[b].x), "=r"(v[b].y), "=r"(v[b].z), "=r"(w_raw[b])
						             : "l"(stage_in + tid + b * NT)
						             : "memory");
				}
#pragma unroll
				for (int b = 0; b < PER; ++b) {
					// The tag word is stored XOR a hash of the samples (zen_group_key: per 64-byte line, per group behind
					// the last whole line): a group caught half-written does not match.  A line is taken as a whole.
					const int g = tid + b * NT;
					const unsigned t = w_raw[b] ^ zen_group_key(v[b].x, v[b].y, v[b].z, g, NLG);
					unsigned ok = ((t ^ want) <= 0xffu) ? 1u : 0u;
					if (g < NLG) {
						ok &= __shfl_xor_sync(0xffffffffu, ok, 1);
						ok &= __shfl_xor_sync(0xffffffffu, ok, 2);
					}
					else {
						(void)__shfl_xor_sync(0xffffffffu, ok, 1);
						(void)__shfl_xor_sync(0xffffffffu, ok, 2);
					}
					if ((pending & (1u << b)) && ok) {
						pending &= ~(1u << b);
						v[b].w = t;
					}
				}
				got = pending == 0u;
				if (tid == 0)
					S.timeout[round & 1u] = (!got && (round & 63u) == 63u && rt_globaltimer() - t0 > idle_ns) ? 1u : 0u;
				const int n_got = __syncthreads_count(got);
				if (n_got == NT) {
					all = true;
					break;
				}
				if (S.timeout[round & 1u])
					break;
				++round;
			}
			// (uniform from here: the count and the flag are the same for every thread)
			if (all) {
				opw = v[0].w & 0xffu;
				if (tid == 0) {
					if (opw & RT_F_NEW_ARGS) {
						// pointers changed since the last request: fetch them (one more round trip), else reuse the cached ones
						S.in = *reinterpret_cast<const float* volatile*>(&ctrl->in);
						S.out[0] = *reinterpret_cast<float* volatile*>(&ctrl->out[0]);
						S.out[1] = *reinterpret_cast<float* volatile*>(&ctrl->out[1]);
						S.out[2] = *reinterpret_cast<float* volatile*>(&ctrl->out[2]);
						S.which = *reinterpret_cast<volatile int*>(&ctrl->which);
					}
					if (opw & RT_F_STAMPS) {
						S.stamps[9] = rt_globaltimer();
						S.stamps[10] = (unsigned long long)clock64();
					}
				}
			}
			else {
				opw = RT_OP_EXIT;  // idle time-out
			}
			const bool pushed = (opw & RT_F_PUSH_IN) && (opw & RT_OP_MASK) == RT_OP_PROCESS;
			float* stash = S.hopbuf[prev_idx ^ 1];
			if (pushed) {
#pragma unroll
				for (int b = 0; b < PER; ++b) {
					const int g = tid + b * NT;
					if (g < NG) {
						stash[3 * g] = __uint_as_float(v[b].x);
						if (3 * g + 1 < HOP) stash[3 * g + 1] = __uint_as_float(v[b].y);
						if (3 * g + 2 < HOP) stash[3 * g + 2] = __uint_as_float(v[b].z);
					}
				}
			}
			if (tid == 0) fill_request(opw, seq + 1u);
			__syncthreads();  // the request starts here: the stash and S are complete
			if constexpr (SPLIT) {
				// The command goes to the other CTAs by a few remote stores; the hop itself stays here and is PULLED by
				// them while they window it (rt_microbench: pushing 1026 floats costs ~0.7 us per CTA with 4-byte stores,
				// pulling them ~0.35 us for all CTAs at once).  The command word doubles as the flag the other CTAs spin
				// on: no cluster barrier (0.2 us + the leader waiting for it) at the start of a hop.
				// The LAST thread signals, after the barrier, while the other warps already start on the hop.  The
				// command word is released at cluster scope (fence.acq_rel.cluster before the store, and again in the
				// reader once it has seen the word), as the PTX memory model asks for a hand-off through another CTA's
				// shared memory.  Measured: the fences sit on one thread while the other warps work and cost nothing
				// at the host (p50 10.68 us with, 10.70 us without, profiles/r02_rt_latency_fenced.txt).  Without
				// them the protocol has never delivered a stale word either (2 x 10^6 rounds x 7 readers of
				// tests/cpp/handoff_litmus.cu: shared memory is one physical copy and BAR.SYNC has drained the
				// stores), but that is evidence, not a guarantee: `fenced` is the default, ZEN_B200_RT_FENCED=0
				// the experiment.
				if (tid == NT - 1) {
					const unsigned cw = S.op;  // (opw is thread 0's: this thread may not own a tagged group)
					if (cw & RT_F_NEW_ARGS) {
						for (int r = 1; r < C; ++r) {
							RtShared<NFFT>* R = cg::this_cluster().map_shared_rank(&S, r);
							R->in = S.in;
							R->out[0] = S.out[0];
							R->out[1] = S.out[1];
							R->out[2] = S.out[2];
							R->which = S.which;
						}
						asm volatile("fence.acq_rel.cluster;" ::: "memory");
					}
					for (int r = 1; r < C; ++r)
						*reinterpret_cast<volatile unsigned*>(&cg::this_cluster().map_shared_rank(&S, r)->cmd) = ((seq + 1u) << 8) | cw;
				}
			}
		}
		else {
			// the other CTAs of the cluster wait for the leader's command word
			if (tid == 0) {
				volatile unsigned* cmd = &S.cmd;
				unsigned c;
				do {
					c = *cmd;
				} while (((c ^ ((seq + 1u) << 8)) >> 8) != 0u);
				if (c & RT_F_NEW_ARGS)
					asm volatile("fence.acq_rel.cluster;" ::: "memory");
				else
					asm volatile("" ::: "memory");
				fill_request(c & 0xffu, seq + 1u);
				if ((c & RT_F_STAMPS) && rank == 1) {
					S.stamps[9] = rt_globaltimer();
					S.stamps[10] = (unsigned long long)clock64();
				}
			}
			__syncthreads();
		}
		opw = S.op;
		const unsigned op = opw & RT_OP_MASK;
		if (op == RT_OP_EXIT) {
			exit_reason = 2u;
			break;
		}
		++seq;
		if (op == RT_OP_STOP) {
			exit_reason = 1u;
			break;
		}
		const bool tagged = (opw & RT_F_TAG_OUT) != 0;
		if (op == RT_OP_PROCESS) {
			const bool pushed = (opw & RT_F_PUSH_IN) != 0;
			unsigned long long* stamps = ((opw & RT_F_STAMPS) && (leader || rank == 1)) ? S.stamps : nullptr;
			if constexpr (!SPLIT) {
				hpr_iteration<NFFT, NT, rt_u_for<NFFT>(), true>(P, sm, S.st, iter, S.hopbuf[prev_idx],
				                                                pushed ? S.hopbuf[prev_idx ^ 1] : S.in, true, false, S.em,
				                                                pushed ? nullptr : S.hopbuf[prev_idx ^ 1], stamps, nullptr, &S.tb, &S.pk);
				if ((P.out_flags & ZEN_OUTPUT_RESIDUAL) && (P.soft || P.sse) && !tagged && S.out[2])
					for (int n = tid; n < HOP; n += NT)
						S.out[2][n] = 0.0f;
			}
			else {
				if (!pushed) {  // pulled hop: every CTA fetches it into its own stash first
					const float* src = S.in;
					float* stash = S.hopbuf[prev_idx ^ 1];
					for (int n = tid; n < HOP; n += NT)
						stash[n] = src[n];
					__syncthreads();
				}
				// a pushed hop sits in the leader's stash: the other CTAs read it from there and keep a copy as the next `prev`
				float* stash = S.hopbuf[prev_idx ^ 1];
				const bool remote = pushed && !leader;
				const float* cur = remote ? cg::this_cluster().map_shared_rank(stash, 0) : stash;
				hpr_split_analyse<NFFT, NT, US>(P, sm, S.st, iter, S.hopbuf[prev_idx], cur, remote ? stash : nullptr, S.tb, S.sp,
				                                reinterpret_cast<float2*>(smem_raw + S.zrecv_off) + 2 * fpad_size(M), (iter & 1) * fpad_size(M), stamps);
				cg::this_cluster().sync();  // the masked spectra have arrived at their owners
				// (two receive buffers, by request parity: a CTA may already receive the next hop's spectrum while it
				// still finishes this one's overlap-add)
				float2* zrecv = reinterpret_cast<float2*>(smem_raw + S.zrecv_off) + (iter & 1) * fpad_size(M);
				float2* zpp = reinterpret_cast<float2*>(smem_raw + S.zrecv_off) + 2 * fpad_size(M);
				if (S.sp.owner[1] == rank)
					hpr_split_synth<NFFT, NT>(P, zrecv, zpp, S.st.tail[1], S.em.a[1], S.em.b[1], S.pk.buf, S.pk.dst[1], S.pk.tag, S.tb, stamps);
				if (S.sp.owner[0] == rank)
					hpr_split_synth<NFFT, NT>(P, zrecv, zpp, S.st.tail[0], S.em.a[0], S.em.b[0], S.pk.buf, S.pk.dst[0], S.pk.tag, S.tb, nullptr);
				if (S.sp.owner[2] == rank)
					hpr_split_synth<NFFT, NT>(P, zrecv, zpp, S.st.tail[2], S.em.a[2], S.em.b[2], S.pk.buf, S.pk.dst[2], S.pk.tag, S.tb, nullptr);
			}
			prev_idx ^= 1;
			++iter;
			if (tid == 0 && stamps) {
				S.stamps[11] = (unsigned long long)clock64();
				S.stamps[12] = rt_globaltimer();
			}
		}
		else if (op == RT_OP_COPY && leader) {
			const float* src = S.which == 0 ? g_ola_h : (S.which == 1 ? g_ola_p : g_ola_r);
			float* dst = S.out[0];
			for (int n = tid; n < HOP; n += NT)
				dst[n] = __ldcg(src + n);  // written by another CTA in cluster mode: not through this SM's L1
		}
		// Completion flag, needed only when the host is not already watching the tagged output groups.  A release
		// store: the fence in front of it orders the plain output stores, and - unlike __threadfence_system() - no L1
		// invalidation follows it.  (In a cluster the next request's first barrier is what keeps the CTAs in step.)
		if (!(op == RT_OP_PROCESS && tagged)) {
			if (SPLIT)
				cg::this_cluster().sync();  // the owners of the outputs have stored them
			else
				__syncthreads();
			if (leader && tid == 0)
				asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(&ctrl->seq_out), "r"(seq) : "memory");
		}
		if ((leader || rank == 1) && op == RT_OP_PROCESS && (opw & RT_F_STAMPS)) {
			__syncthreads();
			if (tid < 16) (leader ? ctrl->stamps : ctrl->stamps_r1)[tid] = S.stamps[tid];  // diagnostics
		}
	}

	// ---- leave: write the state back (each CTA the part it owns)
	__syncthreads();
	if (state_in_smem) {
		const float* ring = S.st.mag_ring;
		const int a0 = S.sp.a0, nA = S.sp.a1 - S.sp.a0, b0 = S.sp.b0, nB = S.sp.b1 - S.sp.b0;
		for (int r = 0; r < P.W; ++r)
			for (int idx = tid; idx < nA + nB; idx += NT) {
				const int k = idx < nA ? a0 + idx : b0 + (idx - nA);
				g_ring[(size_t)r * (M + 1) + k] = ring[(size_t)r * (M + 1) + k];
			}
		// overlap-add tails: the owner of an output; the tails of disabled outputs pass through the leader unchanged
		const bool w0 = S.sp.owner[0] == rank || (S.sp.owner[0] < 0 && leader);
		const bool w1 = S.sp.owner[1] == rank || (S.sp.owner[1] < 0 && leader);
		const bool w2 = S.sp.owner[2] == rank || (S.sp.owner[2] < 0 && leader);
		for (int n = tid; n < HOP; n += NT) {
			if (w0) g_ola_h[HOP + n] = S.st.tail[0][n];
			if (w1) g_ola_p[HOP + n] = S.st.tail[1][n];
			if (w2) g_ola_r[HOP + n] = S.st.tail[2][n];
		}
	}
	if (leader && (state_in_smem || prev_idx == 0)) {
		// `input` must read [older hop | newest hop]; when it served as the double buffer itself, swap the halves if needed
		const float* ho = S.hopbuf[prev_idx ^ 1];
		const float* hn = S.hopbuf[prev_idx];
		for (int n = tid; n < HOP; n += NT) {
			float older = ho[n], newest = hn[n];
			g_input[n] = older;
			g_input[HOP + n] = newest;
		}
	}
	if (leader && tid == 0) *g_iter = iter;
	__syncthreads();
	if (SPLIT) cg::this_cluster().sync();  // nobody touches another CTA's shared memory after this point
	if (leader && tid == 0) {
		ctrl->exit_reason = exit_reason;
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
	const long hops_big = (long)a.n_big * a.big_hops;
	const int n_tiles = a.n_big + (int)((a.n_hops - hops_big + a.tile_hops - 1) / a.tile_hops);
	const long total = (long)n_tiles * a.n_streams;
	static const int order = std::getenv("ZEN_B200_STREAM_MAJOR") ? -1 : 1;  // A/B switch of the queue order
	if (total > 0x7fffffffL)
		return ZEN_ERR_UNSUPPORTED;
	ZEN_CUDA_CHECK(cudaMemsetAsync(a.work_counter, 0, sizeof(int), a.stream));
	const int grid = (int)std::min<long>(total, a.resident_ctas);
	const bool want_peaks = a.peaks[0] || a.peaks[1] || a.peaks[2];
	if (hpr_fast_supported(a.dev) && !a.force_general) {
		size_t smem = FastSmem<NFFT>::bytes(a.dev.Lp);
		auto launch = [&](auto kern) -> int {
			ZEN_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
			ZEN_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
			kern<<<grid, NT, smem, a.stream>>>(a.dev, a.in, a.in_stride, a.out_h, a.out_p, a.out_r, a.out_stride, a.n_hops, a.tile_hops, a.big_hops, a.n_big,
			                                   order * n_tiles, (int)total, a.work_counter, a.scratch, a.scratch_per_cta, a.peaks[0], a.peaks[1],
			                                   a.peaks[2]);
			ZEN_CUDA_CHECK(cudaGetLastError());
			return ZEN_OK;
		};
		return want_peaks ? launch(hpr_tile_fast_kernel<NFFT, NT, true>) : launch(hpr_tile_fast_kernel<NFFT, NT, false>);
	}
	if (want_peaks)
		return ZEN_ERR_UNSUPPORTED;  // the caller falls back to a separate peak pass
	auto kern = hpr_tile_kernel<NFFT, NT>;
	size_t smem = HprSmem<NFFT>::bytes(a.dev.Lp);
	ZEN_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	ZEN_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
	kern<<<grid, NT, smem, a.stream>>>(a.dev, a.in, a.in_stride, a.out_h, a.out_p, a.out_r, a.out_stride, a.n_hops, a.tile_hops, a.big_hops, a.n_big,
	                                   order * n_tiles, (int)total, a.work_counter, a.scratch, a.scratch_per_cta);
	ZEN_CUDA_CHECK(cudaGetLastError());
	return ZEN_OK;
}

// resident CTAs per SM of the tile kernel (occupancy query), times the SM count
template <int NFFT>
int tile_resident_ctas(const HprDev& d)
{
	constexpr int NT = nt_for<NFFT>();
	int per_sm = 0, dev = 0, sms = 0;
	auto occ = [&](auto kern, size_t smem) -> bool {
		return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) == cudaSuccess
		       && cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared) == cudaSuccess
		       && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NT, smem) == cudaSuccess;
	};
	// the general kernel needs at least as much shared memory as the fast one: size the scratch for whichever holds more CTAs
	int best = 0;
	if (occ(hpr_tile_kernel<NFFT, NT>, HprSmem<NFFT>::bytes(d.Lp))) best = per_sm;
	if (hpr_fast_supported(d) && occ(hpr_tile_fast_kernel<NFFT, NT, false>, FastSmem<NFFT>::bytes(d.Lp))) {
		int f = per_sm;
		if (occ(hpr_tile_fast_kernel<NFFT, NT, true>, FastSmem<NFFT>::bytes(d.Lp))) f = std::min(f, per_sm);
		best = std::max(best, f);
	}
	if (best < 1 || cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
		return 0;
	return best * sms;
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
size_t rt_smem_bytes(const HprDev& d, int smem_flags, int cluster)
{
	constexpr int M = NFFT / 2, HOP = M / 2;
	size_t b = (HprSmem<NFFT>::bytes(d.Lp) + 15) & ~(size_t)15;
	b += sizeof(float) * (size_t)HOP;  // packbuf
	if (smem_flags & 1)
		b += sizeof(float) * (2 * (size_t)HOP + (((size_t)d.W * (M + 1) + 3) & ~(size_t)3) + 3 * (size_t)HOP);
	if (smem_flags & 2)
		b += sizeof(float) * (size_t)rt_table_floats<NFFT>();
	if (cluster > 1)
		b += 3 * sizeof(float2) * (size_t)fpad_size(M);  // two receive buffers of the masked spectrum (request parity) + ping-pong partner of the FFTs
	return b;
}

template <int NFFT, int NT, bool SPLIT>
int launch_rt_variant(const RtArgs& a)
{
	auto kern = hpr_rt_kernel<NFFT, NT, SPLIT>;
	size_t smem = rt_smem_bytes<NFFT>(a.dev, a.state_in_smem, a.cluster);
	ZEN_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	cudaLaunchConfig_t cfg = {};
	cfg.gridDim = dim3((unsigned)a.cluster);
	cfg.blockDim = dim3(NT);
	cfg.dynamicSmemBytes = smem;
	cfg.stream = a.stream;
	cudaLaunchAttribute attr[1];
	attr[0].id = cudaLaunchAttributeClusterDimension;
	attr[0].val.clusterDim.x = (unsigned)a.cluster;
	attr[0].val.clusterDim.y = 1;
	attr[0].val.clusterDim.z = 1;
	cfg.attrs = attr;
	cfg.numAttrs = SPLIT ? 1 : 0;
	ZEN_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, a.dev, a.ctrl, a.stage_in, a.stage_out[0], a.stage_out[1], a.stage_out[2], a.mag_ring,
	                                  a.input, a.ola[0], a.ola[1], a.ola[2], a.iter, a.seq0, a.idle_ns, a.state_in_smem | (a.fenced ? 4 : 0)));
	return ZEN_OK;
}

template <int NFFT>
int launch_rt_impl(const RtArgs& a)
{
	if (a.cluster > 1)
		return launch_rt_variant<NFFT, nt_rt_split_for<NFFT>(), true>(a);
	return launch_rt_variant<NFFT, nt_rt_for<NFFT>(), false>(a);
}

template int launch_tile_impl<ZEN_HPR_INSTANTIATE>(const TileArgs&);
template int tile_resident_ctas<ZEN_HPR_INSTANTIATE>(const HprDev&);
template int launch_hop_impl<ZEN_HPR_INSTANTIATE>(const HopArgs&);
template int launch_rt_impl<ZEN_HPR_INSTANTIATE>(const RtArgs&);
template size_t rt_smem_bytes<ZEN_HPR_INSTANTIATE>(const HprDev&, int, int);

}  // namespace zen_b200
#endif
