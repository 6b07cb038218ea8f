// Can the resident real-time kernel receive the pushed hop by ONE bulk (TMA) read of the mapped host staging buffer,
// multicast into the shared memory of every CTA of its cluster, instead of 342 polling threads in the leader CTA
// followed by the other CTAs pulling the hop out of the leader's shared memory?
//
// Measures, against mapped pinned host memory that the host keeps rewriting:
//   1. rounds / s of   (a) 342 threads x ld.volatile.v4 (today's poll round)
//                      (b) one cp.async.bulk global -> shared::cta of the same 5472 bytes
//                      (c) the same copy multicast to a cluster of C CTAs
//   2. whether bulk reads of host memory ever return stale data (the host bumps a sequence number in every group)
//   3. the host-observed round trip: host writes the groups, kernel detects them in ALL CTAs, CTA 0 writes an ack
//      word to mapped memory, host sees it - for the three variants.
// build: nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a tools/tma_poll_microbench.cu -o tools/_build/tma_poll_microbench
#include <chrono>
#include <cooperative_groups.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>
#include <emmintrin.h>
#include <algorithm>
#include <vector>

namespace cg = cooperative_groups;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { std::fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); std::exit(1); } } while (0)

constexpr int NG = 342;            // groups of a 1024-sample hop
constexpr int BYTES = NG * 16;     // 5472
constexpr int NT = 512;

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned long long gtime()
{
	unsigned long long t;
	asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
	return t;
}
__device__ __forceinline__ void mbar_init(unsigned long long* b, unsigned n)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(n));
}
__device__ __forceinline__ void mbar_expect(unsigned long long* b, unsigned bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try(unsigned long long* b, unsigned parity)
{
	unsigned ok;
	asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
	return ok != 0;
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar)
{
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_g2s_mc(void* dst, const void* src, unsigned bytes, unsigned long long* bar, unsigned short mask)
{
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "h"(mask) : "memory");
}

struct Ctl {
	volatile unsigned ack;       // written by the kernel: last sequence number every CTA has seen complete
	unsigned pad[15];
	volatile unsigned long long rounds;  // poll rounds done so far (leader)
	volatile unsigned long long stale;   // rounds whose data went BACKWARDS in sequence (stale read)
	volatile unsigned long long torn;    // rounds in which the groups carried more than one sequence number
	volatile unsigned long long disagree;  // multicast: a CTA saw a complete hop in a different round than the leader
};

constexpr unsigned STOP_SEQ = 0xffffffffu;
__device__ __forceinline__ void spin_guard(unsigned& n)
{
	if (++n > (1u << 26)) __trap();  // never hang the box: a protocol error aborts the kernel
}

// MODE 0: 342 threads poll with ld.volatile.v4 (one CTA)
// MODE 1: bulk copies into two alternating buffers, issued by the leader, multicast when the cluster has more than one
//         CTA; one copy in flight
// MODE 2: the same with TWO copies in flight (the next one is issued before the current one is waited for)
// A CTA arms its mbarrier for round n + 1 before it waits for round n and tells the leader (a counter in the leader's
// shared memory); the leader issues round n only when every CTA has armed it, so a barrier never receives the bytes
// of a round it has not armed.
template <int MODE>
__global__ void __launch_bounds__(NT, 1) poll_kernel(const uint4* __restrict__ stage, Ctl* ctl, unsigned* seen /* per CTA: last complete seq */)
{
	__shared__ __align__(128) uint4 buf[2][NG];
	__shared__ __align__(8) unsigned long long bar[2];
	__shared__ unsigned s_min, s_max;
	__shared__ volatile unsigned armed[8];  // leader: rounds armed by CTA r
	__shared__ volatile unsigned done[8];   // leader: last complete sequence number seen by CTA r
	__shared__ volatile unsigned done_round[8];
	auto cluster = cg::this_cluster();
	const int C = MODE ? (int)cluster.dim_blocks().x : 1;
	const int rank = MODE ? (int)cluster.block_rank() : 0;
	const int tid = threadIdx.x;
	if (tid == 0) {
		mbar_init(&bar[0], 1);
		mbar_init(&bar[1], 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		for (int r = 0; r < 8; ++r)
			armed[r] = done[r] = done_round[r] = 0;
	}
	__syncthreads();
	if (MODE) cluster.sync();
	volatile unsigned* l_armed = MODE ? cluster.map_shared_rank(const_cast<unsigned*>(&armed[0]), 0) : &armed[0];
	volatile unsigned* l_done = MODE ? cluster.map_shared_rank(const_cast<unsigned*>(&done[0]), 0) : &done[0];
	volatile unsigned* l_done_round = MODE ? cluster.map_shared_rank(const_cast<unsigned*>(&done_round[0]), 0) : &done_round[0];
	const unsigned short mask = (unsigned short)((1u << C) - 1u);
	unsigned last = 0, issued = 0;
	unsigned long long rounds = 0, stale = 0, torn = 0, disagree = 0;
	auto issue = [&](unsigned n) {  // leader, thread 0
		unsigned g = 0;
		for (int r = 0; r < C; ++r)
			while (armed[r] < n + 1) spin_guard(g);
		if (C > 1)
			bulk_g2s_mc(buf[n & 1], stage, BYTES, &bar[n & 1], mask);
		else
			bulk_g2s(buf[n & 1], stage, BYTES, &bar[n & 1]);
	};
	if (MODE && tid == 0) {
		mbar_expect(&bar[0], BYTES);
		l_armed[rank] = 1;
	}
	for (unsigned n = 0;; ++n) {
		uint4 v = make_uint4(0, 0, 0, 0);
		if (MODE == 0) {
			if (tid < NG) asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(stage + tid));
		}
		else {
			if (tid == 0) {
				mbar_expect(&bar[(n + 1) & 1], BYTES);  // arm the next round, tell the leader
				l_armed[rank] = n + 2;
				if (rank == 0) {
					while (issued <= n + (MODE == 2 ? 1u : 0u)) {
						issue(issued);
						++issued;
					}
				}
				unsigned g = 0;
				while (!mbar_try(&bar[n & 1], (n >> 1) & 1)) spin_guard(g);
			}
			__syncthreads();
			if (tid < NG) v = buf[n & 1][tid];
		}
		// every group carries its sequence number in .w; the hop is complete when all groups agree
		if (tid == 0) {
			s_min = 0xffffffffu;
			s_max = 0;
		}
		__syncthreads();
		if (tid < NG) {
			atomicMin(&s_min, v.w);
			atomicMax(&s_max, v.w);
		}
		__syncthreads();
		const unsigned mn = s_min, mx = s_max;
		++rounds;
		if (mn != mx) ++torn;
		if (mx < last) ++stale;
		if (mn == mx && mn == STOP_SEQ) {
			if (MODE == 2 && tid == 0) {  // a second copy is still in flight: it must land before this CTA's shared memory goes away
				unsigned g = 0;
				while (!mbar_try(&bar[(n + 1) & 1], ((n + 1) >> 1) & 1)) spin_guard(g);
			}
			break;
		}
		if (mn == mx && mn > last) {
			last = mn;
			if (tid == 0) {
				seen[rank] = last;
				l_done_round[rank] = n;
				l_done[rank] = last;
				if (rank == 0) {
					unsigned g = 0;
					for (int r = 0; r < C; ++r) {
						while (done[r] != last) {
							spin_guard(g);
							if (MODE && issued <= n + 1 && g > 4096u) {  // a CTA missed this round: keep the rounds coming
								issue(issued);
								++issued;
							}
						}
						if (done_round[r] != n) ++disagree;
					}
					ctl->ack = last;
				}
			}
		}
		__syncthreads();
	}
	if (rank == 0 && tid == 0) {
		ctl->rounds = rounds;
		ctl->stale = stale;
		ctl->torn = torn;
		ctl->disagree = disagree;
	}
	if (MODE) cluster.sync();
}

static double now_us()
{
	return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

template <int MODE>
static void run(const char* name, int C, uint4* h_stage, uint4* d_stage, Ctl* h_ctl, Ctl* d_ctl, unsigned* d_seen, int iters)
{
	std::memset((void*)h_ctl, 0, sizeof(Ctl));
	std::memset(h_stage, 0, BYTES);
	_mm_sfence();
	cudaLaunchConfig_t cfg = {};
	cfg.gridDim = dim3((unsigned)C);
	cfg.blockDim = dim3(NT);
	cudaLaunchAttribute attr[1];
	attr[0].id = cudaLaunchAttributeClusterDimension;
	attr[0].val.clusterDim.x = (unsigned)C;
	attr[0].val.clusterDim.y = 1;
	attr[0].val.clusterDim.z = 1;
	cfg.attrs = attr;
	cfg.numAttrs = 1;
	cudaStream_t s;
	CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
	cfg.stream = s;
	CK(cudaLaunchKernelEx(&cfg, poll_kernel<MODE>, (const uint4*)d_stage, d_ctl, d_seen));
	std::vector<double> lat;
	lat.reserve(iters);
	const double t_begin = now_us();
	bool failed = false;
	for (int k = 1; k <= iters && !failed; ++k) {
		// ~20 us between hops so that the kernel is idle-polling when the hop arrives
		const double w = now_us();
		while (now_us() - w < 20.0) {
		}
		const double t0 = now_us();
		for (int g = 0; g < NG; ++g) {
			__m128i x = _mm_set_epi32(k, g * 3 + 2 + k, g * 3 + 1 + k, g * 3 + k);
			_mm_store_si128((__m128i*)(h_stage + g), x);
		}
		_mm_sfence();
		while (h_ctl->ack != (unsigned)k) {
			if (now_us() - t0 > 2e6) {
				failed = true;
				break;
			}
		}
		lat.push_back(now_us() - t0);
	}
	const double t_end = now_us();
	for (int g = 0; g < NG; ++g)
		_mm_store_si128((__m128i*)(h_stage + g), _mm_set1_epi32((int)0xffffffffu));
	_mm_sfence();
	cudaError_t rc = cudaStreamSynchronize(s);
	std::vector<unsigned> seen(8, 0);
	cudaMemcpy(seen.data(), d_seen, 8 * sizeof(unsigned), cudaMemcpyDeviceToHost);
	std::sort(lat.begin(), lat.end());
	const double p50 = lat.empty() ? -1 : lat[lat.size() / 2], p99 = lat.empty() ? -1 : lat[(lat.size() * 99) / 100], mn = lat.empty() ? -1 : lat[0];
	std::printf(" \"%s\": {\"cluster\": %d, \"ok\": %s, \"cuda\": \"%s\", \"round_trip_p50_us\": %.2f, \"p99_us\": %.2f, \"min_us\": %.2f, "
	            "\"poll_round_us\": %.3f, \"rounds\": %llu, \"stale_rounds\": %llu, \"torn_rounds\": %llu, \"ctas_disagreeing_on_the_round\": %llu, \"seen_last\": [%u, %u]},\n",
	            name, C, failed ? "false" : "true", cudaGetErrorString(rc), p50, p99, mn,
	            h_ctl->rounds ? (t_end - t_begin) / (double)h_ctl->rounds : -1.0, (unsigned long long)h_ctl->rounds,
	            (unsigned long long)h_ctl->stale, (unsigned long long)h_ctl->torn, (unsigned long long)h_ctl->disagree, seen[0], seen[C - 1]);
	cudaStreamDestroy(s);
}

int main(int argc, char** argv)
{
	const int iters = argc > 1 ? std::atoi(argv[1]) : 3000;
	uint4 *h_stage, *d_stage;
	Ctl *h_ctl, *d_ctl;
	unsigned* d_seen;
	CK(cudaSetDeviceFlags(cudaDeviceMapHost));
	CK(cudaHostAlloc((void**)&h_stage, 8192, cudaHostAllocMapped));
	CK(cudaHostAlloc((void**)&h_ctl, sizeof(Ctl), cudaHostAllocMapped));
	CK(cudaHostGetDevicePointer((void**)&d_stage, h_stage, 0));
	CK(cudaHostGetDevicePointer((void**)&d_ctl, h_ctl, 0));
	CK(cudaMalloc(&d_seen, 64));
	std::printf("{\n");
	run<0>("threads_ld_volatile", 1, h_stage, d_stage, h_ctl, d_ctl, d_seen, iters);
	run<1>("bulk_copy_one_cta", 1, h_stage, d_stage, h_ctl, d_ctl, d_seen, iters);
	run<2>("bulk_copy_one_cta_two_in_flight", 1, h_stage, d_stage, h_ctl, d_ctl, d_seen, iters);
	run<1>("bulk_multicast_c2", 2, h_stage, d_stage, h_ctl, d_ctl, d_seen, iters);
	run<1>("bulk_multicast_c4", 4, h_stage, d_stage, h_ctl, d_ctl, d_seen, iters);
	run<1>("bulk_multicast_c8", 8, h_stage, d_stage, h_ctl, d_ctl, d_seen, iters);
	run<2>("bulk_multicast_c8_two_in_flight", 8, h_stage, d_stage, h_ctl, d_ctl, d_seen, iters);
	std::printf(" \"bytes_per_round\": %d\n}\n", BYTES);
	return 0;
}
