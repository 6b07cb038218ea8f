// Litmus / stress test of the hand-off the resident real-time kernel uses inside its thread-block cluster
// (zen_b200/csrc/hpr_launch.cuh, hpr_rt_kernel, "the command goes to the other CTAs by a few remote stores"):
//
//   leader CTA:   every thread stores its part of the hop into the leader's OWN shared memory (plain st.shared),
//                 bar.sync, then ONE thread stores the command word into the shared memory of every other CTA
//                 (st.volatile through the cluster window)
//   other CTAs:   one thread spins on its LOCAL command word (ld.volatile.shared), bar.sync, then every thread
//                 reads the hop out of the leader's shared memory (ld.shared::cluster)
//
// PTX's memory model wants a release / acquire pair at cluster scope around the command word for the readers to
// be guaranteed the new hop.  The product path leaves the fence out of the common case (it costs every CTA
// ~0.5 us per hop, measured below) on the argument that shared memory is one physical copy with no cache in front
// of it and that bar.sync orders the leader's stores before the command store is issued.  This program is the
// evidence for that argument on the silicon it runs on: it replays the protocol with a fresh pattern per round,
// with the other warps of the leader hammering shared memory like the FFT that follows, and counts every word a
// reader saw that was not this round's.  It runs the plain and the fenced variant, and prints the time per round
// of both, i.e. what the fence would cost.
//
//   handoff_litmus [rounds] [cluster]      exit code 0: no stale word in either variant
#include <cooperative_groups.h>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

namespace cg = cooperative_groups;

constexpr int NT = 256;
constexpr int HOPW = 1024;  // words handed over per round, as at hop 1024

__device__ __forceinline__ unsigned pattern(unsigned round, unsigned i)
{
	unsigned x = round * 2654435761u + i * 40503u + 0x9e3779b9u;
	x ^= x >> 15;
	x *= 2246822519u;
	x ^= x >> 13;
	return x;
}

template <bool FENCED>
__global__ void __launch_bounds__(NT, 1) litmus_kernel(unsigned rounds, unsigned long long* stale, unsigned long long* first_bad, float* sink)
{
	__shared__ unsigned stash[2][HOPW];  // double-buffered by round parity, like S.hopbuf
	__shared__ unsigned cmd;
	__shared__ float noise[NT * 4];
	auto cluster = cg::this_cluster();
	const int C = (int)cluster.dim_blocks().x;
	const int rank = (int)cluster.block_rank();
	const int tid = threadIdx.x;
	if (tid == 0) cmd = 0;
	for (int i = tid; i < NT * 4; i += NT)
		noise[i] = (float)i;
	cluster.sync();
	unsigned long long bad = 0;
	float acc = 0.0f;
	for (unsigned r = 1; r <= rounds; ++r) {
		unsigned* mine = stash[r & 1];
		if (rank == 0) {
			for (int i = tid; i < HOPW; i += NT)
				mine[i] = pattern(r, (unsigned)i);
			__syncthreads();
			if (tid == NT - 1) {
				if (FENCED) asm volatile("fence.acq_rel.cluster;" ::: "memory");
				for (int q = 1; q < C; ++q)
					*reinterpret_cast<volatile unsigned*>(cluster.map_shared_rank(&cmd, q)) = r;
			}
			// the leader goes on with the hop: shared-memory traffic of its own while the others pull the stash
			for (int k = 0; k < 8; ++k) {
				acc += noise[(tid * 4 + k * 37) & (NT * 4 - 1)];
				noise[(tid * 4 + k) & (NT * 4 - 1)] = acc;
			}
		}
		else {
			if (tid == 0) {
				volatile unsigned* c = &cmd;
				while (*c != r) {
				}
				if (FENCED)
					asm volatile("fence.acq_rel.cluster;" ::: "memory");
				else
					asm volatile("" ::: "memory");
			}
			__syncthreads();
			const unsigned* theirs = cluster.map_shared_rank(mine, 0);
			// same access shape as the windowing loader: strided by the thread count, and once more in reverse order
			for (int i = tid; i < HOPW; i += NT) {
				const unsigned v = theirs[i];
				if (v != pattern(r, (unsigned)i)) {
					if (!bad) atomicMin(first_bad, (unsigned long long)r);
					++bad;
				}
			}
			for (int i = HOPW - 1 - tid; i >= 0; i -= NT) {
				const unsigned v = theirs[i];
				if (v != pattern(r, (unsigned)i)) ++bad;
			}
		}
		// the real hop has one cluster barrier (masked spectra have arrived) before anybody can see the next request
		cluster.sync();
	}
	if (bad) atomicAdd(stale, bad);
	if (acc == 123.456f) sink[0] = acc;
}

template <bool FENCED>
static int run(unsigned rounds, int C, double* us_per_round, unsigned long long* stale_out, unsigned long long* first_out)
{
	unsigned long long *d_stale, *d_first;
	float* d_sink;
	cudaMalloc(&d_stale, 8);
	cudaMalloc(&d_first, 8);
	cudaMalloc(&d_sink, 4);
	cudaMemset(d_stale, 0, 8);
	cudaMemset(d_first, 0xff, 8);
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
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0);
	cudaEventCreate(&e1);
	cudaEventRecord(e0);
	cudaError_t rc = cudaLaunchKernelEx(&cfg, litmus_kernel<FENCED>, rounds, d_stale, d_first, d_sink);
	cudaEventRecord(e1);
	if (rc == cudaSuccess) rc = cudaDeviceSynchronize();
	if (rc != cudaSuccess) {
		std::fprintf(stderr, "handoff_litmus: %s\n", cudaGetErrorString(rc));
		return 1;
	}
	float ms = 0;
	cudaEventElapsedTime(&ms, e0, e1);
	*us_per_round = 1e3 * ms / rounds;
	cudaMemcpy(stale_out, d_stale, 8, cudaMemcpyDeviceToHost);
	cudaMemcpy(first_out, d_first, 8, cudaMemcpyDeviceToHost);
	cudaFree(d_stale);
	cudaFree(d_first);
	cudaFree(d_sink);
	return 0;
}

int main(int argc, char** argv)
{
	const unsigned rounds = argc > 1 ? (unsigned)std::strtoul(argv[1], nullptr, 10) : 2000000u;
	const int C = argc > 2 ? std::atoi(argv[2]) : 8;
	double us_plain = 0, us_fenced = 0;
	unsigned long long stale_plain = 0, stale_fenced = 0, first_plain = 0, first_fenced = 0;
	if (run<false>(rounds, C, &us_plain, &stale_plain, &first_plain)) return 2;
	if (run<true>(rounds, C, &us_fenced, &stale_fenced, &first_fenced)) return 2;
	std::printf("{\"rounds\": %u, \"cluster\": %d, \"words_per_round\": %d, \"readers\": %d, "
	            "\"plain\": {\"stale_words\": %llu, \"us_per_round\": %.4f}, "
	            "\"fenced\": {\"stale_words\": %llu, \"us_per_round\": %.4f}}\n",
	            rounds, C, HOPW, C - 1, stale_plain, us_plain, stale_fenced, us_fenced);
	return (stale_plain || stale_fenced) ? 1 : 0;
}
