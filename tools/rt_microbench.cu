// Latency constants behind the resident real-time kernel's design (hpr_launch.cuh): how long a poll round over
// PCIe takes as a function of the number of 16-byte requests, what a cluster barrier costs, what forwarding a hop
// through distributed shared memory costs.  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a
// tools/rt_microbench.cu -o tools/_build/rt_microbench ; run on a B200; prints one JSON object.
#include <cooperative_groups.h>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { std::fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); std::exit(1); } } while (0)

__device__ __forceinline__ uint4 ld_sys(const uint4* p)
{
	uint4 v;
	asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
	return v;
}

// rounds of: `active` threads load `per` consecutive 16-byte groups each, then a CTA barrier
__global__ void poll_rounds(const uint4* host, int active, int per, int rounds, long long* out, unsigned* sink)
{
	const int tid = threadIdx.x;
	unsigned acc = 0;
	__syncthreads();
	long long t0 = clock64();
	for (int r = 0; r < rounds; ++r) {
		if (tid < active)
			for (int b = 0; b < per; ++b)
				acc += ld_sys(host + tid * per + b).w;
		__syncthreads();
	}
	long long t1 = clock64();
	if (tid == 0) out[0] = (t1 - t0) / rounds;
	sink[tid] = acc;
}

__global__ void cluster_sync_lat(int iters, long long* out)
{
	cg::cluster_group cl = cg::this_cluster();
	cl.sync();
	long long t0 = clock64();
	for (int i = 0; i < iters; ++i)
		cl.sync();
	long long t1 = clock64();
	if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (t1 - t0) / iters;
}

// rank 0 forwards 1026 floats to every other rank: scalar 4-byte stores (mode 0) or 16-byte stores (mode 1), then a cluster barrier
__global__ void dsmem_push(int mode, int iters, long long* out)
{
	__shared__ __align__(16) float buf[1368];
	cg::cluster_group cl = cg::this_cluster();
	const int C = cl.num_blocks(), rank = cl.block_rank(), tid = threadIdx.x;
	cl.sync();
	long long t0 = clock64();
	for (int i = 0; i < iters; ++i) {
		if (rank == 0 && tid < 342) {
			for (int r = 1; r < C; ++r) {
				float* rb = cl.map_shared_rank(buf, r);
				if (mode == 0) {
					rb[3 * tid] = (float)i;
					rb[3 * tid + 1] = (float)i;
					rb[3 * tid + 2] = (float)i;
				}
				else {
					reinterpret_cast<float4*>(rb)[tid] = make_float4((float)i, (float)i, (float)i, (float)i);
				}
			}
		}
		cl.sync();
	}
	long long t1 = clock64();
	if (tid == 0 && rank == 0) out[0] = (t1 - t0) / iters;
	if (buf[tid] == 12345.f) out[1] = 1;
}

// every rank > 0 pulls 1024 floats (float2 per thread pair) out of rank 0's shared memory after a cluster barrier
__global__ void dsmem_pull(int iters, long long* out)
{
	__shared__ __align__(16) float buf[1024];
	__shared__ __align__(16) float loc[1024];
	cg::cluster_group cl = cg::this_cluster();
	const int rank = cl.block_rank(), tid = threadIdx.x;
	for (int n = tid; n < 1024; n += blockDim.x)
		buf[n] = (float)n;
	cl.sync();
	const float2* src = reinterpret_cast<const float2*>(cl.map_shared_rank(buf, 0));
	long long t0 = clock64();
	for (int i = 0; i < iters; ++i) {
		for (int n = tid; n < 512; n += blockDim.x)
			reinterpret_cast<float2*>(loc)[n] = src[n];
		__syncthreads();
	}
	long long t1 = clock64();
	if (tid == 0 && rank == 1) out[0] = (t1 - t0) / iters;
	cl.sync();
	if (loc[tid] == 12345.f) out[1] = 1;
}

// flag ping-pong between rank 0 and rank 1 through remote shared-memory stores: cycles per one-way hop
__global__ void dsmem_pingpong(int iters, long long* out)
{
	__shared__ unsigned flag;
	cg::cluster_group cl = cg::this_cluster();
	const int rank = cl.block_rank();
	if (threadIdx.x == 0) flag = 0;
	cl.sync();
	if (threadIdx.x == 0 && rank < 2) {
		volatile unsigned* mine = &flag;
		volatile unsigned* other = cl.map_shared_rank(&flag, rank ^ 1);
		long long t0 = clock64();
		for (int i = 1; i <= iters; ++i) {
			if (rank == 0) {
				*other = (unsigned)i;
				while (*mine != (unsigned)i) {
				}
			}
			else {
				while (*mine != (unsigned)i) {
				}
				*other = (unsigned)i;
			}
		}
		long long t1 = clock64();
		if (rank == 0) out[0] = (t1 - t0) / (2 * iters);
	}
	cl.sync();
}

template <typename K, typename... A>
static void launch_cluster(K kern, int C, int nt, A... args)
{
	cudaLaunchConfig_t cfg = {};
	cfg.gridDim = dim3(C);
	cfg.blockDim = dim3(nt);
	cudaLaunchAttribute attr[1];
	attr[0].id = cudaLaunchAttributeClusterDimension;
	attr[0].val.clusterDim.x = C;
	attr[0].val.clusterDim.y = 1;
	attr[0].val.clusterDim.z = 1;
	cfg.attrs = attr;
	cfg.numAttrs = 1;
	CK(cudaLaunchKernelEx(&cfg, kern, args...));
	CK(cudaDeviceSynchronize());
}

int main()
{
	uint4* host = nullptr;
	CK(cudaHostAlloc((void**)&host, 64 * 1024, cudaHostAllocMapped));
	uint4* hdev = nullptr;
	CK(cudaHostGetDevicePointer((void**)&hdev, host, 0));
	long long* out = nullptr;
	CK(cudaMallocManaged(&out, 64));
	unsigned* sink = nullptr;
	CK(cudaMalloc(&sink, 4096 * 4));
	int khz = 0;
	CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0));
	std::printf("{\"note\": \"cycles (SM clock, nominal %d MHz)\",\n", khz / 1000);
	const int cfgs[][2] = {{1, 1}, {32, 1}, {86, 1}, {128, 1}, {342, 1}, {512, 1}, {1024, 1}, {86, 4}, {171, 2}, {43, 8}, {342, 4}};
	std::printf(" \"poll_round_cycles\": {");
	for (unsigned i = 0; i < sizeof(cfgs) / sizeof(cfgs[0]); ++i) {
		poll_rounds<<<1, 1024>>>(hdev, cfgs[i][0], cfgs[i][1], 200, out, sink);
		CK(cudaDeviceSynchronize());
		poll_rounds<<<1, 1024>>>(hdev, cfgs[i][0], cfgs[i][1], 500, out, sink);
		CK(cudaDeviceSynchronize());
		std::printf("%s\"%dthreads_x%d\": %lld", i ? ", " : "", cfgs[i][0], cfgs[i][1], out[0]);
	}
	std::printf("},\n \"cluster_sync_cycles\": {");
	for (int C = 1; C <= 8; C *= 2) {
		launch_cluster(cluster_sync_lat, C, 256, 1000, out);
		std::printf("%s\"C%d\": %lld", C > 1 ? ", " : "", C, out[0]);
	}
	std::printf("},\n \"dsmem_push_1026_floats_plus_cluster_sync_cycles\": {");
	for (int C = 2; C <= 8; C *= 2)
		for (int mode = 0; mode < 2; ++mode) {
			launch_cluster(dsmem_push, C, 512, mode, 300, out);
			std::printf("%s\"C%d_%s\": %lld", (C > 2 || mode) ? ", " : "", C, mode ? "16B" : "4B", out[0]);
		}
	std::printf("},\n \"dsmem_pull_1024_floats_cycles\": {");
	for (int C = 2; C <= 8; C *= 2) {
		launch_cluster(dsmem_pull, C, 256, 300, out);
		std::printf("%s\"C%d\": %lld", C > 2 ? ", " : "", C, out[0]);
	}
	launch_cluster(dsmem_pingpong, 2, 32, 2000, out);
	std::printf("},\n \"dsmem_flag_one_way_cycles\": %lld\n}\n", out[0]);
	return 0;
}
