// Kernels + C ABI of the streaming / batched / offline HPR path.
// Reference: libzen/hps.h, libzen/hps.cu (HPR<GPU>, HPRRealtime<GPU>, HPRIOffline<GPU>).
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <new>
#include <chrono>
#include <vector>

// The host side of the tagged 16-byte groups (RtCtrl, hpr_launch.cuh) has an SSE2 fast path and a portable one
// (GCC / clang vector extensions: one aligned 16-byte load / store per group on x86-64 and aarch64 alike);
// -DZEN_PORTABLE_GROUPS selects the portable one on x86-64 too (tests/test_capi_host.py builds and checks both).
#if defined(__x86_64__) && !defined(ZEN_PORTABLE_GROUPS)
#define ZEN_GROUPS_SSE 1
#include <xmmintrin.h>
#include <emmintrin.h>
static inline void zen_host_store_fence() { _mm_sfence(); }
#else
#define ZEN_GROUPS_SSE 0
#include <atomic>
static inline void zen_host_store_fence() { std::atomic_thread_fence(std::memory_order_seq_cst); }
#endif

#include "hpr_launch.cuh"

using namespace zen_b200;

__global__ void copy_hop_kernel(const float* __restrict__ src, float* __restrict__ dst, int n)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n)
		dst[i] = src[i];
}

// intermediate = percussive + residual of pass 1, un-lagged (hps.cu:154-176).
// The reference reads pass-2 hops past the truncated size from the same
// allocation, where the un-shifted tail still sits; restated here.
__global__ void offline_intermediate_kernel(const float* __restrict__ p1, const float* __restrict__ r1,
                                            float* __restrict__ inter, long padded1, long shift1, long padded2)
{
	long j = (long)blockIdx.x * blockDim.x + threadIdx.x;
	if (j >= padded2)
		return;
	long src = (j + shift1 < padded1) ? j + shift1 : j;
	inter[j] = src < padded1 ? p1[src] + r1[src] : 0.0f;
}

// ------------------------------------------------------------------- host ---

namespace {

struct Plan {
	HprDev dev;
	zen_geometry geom;
	int nfft;
	int causality;
	float* d_window = nullptr;
	float2* d_tw = nullptr;
	float2* d_twr = nullptr;
	size_t smem_bytes = 0;
};

// rounding boundary of RN(x / y) >= c: the midpoint between c and its predecessor (hpr_core.cuh: thr_ratio_ge)
RatioRule make_ratio_rule(float c)
{
	RatioRule r;
	r.m = 0.0;
	r.even = 0;
	r.trivial = 0;
	if (c != c) {
		r.trivial = -1;
		return r;
	}
	if (!(c > 0.0f)) {
		r.trivial = 1;
		return r;
	}
	if (std::isinf(c)) {
		r.trivial = -1;
		return r;
	}
	uint32_t bits;
	std::memcpy(&bits, &c, sizeof(bits));
	uint32_t pb = bits - 1u;
	float pred;
	std::memcpy(&pred, &pb, sizeof(pred));
	r.m = 0.5 * ((double)pred + (double)c);
	r.even = (bits & 1u) == 0u;
	return r;
}

int build_plan(Plan& pl, float fs, int hop, float beta, unsigned flags, int causality, int copy_bord, bool sse, bool soft)
{
	if (hop < 32 || hop > 4096 || !is_pow2(hop))
		return ZEN_ERR_UNSUPPORTED;
	if (zen_hpr_geometry(fs, hop, causality == ZEN_TIME_CAUSAL, &pl.geom) != ZEN_OK)
		return ZEN_ERR_ARG;
	const zen_geometry& g = pl.geom;
	// the four filter constructors of HPR<B> throw when the filter is longer than its axis
	// (mfilt.h:80-87, box.h:71-78)
	if (g.stft_width < 1 || g.l_harm > g.stft_width || g.l_perc > g.nfft)
		return ZEN_ERR_GEOMETRY;
	pl.nfft = g.nfft;
	pl.causality = causality;
	HprDev& d = pl.dev;
	std::memset(&d, 0, sizeof(d));
	const int M = g.nfft / 2;
	d.hop = hop;
	d.W = g.stft_width;
	d.lag = g.lag;
	d.Lp = odd_len(g.l_perc);
	d.midp = d.Lp / 2;
	d.Kp = sliding_K_for(d.Lp);
	d.Cp = thread_window_capacity(d.Lp);
	if ((d.Kp == 0 && d.Cp == 0) || d.Lp > M)
		return ZEN_ERR_UNSUPPORTED;
	d.copy_bord = copy_bord ? 1 : 0;
	d.out_flags = (int)(flags & 7u);
	d.soft = soft ? 1 : 0;
	d.decide = std::getenv("ZEN_B200_NO_DECIDE") ? 0 : 1;  // debugging aid: force the median-selection path
	d.sse = sse ? 1 : 0;
	d.power = (float)(int)beta;
	d.beta = beta;
	d.beta_h = beta - ZEN_EPS;
	d.rule_p = make_ratio_rule(d.beta);
	d.rule_h = make_ratio_rule(d.beta_h);
	d.cola = g.cola_factor;
	d.lh1 = (float)g.l_harm + 1.0f;
	d.lp1 = (float)g.l_perc + 1.0f;
	d.inv_lp = 1.0f / (float)d.Lp;
	// time-axis taps of the consumed row r = W - lag, as ages behind the newest frame
	// (SURVEY.md section 8(a'); mfilt.h:111-189)
	{
		const int Lh = odd_len(g.l_harm), mid = Lh / 2, W = g.stft_width, r = W - g.lag;
		std::vector<int> rows;
		const bool wrap = copy_bord || sse;  // BoxFilterGPU always wrap-pads (box.h:194-205)
		if (wrap) {
			for (int t = 0; t < Lh; ++t)
				rows.push_back(((r - mid + t) % W + W) % W);
		}
		else if (causality == ZEN_TIME_CAUSAL) {
			if (r >= Lh && r < W)
				for (int t = 0; t < Lh; ++t)
					rows.push_back(r - Lh + t);
		}
		else {
			if (r >= mid && r < mid + W - Lh)
				for (int t = 0; t < Lh; ++t)
					rows.push_back(r - mid + t);
		}
		if ((int)rows.size() > ZEN_MAX_TAPS)
			return ZEN_ERR_UNSUPPORTED;
		d.n_taps = (int)rows.size();
		for (size_t t = 0; t < rows.size(); ++t)
			d.tap_age[t] = (short)(W - 1 - rows[t]);
	}
	// tables
	std::vector<float> win(g.nwin);
	zen_window(ZEN_WIN_SQRT_VON_HANN, g.nwin, win.data());
	const int ntw = std::max(1, fft_twiddle_count_rt(M));
	std::vector<float2> tw(ntw), twr(M / 2 + 1);
	const double two_pi = 6.283185307179586476925286766559;
	fft_fill_twiddles(M, tw.data());
	for (int k = 0; k <= M / 2; ++k)
		twr[k] = make_float2((float)std::cos(two_pi * k / g.nfft), (float)-std::sin(two_pi * k / g.nfft));
	ZEN_CUDA_CHECK(cudaMalloc(&pl.d_window, sizeof(float) * g.nwin));
	ZEN_CUDA_CHECK(cudaMalloc(&pl.d_tw, sizeof(float2) * ntw));
	ZEN_CUDA_CHECK(cudaMalloc(&pl.d_twr, sizeof(float2) * (M / 2 + 1)));
	ZEN_CUDA_CHECK(cudaMemcpy(pl.d_window, win.data(), sizeof(float) * g.nwin, cudaMemcpyHostToDevice));
	ZEN_CUDA_CHECK(cudaMemcpy(pl.d_tw, tw.data(), sizeof(float2) * ntw, cudaMemcpyHostToDevice));
	ZEN_CUDA_CHECK(cudaMemcpy(pl.d_twr, twr.data(), sizeof(float2) * (M / 2 + 1), cudaMemcpyHostToDevice));
	d.window = pl.d_window;
	d.tw = pl.d_tw;
	d.twr = pl.d_twr;
	return ZEN_OK;
}

void free_plan(Plan& pl)
{
	cudaFree(pl.d_window);
	cudaFree(pl.d_tw);
	cudaFree(pl.d_twr);
	pl.d_window = nullptr;
	pl.d_tw = nullptr;
	pl.d_twr = nullptr;
}

size_t tile_scratch_floats(const Plan& pl)
{
	const int M = pl.nfft / 2;
	size_t ring = (size_t)pl.dev.W * (M + 1);
	ring += ring & 1;
	size_t xr = pl.dev.lag > 1 ? 2 * (size_t)pl.dev.lag * (M + 1) : 0;
	size_t tails = 3 * (size_t)pl.dev.hop;
	const size_t general = ring + xr + tails;
	const size_t fast = (size_t)pl.dev.W * (M + 4) + tails;   // hpr_tile_fast_kernel: 16-byte aligned ring rows
	return (std::max(general, fast) + 3) & ~(size_t)3;
}

int resident_ctas_for(const Plan& pl)
{
	switch (pl.nfft) {
	case 128: return tile_resident_ctas<128>(pl.dev);
	case 256: return tile_resident_ctas<256>(pl.dev);
	case 512: return tile_resident_ctas<512>(pl.dev);
	case 1024: return tile_resident_ctas<1024>(pl.dev);
	case 2048: return tile_resident_ctas<2048>(pl.dev);
	case 4096: return tile_resident_ctas<4096>(pl.dev);
	case 8192: return tile_resident_ctas<8192>(pl.dev);
	case 16384: return tile_resident_ctas<16384>(pl.dev);
	}
	return 0;
}

// peaks (optional): per output a device array of n_streams floats, zeroed by the caller, that receives max |emitted
// sample| per stream (only where the fast path applies: tile_peaks_supported)
bool tile_peaks_supported(const Plan& pl) { return hpr_fast_supported(pl.dev) && !std::getenv("ZEN_B200_NO_FAST"); }

int dispatch_tile(const Plan& pl, const float* in, long in_stride, float* oh, float* op, float* orr, long out_stride,
                  int n_streams, long n_hops, int tile_hops, float* scratch, int* work_counter, int resident, cudaStream_t s,
                  float* const* peaks = nullptr)
{
	TileArgs a{pl.dev, in, in_stride, oh, op, orr, out_stride, n_streams, n_hops, tile_hops, scratch, tile_scratch_floats(pl),
	           work_counter, resident, s};
	a.force_general = std::getenv("ZEN_B200_NO_FAST") ? 1 : 0;  // debugging aid / A-B test: the general kernel everywhere
	// Fewer streams than resident CTAs (e.g. a 4096-stream job sharded over eight GPUs): choose_tile_hops cuts the streams
	// into many short tiles so that the queue balances, and every tile re-analyses W hops in front of it.  Only the END of
	// the queue needs short items: the first tiles of every stream are four times as long, the last ones - about four per
	// CTA, queued last (tile-major order) - keep the short length.
	if (resident > 0 && n_streams < resident && tile_hops < n_hops && !std::getenv("ZEN_B200_NO_STRETCH")) {
		const long n_small = (4L * resident + n_streams - 1) / n_streams;
		const long hops_small = std::min<long>(n_hops, n_small * tile_hops);
		const long big = 4L * tile_hops;
		a.big_hops = (int)big;
		a.n_big = (int)((n_hops - hops_small) / big);
	}
	if (peaks)
		for (int o = 0; o < 3; ++o)
			a.peaks[o] = reinterpret_cast<unsigned*>(peaks[o]);
	switch (pl.nfft) {
	case 128: return launch_tile_impl<128>(a);
	case 256: return launch_tile_impl<256>(a);
	case 512: return launch_tile_impl<512>(a);
	case 1024: return launch_tile_impl<1024>(a);
	case 2048: return launch_tile_impl<2048>(a);
	case 4096: return launch_tile_impl<4096>(a);
	case 8192: return launch_tile_impl<8192>(a);
	case 16384: return launch_tile_impl<16384>(a);
	}
	return ZEN_ERR_UNSUPPORTED;
}

}  // namespace

// ------------------------------------------------------ streaming object ---

struct zen_hpr {
	Plan plan;
	float fs, beta;
	unsigned flags;
	int hop, causality, copy_bord;
	bool sse = false, soft = false;
	bool plan_dirty = false;
	bool state_external = false;  // input / *_out owned by the caller (zen_hpr_bind_state)
	long iter = 0;
	cudaStream_t stream = nullptr;
	// state
	float* d_input = nullptr;   // nwin
	float* d_ola[3] = {nullptr, nullptr, nullptr};  // nwin each (harmonic_out, percussive_out, residual_out)
	float* d_mag_ring = nullptr;
	float2* d_x_ring = nullptr;
	// persistent real-time session (zen_hpr_realtime_begin)
	bool rt_mode = false;      // hops are served by the resident kernel
	bool rt_running = false;   // a resident kernel has been launched and not yet collected
	RtCtrl* rt_ctrl = nullptr;
	RtCtrl* rt_ctrl_dev = nullptr;
	cudaStream_t rt_stream = nullptr;
	int* d_iter = nullptr;
	unsigned rt_seq = 0;
	bool rt_args_valid = false;  // the resident kernel holds the pointers of the previous call
	unsigned long long rt_idle_ns = 250ull * 1000 * 1000;
	// tagged staging buffers of the resident kernel (mapped pinned host memory, see RtCtrl)
	uint4* rt_stage_in = nullptr;
	uint4* rt_stage_in_dev = nullptr;
	uint4* rt_stage_out[3] = {nullptr, nullptr, nullptr};
	uint4* rt_stage_out_dev[3] = {nullptr, nullptr, nullptr};
	int rt_groups = 0;               // ceil(hop / 3)
	bool rt_push = true;             // ZEN_B200_RT_PUSH=0: the kernel pulls the hop / writes the outputs itself
	const float* rt_in_host = nullptr;          // host alias of the current `in` pointer (null: device memory)
	float* rt_out_host[3] = {nullptr, nullptr, nullptr};
	bool rt_out_all_host = false;    // every non-null destination of the current call is host memory
	bool rt_stamps = false;          // ZEN_B200_RT_STAMPS=1: the kernel records its phase boundaries (diagnostics)
	bool rt_fenced = true;           // ZEN_B200_RT_FENCED=0 drops the cluster-scope fences of the hop hand-off (RtArgs::fenced): same latency either way (profiles/r02_rt_latency_fenced.txt)
	int rt_cluster = 8;              // ZEN_B200_RT_CLUSTER: CTAs serving the stream when the plan allows the split hop (8: p50 10.7 us / p99 11.2 us at hop 1024 against 11.0 / 14.1 with 4, profiles/r02_rt_latency.json)
	// A process_next_hop without destinations is only SUBMITTED (like the reference's, which queues its kernels and lets
	// copy_* wait, hps.cu:341-363): its outputs land in the tagged staging buffers and the copy_* that follows unpacks them
	// on the host - the two-call sequence of zen/fakert.h:229-230 costs one round trip to the device, not two.
	struct RtPending {
		bool active = false;
		unsigned op = 0, opw = 0, tag = 0, target = 0;
		const float* push_src = nullptr;
		bool tagged[3] = {false, false, false};
	} rt_pend;
	unsigned rt_last_tag = 0;                          // tag of the last PROCESS whose outputs sit in the staging buffers
	bool rt_last_tagged[3] = {false, false, false};
	const float* rt_copy_ptr[3] = {nullptr, nullptr, nullptr};  // copy_* destination last seen per output and its host alias
	float* rt_copy_host[3] = {nullptr, nullptr, nullptr};
	// ZEN_B200_RT_TRACE=1: host-side split of the per-hop time (publish / wait for the device / unpack), printed on destroy
	bool rt_trace = false;
	double rt_trace_ns[3] = {0.0, 0.0, 0.0};
	long rt_trace_n = 0;
};

namespace {

int dispatch_hop(zen_hpr* h, const float* in_hop, float* eh, float* ep, float* er)
{
	HopArgs a;
	a.dev = h->plan.dev;
	a.st.mag_ring = h->d_mag_ring;
	a.st.x_ring = h->d_x_ring;
	a.st.xdepth = h->plan.dev.W;
	a.st.tail[0] = a.st.tail[1] = a.st.tail[2] = nullptr;
	a.iter = h->iter;
	a.input = h->d_input;
	a.in_hop = in_hop;
	for (int o = 0; o < 3; ++o)
		a.ola[o] = h->d_ola[o];
	a.ext[0] = eh;
	a.ext[1] = ep;
	a.ext[2] = er;
	a.stream = h->stream;
	int rc = ZEN_ERR_UNSUPPORTED;
	switch (h->plan.nfft) {
	case 128: rc = launch_hop_impl<128>(a); break;
	case 256: rc = launch_hop_impl<256>(a); break;
	case 512: rc = launch_hop_impl<512>(a); break;
	case 1024: rc = launch_hop_impl<1024>(a); break;
	case 2048: rc = launch_hop_impl<2048>(a); break;
	case 4096: rc = launch_hop_impl<4096>(a); break;
	case 8192: rc = launch_hop_impl<8192>(a); break;
	case 16384: rc = launch_hop_impl<16384>(a); break;
	}
	if (rc == ZEN_OK)
		h->iter++;
	return rc;
}

int rebuild_plan_fwd(zen_hpr* h);

// ---- persistent real-time session -------------------------------------------

// host alias of a device-visible pointer into mapped pinned host memory, or null for device memory
void* rt_host_alias(const void* p)
{
	if (!p) return nullptr;
	cudaPointerAttributes at;
	if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
		cudaGetLastError();
		return nullptr;
	}
	return at.type == cudaMemoryTypeHost ? at.hostPointer : nullptr;
}

int rt_launch(zen_hpr* h)
{
	if (!h->rt_ctrl) {
		ZEN_CUDA_CHECK(cudaHostAlloc((void**)&h->rt_ctrl, sizeof(RtCtrl), cudaHostAllocMapped | cudaHostAllocPortable));
		std::memset((void*)h->rt_ctrl, 0, sizeof(RtCtrl));
		ZEN_CUDA_CHECK(cudaHostGetDevicePointer((void**)&h->rt_ctrl_dev, (void*)h->rt_ctrl, 0));
		ZEN_CUDA_CHECK(cudaStreamCreateWithFlags(&h->rt_stream, cudaStreamNonBlocking));
		ZEN_CUDA_CHECK(cudaMalloc(&h->d_iter, sizeof(int)));
		if (const char* e = std::getenv("ZEN_B200_RT_IDLE_MS")) {
			long ms = std::atol(e);
			if (ms > 0) h->rt_idle_ns = (unsigned long long)ms * 1000ull * 1000ull;
		}
		if (const char* e = std::getenv("ZEN_B200_RT_PUSH"))
			h->rt_push = std::atoi(e) != 0;
		if (const char* e = std::getenv("ZEN_B200_RT_STAMPS"))
			h->rt_stamps = std::atoi(e) != 0;
		if (const char* e = std::getenv("ZEN_B200_RT_TRACE"))
			h->rt_trace = std::atoi(e) != 0;
		if (const char* e = std::getenv("ZEN_B200_RT_CLUSTER")) {
			const int c = std::atoi(e);
			if (c == 1 || c == 2 || c == 4 || c == 8) h->rt_cluster = c;
		}
		if (const char* e = std::getenv("ZEN_B200_RT_FENCED"))
			h->rt_fenced = std::atoi(e) != 0;
	}
	const int groups = (h->hop + 2) / 3;
	if (!h->rt_stage_in || h->rt_groups != groups) {
		// one allocation: request / input groups, then the three output group arrays
		if (h->rt_stage_in) cudaFreeHost(h->rt_stage_in);
		h->rt_stage_in = nullptr;
		const size_t per = ((size_t)groups * sizeof(uint4) + 127) & ~(size_t)127;
		unsigned char* base = nullptr;
		ZEN_CUDA_CHECK(cudaHostAlloc((void**)&base, 4 * per, cudaHostAllocMapped | cudaHostAllocPortable));
		std::memset(base, 0, 4 * per);
		unsigned char* base_dev = nullptr;
		ZEN_CUDA_CHECK(cudaHostGetDevicePointer((void**)&base_dev, base, 0));
		h->rt_stage_in = reinterpret_cast<uint4*>(base);
		h->rt_stage_in_dev = reinterpret_cast<uint4*>(base_dev);
		for (int o = 0; o < 3; ++o) {
			h->rt_stage_out[o] = reinterpret_cast<uint4*>(base + (size_t)(o + 1) * per);
			h->rt_stage_out_dev[o] = reinterpret_cast<uint4*>(base_dev + (size_t)(o + 1) * per);
		}
		h->rt_groups = groups;
		// tags of the sequence number already served, so that a fresh kernel does not take stale groups for a request
		for (int g = 0; g < groups; ++g)
			h->rt_stage_in[g].w = h->rt_seq << 8;  // (samples zero: hash zero)
	}
	ZEN_CUDA_CHECK(cudaStreamSynchronize(h->stream));
	int it = (int)h->iter;
	ZEN_CUDA_CHECK(cudaMemcpy(h->d_iter, &it, sizeof(int), cudaMemcpyHostToDevice));
	RtArgs a;
	a.dev = h->plan.dev;
	a.ctrl = h->rt_ctrl_dev;
	a.stage_in = h->rt_stage_in_dev;
	for (int o = 0; o < 3; ++o)
		a.stage_out[o] = h->rt_stage_out_dev[o];
	a.mag_ring = h->d_mag_ring;
	a.input = h->d_input;
	for (int o = 0; o < 3; ++o)
		a.ola[o] = h->d_ola[o];
	a.iter = h->d_iter;
	a.seq0 = h->rt_seq;
	a.idle_ns = h->rt_idle_ns;
	a.fenced = h->rt_fenced ? 1 : 0;
	a.stream = h->rt_stream;
	h->rt_args_valid = false;
	h->rt_ctrl->seq_out = h->rt_seq;
	h->rt_ctrl->exit_reason = 0;
	h->rt_ctrl->alive = 1;
	zen_host_store_fence();
	int rc = ZEN_ERR_UNSUPPORTED;
	const size_t limit = 227 * 1024;
	// The hop is split over a thread-block cluster (hpr_split_analyse) for the default real-time path: hard mask decided
	// by counting, copy-border, one CTA per output at most, everything resident in shared memory.
	const HprDev& d = a.dev;
	int n_out = 0;
	for (int o = 0; o < 3; ++o)
		n_out += (d.out_flags >> o) & 1;
	const bool can_split = h->rt_cluster > 1 && d.decide && !d.sse && !d.soft && d.copy_bord && d.lag == 1 && d.Lp >= 8
	                       && n_out >= 1 && n_out <= h->rt_cluster && h->plan.nfft >= 256;
#define ZEN_RT_CASE(N)                                                            \
	case N:                                                                       \
		a.cluster = (can_split && rt_smem_bytes<N>(a.dev, 3, h->rt_cluster) <= limit) ? h->rt_cluster : 1; \
		a.state_in_smem = rt_smem_bytes<N>(a.dev, 3, a.cluster) <= limit ? 3 : (rt_smem_bytes<N>(a.dev, 1, 1) <= limit ? 1 : (rt_smem_bytes<N>(a.dev, 2, 1) <= limit ? 2 : 0)); \
		rc = launch_rt_impl<N>(a);                                                \
		if (rc != ZEN_OK && a.cluster > 1) { /* no room for a cluster (partitioned GPU, SMs taken): one CTA serves the stream */ \
			cudaGetLastError();                                                   \
			a.cluster = 1;                                                        \
			a.state_in_smem = rt_smem_bytes<N>(a.dev, 3, 1) <= limit ? 3 : (rt_smem_bytes<N>(a.dev, 1, 1) <= limit ? 1 : (rt_smem_bytes<N>(a.dev, 2, 1) <= limit ? 2 : 0)); \
			rc = launch_rt_impl<N>(a);                                            \
		}                                                                         \
		break;
	switch (h->plan.nfft) {
		ZEN_RT_CASE(128)
		ZEN_RT_CASE(256)
		ZEN_RT_CASE(512)
		ZEN_RT_CASE(1024)
		ZEN_RT_CASE(2048)
		ZEN_RT_CASE(4096)
		ZEN_RT_CASE(8192)
		ZEN_RT_CASE(16384)
	}
#undef ZEN_RT_CASE
	if (rc == ZEN_OK)
		h->rt_running = true;
	else
		h->rt_ctrl->alive = 0;
	return rc;
}

// the resident kernel has left (stop or idle time-out): fetch the frame counter
int rt_collect(zen_hpr* h)
{
	if (!h->rt_running)
		return ZEN_OK;
	ZEN_CUDA_CHECK(cudaStreamSynchronize(h->rt_stream));
	int it = 0;
	ZEN_CUDA_CHECK(cudaMemcpy(&it, h->d_iter, sizeof(int), cudaMemcpyDeviceToHost));
	h->iter = it;
	h->rt_running = false;
	return ZEN_OK;
}

#if ZEN_GROUPS_SSE
// ---- tagged 16-byte groups {x[3g], x[3g+1], x[3g+2], tag ^ zen_group_hash(x)} (RtCtrl, hpr_launch.cuh), host side ----
inline __m128i rotl32(__m128i v, int n) { return _mm_or_si128(_mm_slli_epi32(v, n), _mm_srli_epi32(v, 32 - n)); }

// zen_group_hash of the three sample lanes of one group, in every lane
inline __m128i group_hash(__m128i g)
{
	return _mm_xor_si128(_mm_shuffle_epi32(g, _MM_SHUFFLE(0, 0, 0, 0)),
	                     _mm_xor_si128(_mm_shuffle_epi32(rotl32(g, 11), _MM_SHUFFLE(1, 1, 1, 1)), _mm_shuffle_epi32(rotl32(g, 22), _MM_SHUFFLE(2, 2, 2, 2))));
}

// one group: the three samples of `x` (lane 3 ignored) and the hashed tag, one aligned 16-byte store
inline void store_group(uint4* dst, __m128 x, __m128i tag_all)
{
	const __m128i keep = _mm_set_epi32(0, -1, -1, -1);
	const __m128i g = _mm_and_si128(_mm_castps_si128(x), keep);
	const __m128i w = _mm_andnot_si128(keep, _mm_xor_si128(group_hash(g), tag_all));
	_mm_store_si128(reinterpret_cast<__m128i*>(dst), _mm_or_si128(g, w));
}

// the plain tag of a group as read (lane 3 XOR the hash of lanes 0-2); a half-written group gives neither the old nor the new tag
inline unsigned group_tag(__m128i g)
{
	return (unsigned)_mm_cvtsi128_si32(_mm_shuffle_epi32(_mm_xor_si128(g, group_hash(g)), 0xFF));
}

// the line hash of twelve samples as they lie in memory (zen_group_key, hpr_core.cuh): V[l] = y[l] ^ rotl(y[4+l], 11) ^ rotl(y[8+l], 22)
inline __m128i line_hash(__m128 s0, __m128 s1, __m128 s2)
{
	return _mm_xor_si128(_mm_castps_si128(s0), _mm_xor_si128(rotl32(_mm_castps_si128(s1), 11), rotl32(_mm_castps_si128(s2), 22)));
}

// pack `hop` samples into ceil(hop / 3) groups
void rt_pack_groups(const float* src, int hop, unsigned tag, uint4* st)
{
	const int groups = (hop + 2) / 3;
	const int line_groups = 4 * (hop / 12);  // groups of whole 64-byte lines: tag words XOR the line hash
	const __m128i tagv = _mm_set1_epi32((int)tag);
	int g = 0;
	// a line (four groups, twelve samples) per step: three loads, four aligned 16-byte stores
	for (; g < line_groups; g += 4) {
		const __m128 s0 = _mm_loadu_ps(src + 3 * g), s1 = _mm_loadu_ps(src + 3 * g + 4), s2 = _mm_loadu_ps(src + 3 * g + 8);
		const __m128 W = _mm_castsi128_ps(_mm_xor_si128(tagv, line_hash(s0, s1, s2)));
		// {a b c ?} + tag word k of W -> {a b c w}
		auto with_tag = [&](__m128 grp, __m128 wk) { return _mm_shuffle_ps(grp, _mm_shuffle_ps(grp, wk, _MM_SHUFFLE(0, 0, 2, 2)), _MM_SHUFFLE(2, 0, 1, 0)); };
		const __m128 t1 = _mm_shuffle_ps(s0, s1, _MM_SHUFFLE(0, 0, 3, 3));   // x3 x3 x4 x4
		_mm_store_ps(reinterpret_cast<float*>(st + g), with_tag(s0, _mm_shuffle_ps(W, W, _MM_SHUFFLE(0, 0, 0, 0))));
		_mm_store_ps(reinterpret_cast<float*>(st + g + 1), with_tag(_mm_shuffle_ps(t1, s1, _MM_SHUFFLE(3, 1, 2, 0)), _mm_shuffle_ps(W, W, _MM_SHUFFLE(1, 1, 1, 1))));
		_mm_store_ps(reinterpret_cast<float*>(st + g + 2), with_tag(_mm_shuffle_ps(s1, s2, _MM_SHUFFLE(1, 0, 3, 2)), _mm_shuffle_ps(W, W, _MM_SHUFFLE(2, 2, 2, 2))));
		_mm_store_ps(reinterpret_cast<float*>(st + g + 3), with_tag(_mm_shuffle_ps(s2, s2, _MM_SHUFFLE(3, 3, 2, 1)), _mm_shuffle_ps(W, W, _MM_SHUFFLE(3, 3, 3, 3))));
	}
	// behind the last whole line: per-group hash
	for (; g < groups; ++g) {
		float x0 = src[3 * g], x1 = 3 * g + 1 < hop ? src[3 * g + 1] : 0.0f, x2 = 3 * g + 2 < hop ? src[3 * g + 2] : 0.0f;
		store_group(st + g, _mm_set_ps(0.0f, x2, x1, x0), tagv);
	}
}

// Unpack the groups that carry `tag` into dst, starting at group g (a multiple of four inside the whole lines;
// updated); true when the whole hop is out, false at the first line / group that is still old or caught half-written.
bool rt_unpack_groups(const uint4* st, int hop, unsigned tag, float* dst, int& g)
{
	const int groups = (hop + 2) / 3;
	const int line_groups = 4 * (hop / 12);
	const __m128i tagv = _mm_set1_epi32((int)tag);
	for (; g < line_groups; g += 4) {
		const __m128i a = _mm_load_si128(reinterpret_cast<const __m128i*>(st + g)), b = _mm_load_si128(reinterpret_cast<const __m128i*>(st + g + 1));
		const __m128i c = _mm_load_si128(reinterpret_cast<const __m128i*>(st + g + 2)), d = _mm_load_si128(reinterpret_cast<const __m128i*>(st + g + 3));
		const __m128 fa = _mm_castsi128_ps(a), fb = _mm_castsi128_ps(b), fc = _mm_castsi128_ps(c), fd = _mm_castsi128_ps(d);
		const __m128 t0 = _mm_shuffle_ps(fa, fb, _MM_SHUFFLE(0, 0, 2, 2));   // a2 a2 b0 b0
		const __m128 t2 = _mm_shuffle_ps(fc, fd, _MM_SHUFFLE(0, 0, 2, 2));   // c2 c2 d0 d0
		const __m128 s0 = _mm_shuffle_ps(fa, t0, _MM_SHUFFLE(2, 0, 1, 0));   // a0 a1 a2 b0
		const __m128 s1 = _mm_shuffle_ps(fb, fc, _MM_SHUFFLE(1, 0, 2, 1));   // b1 b2 c0 c1
		const __m128 s2 = _mm_shuffle_ps(t2, fd, _MM_SHUFFLE(2, 1, 2, 0));   // c2 d0 d1 d2
		const __m128i w = _mm_unpackhi_epi64(_mm_unpackhi_epi32(a, b), _mm_unpackhi_epi32(c, d));  // a3 b3 c3 d3
		if (_mm_movemask_epi8(_mm_cmpeq_epi32(_mm_xor_si128(w, line_hash(s0, s1, s2)), tagv)) != 0xffff)
			return false;  // the line is taken as a whole or not at all
		_mm_storeu_ps(dst + 3 * g, s0);
		_mm_storeu_ps(dst + 3 * g + 4, s1);
		_mm_storeu_ps(dst + 3 * g + 8, s2);
	}
	for (; g < groups; ++g) {
		const __m128i v = _mm_load_si128(reinterpret_cast<const __m128i*>(st + g));
		if (group_tag(v) != tag)
			return false;
		alignas(16) float t[4];
		_mm_store_ps(t, _mm_castsi128_ps(v));
		for (int j = 0; j < 3 && 3 * g + j < hop; ++j)
			dst[3 * g + j] = t[j];
	}
	return true;
}

// the last group of the hop carries `tag` (validated with its line when it belongs to one)
bool rt_last_group_ready(const uint4* st, int hop, unsigned tag)
{
	const int groups = (hop + 2) / 3, line_groups = 4 * (hop / 12);
	if (groups > line_groups)
		return group_tag(_mm_load_si128(reinterpret_cast<const __m128i*>(st + groups - 1))) == tag;
	alignas(16) float scratch[12];
	int g = groups - 4;
	const uint4* line = st + g;
	int g0 = 0;
	return rt_unpack_groups(line, 12, tag, scratch, g0);
}

// publish one request: tag every group of the staging buffer; with `src` the groups carry the hop
void rt_publish(zen_hpr* h, unsigned tag, const float* src)
{
	uint4* st = h->rt_stage_in;
	if (src)
		rt_pack_groups(src, h->hop, tag, st);
	else
		for (int g = 0; g < h->rt_groups; ++g)  // no samples: {0, 0, 0, tag} (the hash of zeros is zero)
			_mm_store_si128(reinterpret_cast<__m128i*>(st + g), _mm_set_epi32((int)tag, 0, 0, 0));
}

#else  // portable: the same groups, bit for bit, without intrinsics
typedef unsigned zen_v4u __attribute__((vector_size(16), aligned(16)));
inline unsigned rotl32s(unsigned v, int n) { return (v << n) | (v >> (32 - n)); }
inline unsigned fbits(float f)
{
	unsigned u;
	std::memcpy(&u, &f, sizeof(u));
	return u;
}
inline float bitsf(unsigned u)
{
	float f;
	std::memcpy(&f, &u, sizeof(f));
	return f;
}
// one aligned 16-byte store / load (str q / ldr q, movdqa): a group never becomes visible in two pieces
inline void st16(uint4* dst, unsigned a, unsigned b, unsigned c, unsigned d)
{
	const zen_v4u v = {a, b, c, d};
	*reinterpret_cast<volatile zen_v4u*>(dst) = v;
}
inline void ld16(const uint4* src, unsigned (&o)[4])
{
	const zen_v4u v = *reinterpret_cast<const volatile zen_v4u*>(src);
	o[0] = v[0];
	o[1] = v[1];
	o[2] = v[2];
	o[3] = v[3];
}

void rt_pack_groups(const float* src, int hop, unsigned tag, uint4* st)
{
	const int groups = (hop + 2) / 3;
	const int line_groups = 4 * (hop / 12);
	int g = 0;
	for (; g < line_groups; g += 4) {
		unsigned y[12], V[4];
		for (int l = 0; l < 12; ++l)
			y[l] = fbits(src[3 * g + l]);
		for (int l = 0; l < 4; ++l)
			V[l] = y[l] ^ rotl32s(y[4 + l], 11) ^ rotl32s(y[8 + l], 22);  // zen_group_key's line hash
		for (int q = 0; q < 4; ++q)
			st16(st + g + q, y[3 * q], y[3 * q + 1], y[3 * q + 2], tag ^ V[q]);
	}
	for (; g < groups; ++g) {
		const unsigned x0 = fbits(src[3 * g]), x1 = 3 * g + 1 < hop ? fbits(src[3 * g + 1]) : 0u, x2 = 3 * g + 2 < hop ? fbits(src[3 * g + 2]) : 0u;
		st16(st + g, x0, x1, x2, tag ^ zen_group_hash(x0, x1, x2));
	}
}

bool rt_unpack_groups(const uint4* st, int hop, unsigned tag, float* dst, int& g)
{
	const int groups = (hop + 2) / 3;
	const int line_groups = 4 * (hop / 12);
	for (; g < line_groups; g += 4) {
		unsigned y[12], w[4], v[4];
		for (int q = 0; q < 4; ++q) {
			ld16(st + g + q, v);
			y[3 * q] = v[0];
			y[3 * q + 1] = v[1];
			y[3 * q + 2] = v[2];
			w[q] = v[3];
		}
		for (int l = 0; l < 4; ++l)
			if ((w[l] ^ y[l] ^ rotl32s(y[4 + l], 11) ^ rotl32s(y[8 + l], 22)) != tag)
				return false;  // the line is taken as a whole or not at all
		for (int l = 0; l < 12; ++l)
			dst[3 * g + l] = bitsf(y[l]);
	}
	for (; g < groups; ++g) {
		unsigned v[4];
		ld16(st + g, v);
		if ((v[3] ^ zen_group_hash(v[0], v[1], v[2])) != tag)
			return false;
		for (int j = 0; j < 3 && 3 * g + j < hop; ++j)
			dst[3 * g + j] = bitsf(v[j]);
	}
	return true;
}

bool rt_last_group_ready(const uint4* st, int hop, unsigned tag)
{
	const int groups = (hop + 2) / 3, line_groups = 4 * (hop / 12);
	if (groups > line_groups) {
		unsigned v[4];
		ld16(st + groups - 1, v);
		return (v[3] ^ zen_group_hash(v[0], v[1], v[2])) == tag;
	}
	float scratch[12];
	int g0 = 0;
	return rt_unpack_groups(st + groups - 4, 12, tag, scratch, g0);
}

void rt_publish(zen_hpr* h, unsigned tag, const float* src)
{
	uint4* st = h->rt_stage_in;
	if (src)
		rt_pack_groups(src, h->hop, tag, st);
	else
		for (int g = 0; g < h->rt_groups; ++g)  // no samples: {0, 0, 0, tag} (the hash of zeros is zero)
			st16(st + g, 0u, 0u, 0u, tag);
}
#endif

bool rt_unpack(const zen_hpr* h, int o, unsigned tag, float* dst, int& g)
{
	return rt_unpack_groups(h->rt_stage_out[o], h->hop, tag, dst, g);
}

// all groups of output o carry `tag`
bool rt_tags_ready(const zen_hpr* h, int o, unsigned tag)
{
	static thread_local std::vector<float> scratch;
	if ((int)scratch.size() < h->hop) scratch.resize((size_t)h->hop);
	int g = 0;
	return rt_unpack_groups(h->rt_stage_out[o], h->hop, tag, scratch.data(), g);
}

int rt_wait(zen_hpr* h, unsigned op, unsigned& opw, unsigned& tag, unsigned target, const float* push_src, const bool wait_out[3],
            float* const dst[3], std::chrono::steady_clock::time_point* t_seen)
{
	RtCtrl* c = h->rt_ctrl;
	const bool any_out = wait_out[0] || wait_out[1] || wait_out[2];
	std::chrono::steady_clock::time_point t0;  // taken at the first check-point (a hop is over long before)
	unsigned spins = 0;
	// The kernel emits P, H, R in that order (hps.cu:498-579): watch the last group of the last output, then take
	// everything.  Two alternatives were measured and dropped: unpacking each group the moment its tag shows up gave
	// wrong samples now and then (a group is not guaranteed to become visible to the CPU in one piece at that very
	// moment; by the time the LAST group is in, the earlier ones have long settled, and every tag is still checked);
	// and reading the groups while they land, with a second pass at the end, keeps the lines bouncing between the CPU
	// cache and the incoming DMA writes - the hop got 0.9 us slower although the final copy shrank to 0.2 us.
	int prog[3] = {0, 0, 0};
	int last_o = -1;
	if (any_out) last_o = wait_out[2] ? 2 : (wait_out[0] ? 0 : 1);
	for (;;) {
		asm volatile("" ::: "memory");  // the staging buffers change under us: reload them every time round
		if (any_out) {
			if (rt_last_group_ready(h->rt_stage_out[last_o], h->hop, tag)) {
				if (t_seen) *t_seen = std::chrono::steady_clock::now();
				bool ok = true;
				for (int o = 0; o < 3 && ok; ++o) {
					prog[o] = 0;
					if (wait_out[o]) ok = dst[o] ? rt_unpack(h, o, tag, dst[o], prog[o]) : rt_tags_ready(h, o, tag);
				}
				if (ok) break;
			}
		}
		else if (c->seq_out == target)
			break;
		if ((++spins & 1023u) == 0) {
			if (!c->alive) {
				if (c->seq_out == target && !any_out)
					break;
				if (c->seq_out != target) {
					// the kernel left (idle time-out) before it saw the request: bring it back.  The new kernel has no
					// cached pointers, so the request is published again with RT_F_NEW_ARGS (nobody reads the staging
					// buffer between rt_collect and rt_launch).
					int rc = rt_collect(h);
					if (rc == ZEN_OK && op != RT_OP_STOP) {
						opw |= (unsigned)RT_F_NEW_ARGS;
						tag = (target << 8) | opw;
						c->op = opw;
						rt_publish(h, tag, push_src);
						prog[0] = prog[1] = prog[2] = 0;
						rc = rt_launch(h);
						h->rt_args_valid = true;
					}
					if (rc != ZEN_OK) return rc;
					if (op == RT_OP_STOP) break;
				}
			}
			if (spins == 1024u) t0 = std::chrono::steady_clock::now();
			if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(5)) {
				std::fprintf(stderr, "zen_b200: the resident real-time kernel did not answer within 5 s\n");
				return ZEN_ERR_CUDA;
			}
		}
	}
	h->rt_seq = target;
	return ZEN_OK;
}

// complete a submitted-only request (see RtPending)
int rt_drain(zen_hpr* h)
{
	if (!h->rt_pend.active)
		return ZEN_OK;
	auto& p = h->rt_pend;
	float* const none[3] = {nullptr, nullptr, nullptr};
	int rc = rt_wait(h, p.op, p.opw, p.tag, p.target, p.push_src, p.tagged, none, nullptr);
	p.active = false;
	if (rc != ZEN_OK) return rc;
	h->rt_last_tag = p.tag;
	for (int o = 0; o < 3; ++o)
		h->rt_last_tagged[o] = p.tagged[o];
	return ZEN_OK;
}

int rt_call(zen_hpr* h, unsigned op, const float* in, float* o0, float* o1, float* o2, int which)
{
	{
		int rc = rt_drain(h);
		if (rc != ZEN_OK) return rc;
	}
	if (h->plan_dirty) {
		int rc = rt_collect(h);  // cannot be running with a dirty plan, but be safe
		if (rc == ZEN_OK) rc = rebuild_plan_fwd(h);
		if (rc != ZEN_OK) return rc;
	}
	if (!h->rt_running) {
		int rc = rt_launch(h);
		if (rc != ZEN_OK) return rc;
	}
	RtCtrl* c = h->rt_ctrl;
	const unsigned target = h->rt_seq + 1;
	// the pointers usually repeat hop after hop (IOGPU buffers): the kernel re-reads them only when told to
	const bool same = h->rt_args_valid && c->in == in && c->out[0] == o0 && c->out[1] == o1 && c->out[2] == o2 && c->which == which;
	if (!same) {
		c->in = in;
		c->out[0] = o0;
		c->out[1] = o1;
		c->out[2] = o2;
		c->which = which;
		h->rt_args_valid = true;
		// which of them can the host itself read / write?  (mapped pinned memory: IOGPU)
		float* outs[3] = {o0, o1, o2};
		h->rt_in_host = h->rt_push ? static_cast<const float*>(rt_host_alias(in)) : nullptr;
		h->rt_out_all_host = h->rt_push;
		for (int o = 0; o < 3; ++o) {
			h->rt_out_host[o] = h->rt_push ? static_cast<float*>(rt_host_alias(outs[o])) : nullptr;
			if (outs[o] && !h->rt_out_host[o]) h->rt_out_all_host = false;
		}
	}
	unsigned opw = op | (same ? 0u : (unsigned)RT_F_NEW_ARGS);
	// outputs the kernel will emit for this call (hpr_iteration step G)
	bool wait_out[3] = {false, false, false};
	bool any_out = false;
	bool defer = false;
	const float* push_src = nullptr;
	if (op == RT_OP_PROCESS) {
		if (h->rt_stamps) opw |= RT_F_STAMPS;
		if (h->rt_in_host) {
			opw |= RT_F_PUSH_IN;
			push_src = h->rt_in_host;
		}
		const unsigned of = (unsigned)h->plan.dev.out_flags;
		const bool no_dst = !o0 && !o1 && !o2;
		if (h->rt_out_all_host) {
			for (int o = 0; o < 3; ++o) {
				const bool emitted = (of & (1u << o)) && !(o == 2 && (h->plan.dev.soft || h->plan.dev.sse));
				wait_out[o] = emitted && (no_dst || h->rt_out_host[o]);
				any_out = any_out || wait_out[o];
			}
			if (any_out) opw |= RT_F_TAG_OUT;
			defer = no_dst && any_out;  // process_next_hop: submit only, the outputs stay in the staging buffers
		}
		h->rt_last_tagged[0] = h->rt_last_tagged[1] = h->rt_last_tagged[2] = false;  // the staging buffers are about to change
	}
	unsigned tag = (target << 8) | opw;
	std::chrono::steady_clock::time_point tr0, tr1, tr2;
	if (h->rt_trace) tr0 = std::chrono::steady_clock::now();
	c->op = opw;
	zen_host_store_fence();  // a pulled hop may sit in write-combined memory: drain it (and the arguments) before the tags
	rt_publish(h, tag, push_src);
	if (h->rt_trace) tr2 = tr1 = std::chrono::steady_clock::now();
	if (defer) {
		auto& p = h->rt_pend;
		p.active = true;
		p.op = op;
		p.opw = opw;
		p.tag = tag;
		p.target = target;
		p.push_src = push_src;
		for (int o = 0; o < 3; ++o)
			p.tagged[o] = wait_out[o];
		return ZEN_OK;
	}
	if (any_out && (h->plan.dev.out_flags & ZEN_OUTPUT_RESIDUAL) && (h->plan.dev.soft || h->plan.dev.sse) && h->rt_out_host[2])
		std::memset(h->rt_out_host[2], 0, sizeof(float) * (size_t)h->hop);  // the reference's rotate-and-zero (hps.cu:435-449)
	float* dst[3] = {h->rt_out_host[0], h->rt_out_host[1], h->rt_out_host[2]};
	int rc = rt_wait(h, op, opw, tag, target, push_src, wait_out, dst, h->rt_trace ? &tr2 : nullptr);
	if (rc != ZEN_OK) return rc;
	if (op == RT_OP_PROCESS && any_out) {
		h->rt_last_tag = tag;
		for (int o = 0; o < 3; ++o)
			h->rt_last_tagged[o] = wait_out[o];
	}
	if (h->rt_trace && any_out) {
		const auto tr3 = std::chrono::steady_clock::now();
		h->rt_trace_ns[0] += std::chrono::duration<double, std::nano>(tr1 - tr0).count();
		h->rt_trace_ns[1] += std::chrono::duration<double, std::nano>(tr2 - tr1).count();
		h->rt_trace_ns[2] += std::chrono::duration<double, std::nano>(tr3 - tr2).count();
		h->rt_trace_n++;
	}
	return ZEN_OK;
}

// copy_{harmonic,percussive,residual} of a resident session: when the last hop's output o sits in the tagged staging
// buffer and the destination is host-visible, the host unpacks it itself; otherwise the kernel copies it
int rt_copy(zen_hpr* h, int o, float* d_out)
{
	int rc = rt_drain(h);
	if (rc != ZEN_OK) return rc;
	if (h->rt_push && h->rt_last_tagged[o]) {
		if (h->rt_copy_ptr[o] != d_out) {
			h->rt_copy_ptr[o] = d_out;
			h->rt_copy_host[o] = static_cast<float*>(rt_host_alias(d_out));
		}
		if (h->rt_copy_host[o]) {
			int g = 0;
			if (rt_unpack(h, o, h->rt_last_tag, h->rt_copy_host[o], g))
				return ZEN_OK;
		}
	}
	return rt_call(h, RT_OP_COPY, nullptr, d_out, nullptr, nullptr, o);
}

// leave the resident kernel (state goes back to global memory); rt_mode is kept
int rt_pause(zen_hpr* h)
{
	if (!h->rt_running)
		return ZEN_OK;
	{
		int rc = rt_drain(h);  // a submitted hop is completed first (this brings a timed-out kernel back if need be)
		if (rc != ZEN_OK) return rc;
	}
	if (h->rt_ctrl->alive) {
		int rc = rt_call(h, RT_OP_STOP, nullptr, nullptr, nullptr, nullptr, 0);
		if (rc != ZEN_OK) return rc;
		const auto t0 = std::chrono::steady_clock::now();
		while (h->rt_ctrl->alive) {
			if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(5))
				return ZEN_ERR_CUDA;
		}
	}
	return rt_collect(h);
}

int rebuild_plan(zen_hpr* h)
{
	free_plan(h->plan);
	int rc = build_plan(h->plan, h->fs, h->hop, h->beta, h->flags, h->causality, h->copy_bord, h->sse, h->soft);
	h->plan_dirty = false;
	return rc;
}

int rebuild_plan_fwd(zen_hpr* h) { return rebuild_plan(h); }

}  // namespace

extern "C" {

int zen_hpr_create(zen_hpr** out, float fs, int hop, float beta, unsigned flags, int causality, int copy_bord)
{
	if (!out || (causality != ZEN_TIME_CAUSAL && causality != ZEN_TIME_ANTICAUSAL))
		return ZEN_ERR_ARG;
	*out = nullptr;
	if (zen_device_count() <= 0)
		return ZEN_ERR_CUDA;
	zen_hpr* h = new (std::nothrow) zen_hpr();
	if (!h)
		return ZEN_ERR_ARG;
	h->fs = fs;
	h->beta = beta;
	h->flags = flags;
	h->hop = hop;
	h->causality = causality;
	h->copy_bord = copy_bord;
	int rc = build_plan(h->plan, fs, hop, beta, flags, causality, copy_bord, false, false);
	if (rc != ZEN_OK) {
		free_plan(h->plan);
		delete h;
		return rc;
	}
	const zen_geometry& g = h->plan.geom;
	const int M = g.nfft / 2;
	cudaError_t e = cudaStreamCreate(&h->stream);
	if (e == cudaSuccess) e = cudaMalloc(&h->d_input, sizeof(float) * g.nwin);
	for (int o = 0; o < 3 && e == cudaSuccess; ++o)
		e = cudaMalloc(&h->d_ola[o], sizeof(float) * g.nwin);
	if (e == cudaSuccess) e = cudaMalloc(&h->d_mag_ring, sizeof(float) * (size_t)g.stft_width * (M + 1));
	if (e == cudaSuccess) e = cudaMalloc(&h->d_x_ring, sizeof(float2) * (size_t)g.stft_width * (M + 1));
	if (e != cudaSuccess) {
		std::fprintf(stderr, "zen_b200: allocation failed: %s\n", cudaGetErrorString(e));
		zen_hpr_destroy(h);
		return ZEN_ERR_CUDA;
	}
	rc = zen_hpr_reset_buffers(h);
	if (rc != ZEN_OK) {
		zen_hpr_destroy(h);
		return rc;
	}
	*out = h;
	return ZEN_OK;
}

void zen_hpr_destroy(zen_hpr* h)
{
	if (!h)
		return;
	rt_pause(h);
	if (h->rt_trace && h->rt_trace_n > 0)
		std::fprintf(stderr, "zen_b200 rt trace: %ld hops, publish %.0f ns, wait for the device %.0f ns, unpack %.0f ns per hop\n", h->rt_trace_n,
		             h->rt_trace_ns[0] / h->rt_trace_n, h->rt_trace_ns[1] / h->rt_trace_n, h->rt_trace_ns[2] / h->rt_trace_n);
	if (h->rt_ctrl) cudaFreeHost((void*)h->rt_ctrl);
	if (h->rt_stage_in) cudaFreeHost((void*)h->rt_stage_in);
	if (h->rt_stream) cudaStreamDestroy(h->rt_stream);
	cudaFree(h->d_iter);
	if (h->stream) cudaStreamSynchronize(h->stream);
	free_plan(h->plan);
	if (!h->state_external) {
		cudaFree(h->d_input);
		for (int o = 0; o < 3; ++o)
			cudaFree(h->d_ola[o]);
	}
	cudaFree(h->d_mag_ring);
	cudaFree(h->d_x_ring);
	if (h->stream) cudaStreamDestroy(h->stream);
	delete h;
}

int zen_hpr_use_sse_filter(zen_hpr* h)
{
	if (!h) return ZEN_ERR_ARG;
	if (rt_pause(h) != ZEN_OK) return ZEN_ERR_CUDA;
	h->sse = true;
	h->plan_dirty = true;
	return ZEN_OK;
}

int zen_hpr_use_soft_mask(zen_hpr* h)
{
	if (!h) return ZEN_ERR_ARG;
	if (rt_pause(h) != ZEN_OK) return ZEN_ERR_CUDA;
	h->soft = true;
	h->plan_dirty = true;
	return ZEN_OK;
}

int zen_hpr_reset_buffers(zen_hpr* h)
{
	if (!h) return ZEN_ERR_ARG;
	if (rt_pause(h) != ZEN_OK) return ZEN_ERR_CUDA;
	const zen_geometry& g = h->plan.geom;
	const int M = g.nfft / 2;
	ZEN_CUDA_CHECK(cudaMemsetAsync(h->d_input, 0, sizeof(float) * g.nwin, h->stream));
	for (int o = 0; o < 3; ++o)
		ZEN_CUDA_CHECK(cudaMemsetAsync(h->d_ola[o], 0, sizeof(float) * g.nwin, h->stream));
	ZEN_CUDA_CHECK(cudaMemsetAsync(h->d_mag_ring, 0, sizeof(float) * (size_t)g.stft_width * (M + 1), h->stream));
	ZEN_CUDA_CHECK(cudaMemsetAsync(h->d_x_ring, 0, sizeof(float2) * (size_t)g.stft_width * (M + 1), h->stream));
	ZEN_CUDA_CHECK(cudaStreamSynchronize(h->stream));
	h->iter = 0;
	return ZEN_OK;
}

int zen_hpr_get_geometry(const zen_hpr* h, zen_geometry* out)
{
	if (!h || !out) return ZEN_ERR_ARG;
	*out = h->plan.geom;
	return ZEN_OK;
}

int zen_hpr_process_hop_io(zen_hpr* h, const float* d_in_hop, float* eh, float* ep, float* er)
{
	if (!h || !d_in_hop) return ZEN_ERR_ARG;
	if (h->rt_mode)
		return rt_call(h, RT_OP_PROCESS, d_in_hop, eh, ep, er, 0);
	if (h->plan_dirty) {
		int rc = rebuild_plan(h);
		if (rc != ZEN_OK) return rc;
	}
	return dispatch_hop(h, d_in_hop, eh, ep, er);
}

int zen_hpr_process_next_hop(zen_hpr* h, const float* d_in_hop)
{
	return zen_hpr_process_hop_io(h, d_in_hop, nullptr, nullptr, nullptr);
}

static int copy_out(zen_hpr* h, int o, float* d_out)
{
	if (!h || !d_out) return ZEN_ERR_ARG;
	if (h->rt_mode)
		return rt_copy(h, o, d_out);
	int hop = h->hop;
	copy_hop_kernel<<<(hop + 255) / 256, 256, 0, h->stream>>>(h->d_ola[o], d_out, hop);
	ZEN_CUDA_CHECK(cudaGetLastError());
	ZEN_CUDA_CHECK(cudaStreamSynchronize(h->stream));
	return ZEN_OK;
}

int zen_hpr_copy_harmonic(zen_hpr* h, float* d_out) { return copy_out(h, 0, d_out); }
int zen_hpr_copy_percussive(zen_hpr* h, float* d_out) { return copy_out(h, 1, d_out); }
int zen_hpr_copy_residual(zen_hpr* h, float* d_out) { return copy_out(h, 2, d_out); }

int zen_hpr_synchronize(zen_hpr* h)
{
	if (!h) return ZEN_ERR_ARG;
	// real-time session: every call is served synchronously by the resident kernel and nothing is ever queued on the
	// object's stream (rt_launch drains it) - a cudaStreamSynchronize here would only add its ~1.5 us to each hop
	if (h->rt_mode && h->rt_running) return rt_drain(h);
	ZEN_CUDA_CHECK(cudaStreamSynchronize(h->stream));
	return ZEN_OK;
}

// After this returns the caller may overwrite the hop it passed to the last zen_hpr_process_next_hop (in the
// reference the copy of in_hop has completed when process_next_hop returns, hps.cu:452-453).
int zen_hpr_wait_input_consumed(zen_hpr* h)
{
	if (!h) return ZEN_ERR_ARG;
	if (h->rt_mode && h->rt_running) {
		// a pushed hop was copied into the tagged staging buffer at submission
		if (!h->rt_pend.active || h->rt_pend.push_src) return ZEN_OK;
		return rt_drain(h);
	}
	ZEN_CUDA_CHECK(cudaStreamSynchronize(h->stream));
	return ZEN_OK;
}

int zen_hpr_realtime_begin(zen_hpr* h)
{
	if (!h) return ZEN_ERR_ARG;
	if (h->plan.dev.lag != 1) return ZEN_ERR_UNSUPPORTED;  // the resident kernel serves causal streams (HPRRealtime)
	if (h->plan_dirty) {
		int rc = rebuild_plan(h);
		if (rc != ZEN_OK) return rc;
	}
	h->rt_mode = true;
	return h->rt_running ? ZEN_OK : rt_launch(h);
}

// diagnostics: device globaltimer stamps (ns) of the last hop served by the resident kernel
int zen_hpr_realtime_stamps(zen_hpr* h, unsigned long long* out16)
{
	if (!h || !out16 || !h->rt_ctrl) return ZEN_ERR_ARG;
	for (int i = 0; i < 16; ++i)
		out16[i] = h->rt_ctrl->stamps[i];
	return ZEN_OK;
}
// the same for CTA 1 of the cluster (when the hop is split): [9] is the globaltimer at which it saw the command
int zen_hpr_realtime_stamps_rank1(zen_hpr* h, unsigned long long* out16)
{
	if (!h || !out16 || !h->rt_ctrl) return ZEN_ERR_ARG;
	for (int i = 0; i < 16; ++i)
		out16[i] = h->rt_ctrl->stamps_r1[i];
	return ZEN_OK;
}

int zen_hpr_realtime_end(zen_hpr* h)
{
	if (!h) return ZEN_ERR_ARG;
	int rc = rt_pause(h);
	h->rt_mode = false;
	return rc;
}

int zen_hpr_bind_state(zen_hpr* h, float* d_input, float* d_harmonic_out, float* d_percussive_out, float* d_residual_out)
{
	if (!h || !d_input || !d_harmonic_out || !d_percussive_out || !d_residual_out)
		return ZEN_ERR_ARG;
	if (((uintptr_t)d_input | (uintptr_t)d_harmonic_out | (uintptr_t)d_percussive_out | (uintptr_t)d_residual_out) & 7)
		return ZEN_ERR_ARG;
	if (rt_pause(h) != ZEN_OK) return ZEN_ERR_CUDA;
	ZEN_CUDA_CHECK(cudaStreamSynchronize(h->stream));
	if (!h->state_external) {
		cudaFree(h->d_input);
		for (int o = 0; o < 3; ++o)
			cudaFree(h->d_ola[o]);
	}
	h->state_external = true;
	h->d_input = d_input;
	h->d_ola[0] = d_harmonic_out;
	h->d_ola[1] = d_percussive_out;
	h->d_ola[2] = d_residual_out;
	return zen_hpr_reset_buffers(h);
}

float* zen_hpr_state_ptr(zen_hpr* h, int which)
{
	if (!h) return nullptr;
	if (rt_pause(h) != ZEN_OK) return nullptr;  // the overlap-add tails live in the resident kernel's shared memory
	switch (which) {
	case 0: return h->d_input;
	case 1: return h->d_ola[0];
	case 2: return h->d_ola[1];
	case 3: return h->d_ola[2];
	case 4: return h->plan.d_window;
	}
	return nullptr;
}

}  // extern "C"

// ------------------------------------------------------- fakert region ---
// The loop zen/fakert.h:217-251 times, hop by hop: host copy into the mapped
// input buffer -> process_next_hop -> copy_percussive -> host copy out.
// fused != 0 uses the single-launch zen_hpr_process_hop_io instead of the
// process_next_hop + copy_percussive pair.

#include <chrono>

// Host-only test hook: the bin ownership of the cluster-split hop (hpr_split_ranges, hpr_core.cuh).
// out6 = {k0, k1, a0, a1, b0, b1}: own pairs [k0, k1), own bins [a0, a1) and [b0, b1) of CTA `rank` out of `cluster`.
extern "C" int zen_rt_split_ranges(int nfft, int rank, int cluster, int* out6)
{
	if (!out6 || nfft < 8 || (nfft & (nfft - 1)) || cluster < 1 || rank < 0 || rank >= cluster) return ZEN_ERR_ARG;
	hpr_split_ranges(nfft / 2, rank, cluster, out6[0], out6[1], out6[2], out6[3], out6[4], out6[5]);
	return ZEN_OK;
}

// Host-only test hooks for the tagged-group format of the resident kernel's staging buffers (no device involved).
// groups: 16-byte aligned, ceil(hop / 3) * 16 bytes.
extern "C" int zen_rt_pack_groups(const float* src, int hop, unsigned tag, void* groups)
{
	if (!src || !groups || hop < 1 || ((uintptr_t)groups & 15)) return ZEN_ERR_ARG;
	rt_pack_groups(src, hop, tag, static_cast<uint4*>(groups));
	return ZEN_OK;
}

/* returns the number of groups unpacked (== ceil(hop / 3) when every group carries `tag`), negative on bad arguments */
extern "C" int zen_rt_unpack_groups(const void* groups, int hop, unsigned tag, float* dst)
{
	if (!dst || !groups || hop < 1 || ((uintptr_t)groups & 15)) return ZEN_ERR_ARG;
	int g = 0;
	rt_unpack_groups(static_cast<const uint4*>(groups), hop, tag, dst, g);
	return g;
}

extern "C" int zen_fakert_run(float fs, int hop, float beta, int options, const float* h_audio, long n_hops,
                              int warmup_iters, int fused, float* h_perc_out, double* h_us_per_hop)
{
	if (!h_audio || n_hops < 1 || !h_perc_out)
		return ZEN_ERR_ARG;
	zen_hpr* h = nullptr;
	int rc = zen_hpr_create(&h, fs, hop, beta, ZEN_OUTPUT_PERCUSSIVE, ZEN_TIME_CAUSAL, !(options & ZEN_OPT_NOCOPYBORD));
	if (rc != ZEN_OK)
		return rc;
	if (options & ZEN_OPT_SSE) zen_hpr_use_sse_filter(h);
	if (options & ZEN_OPT_SOFT_MASK) zen_hpr_use_soft_mask(h);
	zen_io io;
	rc = zen_io_alloc(&io, hop);
	if (rc != ZEN_OK) {
		zen_hpr_destroy(h);
		return rc;
	}
	const bool two_call = fused == 0 || fused == 3;  // process_next_hop + copy_percussive, as zen/fakert.h:229-230
	if (fused >= 2) {
		rc = zen_hpr_realtime_begin(h);  // resident kernel: no launch, no stream synchronisation per hop
		if (rc != ZEN_OK) {
			zen_io_free(&io);
			zen_hpr_destroy(h);
			return rc;
		}
	}
	// HPRRealtime<GPU>::warmup (hps.cu:392-409): iota data, then reset_buffers
	for (int i = 0; i < warmup_iters && rc == ZEN_OK; ++i) {
		for (int j = 0; j < hop; ++j)
			io.host_in[j] = (float)((long)i * hop + j);
		rc = !two_call ? zen_hpr_process_hop_io(h, io.device_in, nullptr, io.device_out, nullptr) : zen_hpr_process_next_hop(h, io.device_in);
		if (rc == ZEN_OK) rc = zen_hpr_synchronize(h);
	}
	if (rc == ZEN_OK) rc = zen_hpr_reset_buffers(h);
	for (long i = 0; i < n_hops && rc == ZEN_OK; ++i) {
		auto t1 = std::chrono::high_resolution_clock::now();
		std::memcpy(io.host_in, h_audio + (size_t)i * hop, sizeof(float) * hop);
		// the destination of this hop's output is cold: ask for its lines while the device works
		for (int b = 0; b < hop * (int)sizeof(float); b += 64)
			__builtin_prefetch(reinterpret_cast<const char*>(h_perc_out + (size_t)i * hop) + b, 1, 3);
		if (!two_call) {
			rc = zen_hpr_process_hop_io(h, io.device_in, nullptr, io.device_out, nullptr);
			if (rc == ZEN_OK) rc = zen_hpr_synchronize(h);
		}
		else {
			rc = zen_hpr_process_next_hop(h, io.device_in);
			if (rc == ZEN_OK) rc = zen_hpr_copy_percussive(h, io.device_out);
		}
		std::memcpy(h_perc_out + (size_t)i * hop, io.host_out, sizeof(float) * hop);
		auto t2 = std::chrono::high_resolution_clock::now();
		if (h_us_per_hop) h_us_per_hop[i] = std::chrono::duration<double, std::micro>(t2 - t1).count();
	}
	zen_io_free(&io);
	zen_hpr_destroy(h);
	return rc;
}

// ------------------------------------------------ debug-view materialise ---
// The reference recomputes the whole stft_width x nfft matrices every hop and
// exposes them as public members (hps.h:185-194); hps.test.cu style callers may
// read them.  We rebuild them on demand from the ring state with the same
// formulas, off the per-hop path.

namespace {

__global__ void expand_ring_kernel(const float2* __restrict__ x_ring, const float* __restrict__ mag_ring, long iter, int W, int M,
                                   float2* __restrict__ stft, float* __restrict__ s_mag, float* __restrict__ recip, int sse)
{
	const int nfft = 2 * M;
	const int k = blockIdx.x * blockDim.x + threadIdx.x;
	const int row = blockIdx.y;
	if (k >= nfft)
		return;
	const long j = (iter - 1) - (W - 1 - row);  // frame held by this row, newest frame in the last row
	const int kk = k <= M ? k : nfft - k;
	float2 X = make_float2(0.0f, 0.0f);
	float m = 0.0f;
	if (j >= 0) {
		X = x_ring[(size_t)(j % W) * (M + 1) + kk];
		if (k > M) X.y = -X.y;
		m = mag_ring[(size_t)(j % W) * (M + 1) + kk];
	}
	const size_t idx = (size_t)row * nfft + k;
	if (stft) stft[idx] = X;
	if (s_mag) s_mag[idx] = m;
	if (recip) recip[idx] = sse ? (1.0f / m) * 1.0f : 0.0f;
}

__global__ void scale_recip_kernel(float* __restrict__ m, size_t n, float factor)
{
	size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n)
		m[i] = (1.0f / m[i]) * factor;  // hps.h:45-56 reciprocal_functor
}

__global__ void mask_rows_kernel(const HprDev P, int nfft, const float* __restrict__ hm, const float* __restrict__ pm,
                                 float* __restrict__ hmask, float* __restrict__ pmask, float* __restrict__ rmask)
{
	const int k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= nfft)
		return;
	const size_t off = (size_t)(P.W - P.lag) * nfft + k;
	const float H = hm[off], Pv = pm[off];
	float mp = 0.0f, mh = 0.0f;
	if (P.sse) {
		mp = mask_sse(Pv, H);
		mh = mask_sse(H, Pv);
	}
	else if (P.soft) {
		mp = mask_soft(Pv, H, P.power);
		mh = mask_soft(H, Pv, P.power);
	}
	else {
		mp = mask_hard(Pv, H, P.beta);
		mh = mask_hard(H, Pv, P.beta_h);
	}
	if (!(P.out_flags & ZEN_OUTPUT_PERCUSSIVE)) mp = 0.0f;
	if (!(P.out_flags & ZEN_OUTPUT_HARMONIC)) mh = 0.0f;
	if (pmask) pmask[off] = mp;
	if (hmask) hmask[off] = mh;
	if (rmask && (P.out_flags & ZEN_OUTPUT_RESIDUAL) && !P.soft && !P.sse) {
		// residual_mask is recomputed over the whole matrix (hps.cu:565-567); rows other
		// than the consumed one only ever hold 1 - (0 + 0)
		for (int r = 0; r < P.W; ++r)
			rmask[(size_t)r * nfft + k] = r == P.W - P.lag ? 1.0f - (mh + mp) : 1.0f;
	}
}

}  // namespace

extern "C" int zen_hpr_materialize(zen_hpr* h, float* d_stft, float* d_s_mag, float* d_hm, float* d_pm,
                                   float* d_hmask, float* d_pmask, float* d_rmask)
{
	if (!h)
		return ZEN_ERR_ARG;
	if (rt_pause(h) != ZEN_OK)
		return ZEN_ERR_CUDA;
	if (h->plan_dirty) {
		int rc = rebuild_plan(h);
		if (rc != ZEN_OK) return rc;
	}
	const zen_geometry& g = h->plan.geom;
	const int W = g.stft_width, nfft = g.nfft, M = nfft / 2;
	const size_t n = (size_t)W * nfft;
	const bool need_filters = d_hm || d_pm || d_hmask || d_pmask || d_rmask;
	float *tmp_mag = nullptr, *tmp_rec = nullptr, *tmp_h = nullptr, *tmp_p = nullptr;
	auto cleanup = [&]() { cudaFree(tmp_mag); cudaFree(tmp_rec); cudaFree(tmp_h); cudaFree(tmp_p); };
	float* s_mag = d_s_mag;
	if (!s_mag && need_filters) {
		ZEN_CUDA_CHECK(cudaMalloc(&tmp_mag, sizeof(float) * n));
		s_mag = tmp_mag;
	}
	if (h->sse && need_filters)
		ZEN_CUDA_CHECK(cudaMalloc(&tmp_rec, sizeof(float) * n));
	dim3 grid((nfft + 255) / 256, W);
	expand_ring_kernel<<<grid, 256, 0, h->stream>>>(h->d_x_ring, h->d_mag_ring, h->iter, W, M, reinterpret_cast<float2*>(d_stft),
	                                               s_mag, tmp_rec, h->sse ? 1 : 0);
	int rc = ZEN_OK;
	if (need_filters) {
		float* hm = d_hm;
		float* pm = d_pm;
		if (!hm) {
			if (cudaMalloc(&tmp_h, sizeof(float) * n) != cudaSuccess || cudaMemsetAsync(tmp_h, 0, sizeof(float) * n, h->stream) != cudaSuccess) {
				cleanup();
				return ZEN_ERR_CUDA;
			}
			hm = tmp_h;
		}
		if (!pm) {
			if (cudaMalloc(&tmp_p, sizeof(float) * n) != cudaSuccess || cudaMemsetAsync(tmp_p, 0, sizeof(float) * n, h->stream) != cudaSuccess) {
				cleanup();
				return ZEN_ERR_CUDA;
			}
			pm = tmp_p;
		}
		if (!h->sse) {
			rc = zen_median_filter(W, nfft, g.l_harm, h->causality, h->copy_bord, s_mag, hm, h->stream);
			if (rc == ZEN_OK)
				rc = zen_median_filter(W, nfft, g.l_perc, ZEN_FREQUENCY, h->copy_bord, s_mag, pm, h->stream);
		}
		else {
			rc = zen_box_filter(W, nfft, g.l_harm, h->causality, tmp_rec, hm, h->stream);
			if (rc == ZEN_OK)
				rc = zen_box_filter(W, nfft, g.l_perc, ZEN_FREQUENCY, tmp_rec, pm, h->stream);
			if (rc == ZEN_OK) {
				scale_recip_kernel<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(pm, n, (float)g.l_perc + 1.0f);
				scale_recip_kernel<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(hm, n, (float)g.l_harm + 1.0f);
			}
		}
		if (rc == ZEN_OK && (d_hmask || d_pmask || d_rmask))
			mask_rows_kernel<<<(nfft + 255) / 256, 256, 0, h->stream>>>(h->plan.dev, nfft, hm, pm, d_hmask, d_pmask, d_rmask);
	}
	cudaError_t e = cudaStreamSynchronize(h->stream);
	cleanup();
	if (e != cudaSuccess || cudaGetLastError() != cudaSuccess)
		return ZEN_ERR_CUDA;
	return rc;
}

// --------------------------------------------------------------- batched ---

// host-buffer pipeline depth: H2D of chunk n+1, kernel of chunk n and D2H of chunk n-1 overlap
constexpr int ZEN_PIPE_SLOTS = 3;

struct zen_hpr_batch {
	Plan plan;
	int max_streams;
	long max_hops;
	int tile_hops;
	float* d_scratch = nullptr;   // [ZEN_PIPE_SLOTS pipeline slots][resident CTAs] scratch states
	int* d_counters = nullptr;    // work-queue heads, one per pipeline slot
	int resident = 0;             // resident CTAs of the tile kernel on this device
	size_t scratch_ctas = 0;
	long last_launches = 0;
	float last_kernel_ms = 0.0f;
	cudaEvent_t ev0 = nullptr, ev1 = nullptr;
	// host pipeline: per slot a float staging buffer for the input and one per requested output, and (PCM16 entry
	// point) the int16 images of both plus the per-stream peaks.  Capacities are tracked in ELEMENTS per buffer, so
	// a later call with longer rows, more streams per chunk or another set of outputs regrows exactly what it lacks.
	cudaStream_t streams[ZEN_PIPE_SLOTS] = {};
	float* d_stage_in[ZEN_PIPE_SLOTS] = {};
	size_t cap_in[ZEN_PIPE_SLOTS] = {};
	float* d_stage_out[ZEN_PIPE_SLOTS][3] = {};
	size_t cap_out[ZEN_PIPE_SLOTS][3] = {};
	int16_t* d_pcm_in[ZEN_PIPE_SLOTS] = {};
	size_t cap_pcm_in[ZEN_PIPE_SLOTS] = {};
	int16_t* d_pcm_out[ZEN_PIPE_SLOTS][3] = {};
	size_t cap_pcm_out[ZEN_PIPE_SLOTS][3] = {};
	float* d_peaks[ZEN_PIPE_SLOTS][3] = {};
	size_t cap_peaks[ZEN_PIPE_SLOTS][3] = {};
};

namespace {

// Tile length.  Work items (stream, tile) are pulled from a queue by the resident CTAs, so what matters
// is (a) enough items per resident CTA that the tail, when the queue runs dry, is short, and (b) tiles long
// enough that the W-hop halo each tile re-analyses stays a small fraction of its work.
int choose_tile_hops(const Plan& pl, int n_streams, long n_hops, int resident)
{
	if (resident < 1) resident = 148;
	const long items_wanted = 32L * resident;  // many small items: the tail when the queue runs dry stays short
	long tiles_wanted = (items_wanted + n_streams - 1) / n_streams;
	if (tiles_wanted < 1) tiles_wanted = 1;
	long tile = (n_hops + tiles_wanted - 1) / tiles_wanted;
	long min_tile = 4L * pl.dev.W;  // the halo (analysis only, roughly a third of a full hop) stays below ~10 % of the tile's work
	if (min_tile < 16) min_tile = 16;
	if (tile < min_tile) tile = min_tile;
	// Items are queued tile-major (all first tiles, then all second tiles, ...), so the LAST tiles of the streams are what
	// the CTAs work on when the queue runs dry.  With enough streams to give every resident CTA one of them, the regular
	// tiles are stretched a little so that the last one is an eighth of their length: the CTAs then finish within an
	// eighth of an item of each other instead of a whole one.
	if (n_streams >= resident && tile < n_hops && !std::getenv("ZEN_B200_NO_STRETCH")) {
		const long k = (n_hops + tile - 1) / tile - 1;  // regular tiles per stream
		if (k >= 1) {
			long stretched = (8 * n_hops + 8 * k) / (8 * k + 1);  // k tiles + one eighth cover n_hops
			if (stretched * k < n_hops && stretched > tile / 2) tile = stretched;
		}
	}
	if (tile > n_hops) tile = n_hops;
	if (tile < 1) tile = 1;
	return (int)tile;
}

int ensure_scratch(zen_hpr_batch* b)
{
	if (b->d_scratch)
		return ZEN_OK;
	b->resident = resident_ctas_for(b->plan);
	if (b->resident < 1)
		return ZEN_ERR_CUDA;
	ZEN_CUDA_CHECK(cudaMalloc(&b->d_scratch, sizeof(float) * tile_scratch_floats(b->plan) * ZEN_PIPE_SLOTS * (size_t)b->resident));
	ZEN_CUDA_CHECK(cudaMalloc(&b->d_counters, sizeof(int) * ZEN_PIPE_SLOTS));
	return ZEN_OK;
}

}  // namespace

extern "C" {

int zen_hpr_batch_create(zen_hpr_batch** out, float fs, int hop, float beta, unsigned flags, int causality,
                         int options, int max_streams, long max_hops)
{
	if (!out || max_streams < 1 || max_hops < 1)
		return ZEN_ERR_ARG;
	*out = nullptr;
	if (zen_device_count() <= 0)
		return ZEN_ERR_CUDA;
	zen_hpr_batch* b = new (std::nothrow) zen_hpr_batch();
	if (!b)
		return ZEN_ERR_ARG;
	int rc = build_plan(b->plan, fs, hop, beta, flags, causality, !(options & ZEN_OPT_NOCOPYBORD),
	                    (options & ZEN_OPT_SSE) != 0, (options & ZEN_OPT_SOFT_MASK) != 0);
	if (rc != ZEN_OK) {
		free_plan(b->plan);
		delete b;
		return rc;
	}
	b->max_streams = max_streams;
	b->max_hops = max_hops;
	cudaEventCreate(&b->ev0);
	cudaEventCreate(&b->ev1);
	*out = b;
	return ZEN_OK;
}

void zen_hpr_batch_destroy(zen_hpr_batch* b)
{
	if (!b)
		return;
	free_plan(b->plan);
	cudaFree(b->d_scratch);
	cudaFree(b->d_counters);
	for (int s = 0; s < ZEN_PIPE_SLOTS; ++s) {
		cudaFree(b->d_stage_in[s]);
		cudaFree(b->d_pcm_in[s]);
		for (int o = 0; o < 3; ++o) {
			cudaFree(b->d_stage_out[s][o]);
			cudaFree(b->d_pcm_out[s][o]);
			cudaFree(b->d_peaks[s][o]);
		}
		if (b->streams[s]) cudaStreamDestroy(b->streams[s]);
	}
	if (b->ev0) cudaEventDestroy(b->ev0);
	if (b->ev1) cudaEventDestroy(b->ev1);
	delete b;
}

int zen_hpr_batch_process(zen_hpr_batch* b, const float* d_in, long in_stride, int n_streams, long n_hops,
                          float* d_out_h, float* d_out_p, float* d_out_r, long out_stride, void* cuda_stream)
{
	if (!b || !d_in || n_streams < 1 || n_hops < 1)
		return ZEN_ERR_ARG;
	if ((in_stride & 1) || (out_stride & 1) || ((uintptr_t)d_in & 7) || ((uintptr_t)d_out_h & 7) || ((uintptr_t)d_out_p & 7)
	    || ((uintptr_t)d_out_r & 7) || in_stride < n_hops * b->plan.dev.hop || out_stride < n_hops * b->plan.dev.hop)
		return ZEN_ERR_ARG;
	cudaStream_t s = (cudaStream_t)cuda_stream;
	int rc = ensure_scratch(b);
	if (rc != ZEN_OK)
		return rc;
	int tile = choose_tile_hops(b->plan, n_streams, n_hops, b->resident);
	b->tile_hops = tile;
	const unsigned f = b->plan.dev.out_flags;
	// the soft-mask / SSE variants never write the residual: the reference's residual_out is only rotated and
	// zero-filled (hps.cu:435-449, 562), so N process_next_hop calls emit zeros
	if ((f & 4) && d_out_r && (b->plan.dev.soft || b->plan.dev.sse))
		ZEN_CUDA_CHECK(cudaMemset2DAsync(d_out_r, (size_t)out_stride * sizeof(float), 0, (size_t)n_hops * b->plan.dev.hop * sizeof(float),
		                                 (size_t)n_streams, s));
	cudaEventRecord(b->ev0, s);
	rc = dispatch_tile(b->plan, d_in, in_stride, (f & 1) ? d_out_h : nullptr, (f & 2) ? d_out_p : nullptr,
	                   (f & 4) ? d_out_r : nullptr, out_stride, n_streams, n_hops, tile, b->d_scratch, b->d_counters, b->resident, s);
	cudaEventRecord(b->ev1, s);
	b->last_launches = 1;
	return rc;
}

long zen_hpr_batch_last_launches(const zen_hpr_batch* b) { return b ? b->last_launches : 0; }

float zen_hpr_batch_last_kernel_ms(const zen_hpr_batch* b)
{
	if (!b)
		return 0.0f;
	float ms = 0.0f;
	if (cudaEventSynchronize(b->ev1) != cudaSuccess || cudaEventElapsedTime(&ms, b->ev0, b->ev1) != cudaSuccess)
		return -1.0f;
	return ms;
}

}  // extern "C"

namespace {

template <typename T>
int grow(T*& p, size_t& cap, size_t need)
{
	if (cap >= need && p)
		return ZEN_OK;
	cudaFree(p);
	p = nullptr;
	cap = 0;
	ZEN_CUDA_CHECK(cudaMalloc(&p, need * sizeof(T)));
	cap = need;
	return ZEN_OK;
}

// streams per chunk so that one chunk's float input is ~chunk_mb
int host_chunk_streams(size_t row, int n_streams)
{
	size_t chunk_mb = 128;
	if (const char* e = std::getenv("ZEN_B200_CHUNK_MB")) {
		long v = std::atol(e);
		if (v > 0) chunk_mb = (size_t)v;
	}
	return (int)std::max<size_t>(1, std::min<size_t>((size_t)n_streams, (chunk_mb << 20) / (row * sizeof(float)) + 1));
}

// The host pipeline shared by the float and the PCM16 entry points.  Streams are cut into chunks; chunk c runs on
// CUDA stream c mod ZEN_PIPE_SLOTS: copy in, (decode,) fused HPR kernel, (peak + encode,) copy out, so the H2D copy of
// one chunk, the kernel of the previous one and the D2H copy of the one before overlap.
template <typename T>
int batch_process_host_impl(zen_hpr_batch* b, const T* h_in, long in_stride, int n_streams, long n_hops, T* h_out_h, T* h_out_p,
                            T* h_out_r, long out_stride, float* h_peaks_h, float* h_peaks_p, float* h_peaks_r)
{
	constexpr bool PCM = sizeof(T) == 2;
	if (!b || !h_in || n_streams < 1 || n_hops < 1)
		return ZEN_ERR_ARG;
	const int hop = b->plan.dev.hop;
	const size_t row = (size_t)n_hops * hop;
	if ((size_t)in_stride < row || (size_t)out_stride < row)
		return ZEN_ERR_ARG;
	const unsigned f = b->plan.dev.out_flags;
	T* h_out[3] = {(f & 1) ? h_out_h : nullptr, (f & 2) ? h_out_p : nullptr, (f & 4) ? h_out_r : nullptr};
	float* h_peaks[3] = {h_peaks_h, h_peaks_p, h_peaks_r};
	const bool zero_r = (f & 4) && (b->plan.dev.soft || b->plan.dev.sse);  // hps.cu:562: residual never written
	const int chunk = host_chunk_streams(row, n_streams);
	// device rows are padded to a multiple of 8 elements so that the 16-byte paths of the PCM kernels apply
	const size_t drow = (row + 7) & ~(size_t)7;
	for (int s = 0; s < ZEN_PIPE_SLOTS; ++s) {
		if (!b->streams[s]) ZEN_CUDA_CHECK(cudaStreamCreateWithFlags(&b->streams[s], cudaStreamNonBlocking));
		int rc = grow(b->d_stage_in[s], b->cap_in[s], drow * chunk);
		if (rc == ZEN_OK && PCM) rc = grow(b->d_pcm_in[s], b->cap_pcm_in[s], drow * chunk);
		for (int o = 0; o < 3 && rc == ZEN_OK; ++o) {
			if (!h_out[o]) continue;
			rc = grow(b->d_stage_out[s][o], b->cap_out[s][o], drow * chunk);
			if (rc == ZEN_OK && PCM) rc = grow(b->d_pcm_out[s][o], b->cap_pcm_out[s][o], drow * chunk);
			if (rc == ZEN_OK && PCM) rc = grow(b->d_peaks[s][o], b->cap_peaks[s][o], (size_t)chunk);
		}
		if (rc != ZEN_OK)
			return rc;
	}
	int rc = ensure_scratch(b);
	if (rc != ZEN_OK)
		return rc;
	int tile = choose_tile_hops(b->plan, chunk, n_hops, b->resident);
	b->tile_hops = tile;
	// the pipeline slots can run concurrently: separate scratch areas and work counters
	const size_t scratch_half = tile_scratch_floats(b->plan) * (size_t)b->resident;
	long launches = 0;
	int slot = 0;
	for (int s0 = 0; s0 < n_streams; s0 += chunk, slot = (slot + 1) % ZEN_PIPE_SLOTS) {
		const int ns = std::min(chunk, n_streams - s0);
		cudaStream_t st = b->streams[slot];
		if (PCM) {
			ZEN_CUDA_CHECK(cudaMemcpy2DAsync(b->d_pcm_in[slot], drow * sizeof(T), h_in + (size_t)s0 * in_stride,
			                                 (size_t)in_stride * sizeof(T), row * sizeof(T), ns, cudaMemcpyHostToDevice, st));
			rc = zen_pcm16_decode_mono_async(reinterpret_cast<const int16_t*>(b->d_pcm_in[slot]), (long)drow, 1, ns, (long)row,
			                                 b->d_stage_in[slot], (long)drow, st);
			if (rc != ZEN_OK)
				return rc;
			++launches;
		}
		else {
			ZEN_CUDA_CHECK(cudaMemcpy2DAsync(b->d_stage_in[slot], drow * sizeof(float), h_in + (size_t)s0 * in_stride,
			                                 (size_t)in_stride * sizeof(T), row * sizeof(T), ns, cudaMemcpyHostToDevice, st));
		}
		if (zero_r && h_out[2])
			ZEN_CUDA_CHECK(cudaMemsetAsync(b->d_stage_out[slot][2], 0, drow * sizeof(float) * ns, st));
		// PCM16: the peaks the outputs are normalised by come out of the fused kernel itself where it can track them
		// (max |emitted sample| per stream, one atomicMax per tile), otherwise from a pass over the outputs
		const bool fused_peaks = PCM && tile_peaks_supported(b->plan) && !(zero_r && h_out[2]);
		float* pk[3] = {nullptr, nullptr, nullptr};
		if (fused_peaks)
			for (int o = 0; o < 3; ++o)
				if (h_out[o]) {
					pk[o] = b->d_peaks[slot][o];
					ZEN_CUDA_CHECK(cudaMemsetAsync(pk[o], 0, sizeof(float) * ns, st));
				}
		rc = dispatch_tile(b->plan, b->d_stage_in[slot], (long)drow, h_out[0] ? b->d_stage_out[slot][0] : nullptr,
		                   h_out[1] ? b->d_stage_out[slot][1] : nullptr, h_out[2] ? b->d_stage_out[slot][2] : nullptr, (long)drow, ns,
		                   n_hops, tile, b->d_scratch + slot * scratch_half, b->d_counters + slot, b->resident, st,
		                   fused_peaks ? pk : nullptr);
		if (rc != ZEN_OK)
			return rc;
		++launches;
		for (int o = 0; o < 3; ++o) {
			if (!h_out[o]) continue;
			if (PCM) {
				// what the command line does with every output: x / max|x|, then PCM16 (zen/offline.h:180-223)
				if (!fused_peaks) {
					rc = zen_pcm16_peaks_async(b->d_stage_out[slot][o], (long)drow, ns, (long)row, b->d_peaks[slot][o], st);
					if (rc != ZEN_OK)
						return rc;
					++launches;
				}
				rc = zen_pcm16_encode_with_peaks_async(b->d_stage_out[slot][o], (long)drow, ns, (long)row, b->d_peaks[slot][o],
				                                       reinterpret_cast<int16_t*>(b->d_pcm_out[slot][o]), (long)drow, st);
				if (rc != ZEN_OK)
					return rc;
				++launches;
				ZEN_CUDA_CHECK(cudaMemcpy2DAsync(h_out[o] + (size_t)s0 * out_stride, (size_t)out_stride * sizeof(T), b->d_pcm_out[slot][o],
				                                 drow * sizeof(T), row * sizeof(T), ns, cudaMemcpyDeviceToHost, st));
				if (h_peaks[o])
					ZEN_CUDA_CHECK(cudaMemcpyAsync(h_peaks[o] + s0, b->d_peaks[slot][o], sizeof(float) * ns, cudaMemcpyDeviceToHost, st));
			}
			else {
				ZEN_CUDA_CHECK(cudaMemcpy2DAsync(h_out[o] + (size_t)s0 * out_stride, (size_t)out_stride * sizeof(T), b->d_stage_out[slot][o],
				                                 drow * sizeof(float), row * sizeof(T), ns, cudaMemcpyDeviceToHost, st));
			}
		}
	}
	for (int s = 0; s < ZEN_PIPE_SLOTS; ++s)
		ZEN_CUDA_CHECK(cudaStreamSynchronize(b->streams[s]));
	b->last_launches = launches;
	return ZEN_OK;
}

}  // namespace

extern "C" int zen_hpr_batch_process_host(zen_hpr_batch* b, const float* h_in, long in_stride, int n_streams, long n_hops,
                                          float* h_out_h, float* h_out_p, float* h_out_r, long out_stride)
{
	return batch_process_host_impl<float>(b, h_in, in_stride, n_streams, n_hops, h_out_h, h_out_p, h_out_r, out_stride, nullptr,
	                                      nullptr, nullptr);
}

extern "C" int zen_hpr_batch_process_host_pcm16(zen_hpr_batch* b, const int16_t* h_in, long in_stride, int n_streams, long n_hops,
                                                int16_t* h_out_h, int16_t* h_out_p, int16_t* h_out_r, long out_stride,
                                                float* h_peaks_h, float* h_peaks_p, float* h_peaks_r)
{
	return batch_process_host_impl<int16_t>(b, h_in, in_stride, n_streams, n_hops, h_out_h, h_out_p, h_out_r, out_stride, h_peaks_h,
	                                        h_peaks_p, h_peaks_r);
}

extern "C" {

// ---------------------------------------------------------------- offline ---

static int chunk_padder(long size, int hop, int lag, long* padded)
{
	// hps.cu:109-126, chunk count evaluated in float as written
	int n_chunks = (int)std::ceil((float)size / (float)hop);
	long pad = (long)n_chunks * hop - size;
	pad += (long)lag * hop;
	n_chunks += lag;
	*padded = size + pad;
	return n_chunks;
}

int zen_offline_process_device(float fs, int hop_h, int hop_p, float beta_h, float beta_p, int options,
                               const float* d_audio, long n, float* d_h, float* d_p, float* d_r, void* cuda_stream)
{
	if (!d_audio || n < 1 || !d_h || !d_p || !d_r)
		return ZEN_ERR_ARG;
	if (hop_p <= 0 || hop_h % hop_p != 0)  // hps.cu:33-36
		return ZEN_ERR_GEOMETRY;
	cudaStream_t s = (cudaStream_t)cuda_stream;
	const int opt = options;
	zen_hpr_batch *bh = nullptr, *bp = nullptr;
	int rc = zen_hpr_batch_create(&bh, fs, hop_h, beta_h, ZEN_OUTPUT_HARMONIC | ZEN_OUTPUT_PERCUSSIVE | ZEN_OUTPUT_RESIDUAL,
	                              ZEN_TIME_ANTICAUSAL, opt, 1, 1);
	if (rc == ZEN_OK)
		rc = zen_hpr_batch_create(&bp, fs, hop_p, beta_p, ZEN_OUTPUT_PERCUSSIVE, ZEN_TIME_ANTICAUSAL, opt, 1, 1);
	float *a1 = nullptr, *h1 = nullptr, *p1 = nullptr, *r1 = nullptr, *a2 = nullptr, *p2 = nullptr;
	auto cleanup = [&]() {
		cudaFree(a1); cudaFree(h1); cudaFree(p1); cudaFree(r1); cudaFree(a2); cudaFree(p2);
		zen_hpr_batch_destroy(bh);
		zen_hpr_batch_destroy(bp);
	};
	if (rc != ZEN_OK) {
		cleanup();
		return rc;
	}
	long padded1 = 0, padded2 = 0;
	const int n1 = chunk_padder(n, hop_h, bh->plan.geom.lag, &padded1);
	const int n2 = chunk_padder(n, hop_p, bp->plan.geom.lag, &padded2);
	const long shift1 = (long)bh->plan.geom.lag * hop_h, shift2 = (long)bp->plan.geom.lag * hop_p;
	cudaError_t e = cudaMalloc(&a1, sizeof(float) * padded1);
	if (e == cudaSuccess) e = cudaMalloc(&h1, sizeof(float) * padded1);
	if (e == cudaSuccess) e = cudaMalloc(&p1, sizeof(float) * padded1);
	if (e == cudaSuccess) e = cudaMalloc(&r1, sizeof(float) * padded1);
	if (e == cudaSuccess) e = cudaMalloc(&a2, sizeof(float) * padded2);
	if (e == cudaSuccess) e = cudaMalloc(&p2, sizeof(float) * padded2);
	if (e == cudaSuccess) e = cudaMemsetAsync(a1, 0, sizeof(float) * padded1, s);
	if (e == cudaSuccess) e = cudaMemcpyAsync(a1, d_audio, sizeof(float) * n, cudaMemcpyDeviceToDevice, s);
	// the soft-mask / SSE variants never write the residual: it stays zero (hps.cu:562)
	if (e == cudaSuccess) e = cudaMemsetAsync(r1, 0, sizeof(float) * padded1, s);
	if (e != cudaSuccess) {
		std::fprintf(stderr, "zen_b200: offline allocation failed: %s\n", cudaGetErrorString(e));
		cleanup();
		return ZEN_ERR_CUDA;
	}
	rc = zen_hpr_batch_process(bh, a1, padded1, 1, n1, h1, p1, r1, padded1, s);
	if (rc == ZEN_OK) {
		offline_intermediate_kernel<<<(unsigned)((padded2 + 255) / 256), 256, 0, s>>>(p1, r1, a2, padded1, shift1, padded2);
		rc = zen_hpr_batch_process(bp, a2, padded2, 1, n2, nullptr, p2, nullptr, padded2, s);
	}
	if (rc == ZEN_OK) {
		e = cudaMemcpyAsync(d_h, h1 + shift1, sizeof(float) * n, cudaMemcpyDeviceToDevice, s);
		if (e == cudaSuccess) e = cudaMemcpyAsync(d_p, p2 + shift2, sizeof(float) * n, cudaMemcpyDeviceToDevice, s);
		if (e == cudaSuccess) e = cudaMemsetAsync(d_r, 0, sizeof(float) * n, s);  // hps.cu:200-204, 219-220
		if (e == cudaSuccess) e = cudaStreamSynchronize(s);
		if (e != cudaSuccess) {
			std::fprintf(stderr, "zen_b200: offline failed: %s\n", cudaGetErrorString(e));
			rc = ZEN_ERR_CUDA;
		}
	}
	cleanup();
	return rc;
}

int zen_offline_process(float fs, int hop_h, int hop_p, float beta_h, float beta_p, int options,
                        const float* h_audio, long n, float* h_h, float* h_p, float* h_r)
{
	if (!h_audio || n < 1 || !h_h || !h_p || !h_r)
		return ZEN_ERR_ARG;
	if (hop_p <= 0 || hop_h % hop_p != 0)
		return ZEN_ERR_GEOMETRY;
	if (zen_device_count() <= 0)
		return ZEN_ERR_CUDA;
	float* d = nullptr;
	ZEN_CUDA_CHECK(cudaMalloc(&d, sizeof(float) * 4 * (size_t)n));
	cudaError_t e = cudaMemcpy(d, h_audio, sizeof(float) * n, cudaMemcpyHostToDevice);
	int rc = e == cudaSuccess ? zen_offline_process_device(fs, hop_h, hop_p, beta_h, beta_p, options, d, n, d + n, d + 2 * n,
	                                                      d + 3 * n, nullptr)
	                          : ZEN_ERR_CUDA;
	if (rc == ZEN_OK) {
		e = cudaMemcpy(h_h, d + n, sizeof(float) * n, cudaMemcpyDeviceToHost);
		if (e == cudaSuccess) e = cudaMemcpy(h_p, d + 2 * n, sizeof(float) * n, cudaMemcpyDeviceToHost);
		if (e == cudaSuccess) e = cudaMemcpy(h_r, d + 3 * n, sizeof(float) * n, cudaMemcpyDeviceToHost);
		if (e != cudaSuccess) rc = ZEN_ERR_CUDA;
	}
	cudaFree(d);
	return rc;
}

}  // extern "C"
