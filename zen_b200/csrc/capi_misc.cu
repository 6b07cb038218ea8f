// Host-side pieces of the C ABI: derived sizes, window, mapped I/O buffers.
#include <cmath>
#include <cstring>
#include <vector>

#include "zen_common.cuh"

extern "C" {

const char* zen_b200_version(void) { return "zen_b200 0.1 (sm_100a)"; }

int zen_device_count(void)
{
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) {
		cudaGetLastError();
		return 0;
	}
	return n;
}

// Window<T>::Window, libzen/win.h:21-53: periodic Hann (or its square root),
// evaluated on the host in float with cosf/sqrtf and PI = 3.14159265359F so
// the table is bit-identical to the reference's.
int zen_window(int type, int n, float* h_out)
{
	if (n < 0 || !h_out || (type != ZEN_WIN_SQRT_VON_HANN && type != ZEN_WIN_VON_HANN))
		return ZEN_ERR_ARG;
	const float PI = 3.14159265359F;
	const float N = (float)n;
	for (int i = 0; i < n; ++i) {
		float h = 0.5F * (1.0F - cosf(2.0F * PI * (float)i / N));
		h_out[i] = (type == ZEN_WIN_SQRT_VON_HANN) ? sqrtf(h) : h;
	}
	return ZEN_OK;
}

// HPR<B>::HPR member initialisers, libzen/hps.h:222-230 and 265-274, with the
// reference's float/double mix: l_harm = roundf(0.2 / ((float)(nfft - hop) / fs)),
// l_perc = roundf(500 / (fs / (float)nfft)), COLA = nfft / sum(win^2) summed
// sequentially in float.
int zen_hpr_geometry(float fs, int hop, int causal, zen_geometry* g)
{
	if (!g || hop < 1 || !(fs > 0.0f))
		return ZEN_ERR_ARG;
	g->hop = hop;
	g->nwin = 2 * hop;
	g->nfft = 4 * hop;
	g->l_harm = (int)roundf((float)(0.2 / ((float)(g->nfft - hop) / fs)));
	g->l_perc = (int)roundf(500 / (fs / (float)g->nfft));
	g->lag = causal ? 1 : g->l_harm;
	g->stft_width = 2 * g->l_harm;
	std::vector<float> w(g->nwin);
	zen_window(ZEN_WIN_SQRT_VON_HANN, g->nwin, w.data());
	volatile float acc = 0.0f;  // volatile: keep the sequential float sum as written (no vectorised reassociation)
	for (int i = 0; i < g->nwin; ++i) {
		float sq = w[i] * w[i];
		acc = acc + sq;
	}
	g->cola_factor = (float)g->nfft / acc;
	return ZEN_OK;
}

// IOGPU, libzen/libzen/io.h:24-70: mapped + portable pinned buffers (input also
// write-combined) and their device aliases.
void zen_io_free(zen_io* io);

int zen_io_alloc(zen_io* io, size_t size)
{
	if (!io || size == 0)
		return ZEN_ERR_ARG;
	std::memset(io, 0, sizeof(*io));
	io->size = size;
	unsigned flags = cudaHostAllocMapped | cudaHostAllocPortable;
	// The reference allocates host_in write-combined (libzen/libzen/io.h:30-34).  Ours is ordinary cached pinned memory:
	// the resident real-time kernel's host side re-reads the hop to push it in tagged groups (rt_publish), and reading
	// write-combined memory back on the host is uncached.
	cudaError_t e = cudaHostAlloc((void**)&io->host_in, size * sizeof(float), flags);
	if (e == cudaSuccess) e = cudaHostAlloc((void**)&io->host_out, size * sizeof(float), flags);
	if (e == cudaSuccess) e = cudaHostGetDevicePointer((void**)&io->device_in, io->host_in, 0);
	if (e == cudaSuccess) e = cudaHostGetDevicePointer((void**)&io->device_out, io->host_out, 0);
	if (e != cudaSuccess) {
		std::fprintf(stderr, "zen_b200: zen_io_alloc failed: %s\n", cudaGetErrorString(e));
		zen_io_free(io);  // nothing leaks on a partial failure
		return ZEN_ERR_CUDA;
	}
	std::memset(io->host_in, 0, size * sizeof(float));
	std::memset(io->host_out, 0, size * sizeof(float));
	return ZEN_OK;
}

void zen_io_free(zen_io* io)
{
	if (!io)
		return;
	if (io->host_in) cudaFreeHost(io->host_in);
	if (io->host_out) cudaFreeHost(io->host_out);
	std::memset(io, 0, sizeof(*io));
}

// pinned (page-locked) host memory for the host-buffer entry points
void* zen_host_alloc(size_t bytes)
{
	void* p = nullptr;
	if (cudaHostAlloc(&p, bytes, cudaHostAllocPortable) != cudaSuccess) {
		cudaGetLastError();
		return nullptr;
	}
	return p;
}

void zen_host_free(void* p)
{
	if (p) cudaFreeHost(p);
}

// plain synchronous copies for hosts that have no CUDA runtime binding of their own
int zen_copy_to_host(void* h_dst, const void* d_src, size_t bytes)
{
	ZEN_CUDA_CHECK(cudaMemcpy(h_dst, d_src, bytes, cudaMemcpyDeviceToHost));
	return ZEN_OK;
}

int zen_copy_to_device(void* d_dst, const void* h_src, size_t bytes)
{
	ZEN_CUDA_CHECK(cudaMemcpy(d_dst, h_src, bytes, cudaMemcpyHostToDevice));
	return ZEN_OK;
}

}  // extern "C"
