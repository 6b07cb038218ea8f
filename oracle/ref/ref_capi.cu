/* TEST INFRASTRUCTURE — flat C entry points around the UNMODIFIED reference
 * classes, compiled together with /root/reference/libzen/hps.cu and core.cu
 * into oracle/_ref/libzen_ref.so (see oracle/Makefile).  Used only by tests/,
 * __graft_entry__.smoke() and bench.py's reference arm as the checker / the
 * baseline; nothing under zen_b200/ links or loads it.
 *
 * Backend codes: 0 = zen::Backend::GPU (thrust + cuFFT + NPP), 1 = CPU (the
 * reference's CPU dataflow; its IPP calls are served by oracle/ref/ippstub). */
#include <chrono>
#include <cstdio>
#include <cstring>
#include <vector>

#include <cuda_runtime.h>
#include <thrust/copy.h>
#include <thrust/device_vector.h>

#include <box.h>
#include <fftw.h>
#include <hps.h>
#include <libzen/hps.h>
#include <libzen/io.h>
#include <libzen/zen.h>
#include <mfilt.h>
#include <win.h>

using zen::Backend;
using zen::internal::hps::HPR;
using zen::internal::hps::mfilt::MedianFilterDirection;
namespace mf = zen::internal::hps::mfilt;
namespace bx = zen::internal::hps::box;

namespace {

struct RefHPR {
	int backend;
	HPR<Backend::GPU>* g;
	HPR<Backend::CPU>* c;
	zen::io::IOGPU* io;
};

template <typename V>
void dump_real(const V& v, float* out)
{
	thrust::copy(v.begin(), v.end(), out);
}

template <typename H>
int get_field(H* h, int which, float* out)
{
	switch (which) {
	case 0: dump_real(h->s_mag, out); return (int)h->s_mag.size();
	case 1: dump_real(h->harmonic_matrix, out); return (int)h->harmonic_matrix.size();
	case 2: dump_real(h->percussive_matrix, out); return (int)h->percussive_matrix.size();
	case 3: dump_real(h->harmonic_mask, out); return (int)h->harmonic_mask.size();
	case 4: dump_real(h->percussive_mask, out); return (int)h->percussive_mask.size();
	case 5: dump_real(h->residual_mask, out); return (int)h->residual_mask.size();
	case 6: dump_real(h->harmonic_out, out); return (int)h->harmonic_out.size();
	case 7: dump_real(h->percussive_out, out); return (int)h->percussive_out.size();
	case 8: dump_real(h->residual_out, out); return (int)h->residual_out.size();
	case 9: dump_real(h->reciprocal, out); return (int)h->reciprocal.size();
	case 10: dump_real(h->input, out); return (int)h->input.size();
	case 11: dump_real(h->window.window, out); return (int)h->window.window.size();
	case 12: {
		std::vector<thrust::complex<float>> tmp(h->sliding_stft.size());
		thrust::copy(h->sliding_stft.begin(), h->sliding_stft.end(), tmp.begin());
		std::memcpy(out, tmp.data(), tmp.size() * sizeof(float) * 2);
		return (int)tmp.size() * 2;
	}
	}
	return -1;
}

template <typename H>
void get_geom(H* h, double* g)
{
	g[0] = h->fs;
	g[1] = (double)h->hop;
	g[2] = (double)h->nwin;
	g[3] = (double)h->nfft;
	g[4] = h->beta;
	g[5] = h->l_harm;
	g[6] = h->l_perc;
	g[7] = h->lag;
	g[8] = (double)h->stft_width;
	g[9] = h->COLA_factor;
}

/* NPP/cuFFT calls can leave a non-sticky error behind that the reference never
 * reads (it ignores their status codes); thrust's next kernel
 * launch would pick it up via cudaPeekAtLastError and throw.  The standalone
 * entry points below therefore report-and-clear it on exit. */
void clear_stale_error(const char* where)
{
	cudaError_t e = cudaGetLastError();
	if (e != cudaSuccess)
		std::fprintf(stderr, "ref_capi: stale CUDA error after %s: %s (cleared)\n", where, cudaGetErrorName(e));
}

} // namespace

extern "C" {

int ref_device_count()
{
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) {
		cudaGetLastError();
		return 0;
	}
	return n;
}

/* causality: 0 TimeCausal, 1 TimeAnticausal.  Returns NULL on ZgException. */
void* ref_hpr_create(int backend, float fs, int hop, float beta, unsigned flags,
                     int causality, int copy_bord)
{
	RefHPR* r = new RefHPR{backend, nullptr, nullptr, nullptr};
	auto dir = causality == 0 ? MedianFilterDirection::TimeCausal
	                          : MedianFilterDirection::TimeAnticausal;
	try {
		if (backend == 0) {
			r->g = new HPR<Backend::GPU>(fs, hop, beta, flags, dir, copy_bord != 0);
			r->io = new zen::io::IOGPU(hop);
		}
		else {
			r->c = new HPR<Backend::CPU>(fs, hop, beta, flags, dir, copy_bord != 0);
		}
	}
	catch (const zen::ZgException& e) {
		delete r;
		return nullptr;
	}
	return r;
}

void ref_hpr_destroy(void* p)
{
	RefHPR* r = (RefHPR*)p;
	delete r->g;
	delete r->c;
	delete r->io;
	delete r;
	/* ~IOGPU calls cudaFree on cudaHostAlloc memory (io.h:72-76) and
	 * ~MedianFilterGPU calls nppsFree on nppiMalloc memory (mfilt.h:218-225);
	 * the error they leave behind would make the next object's first thrust
	 * kernel throw ("invalid device ordinal", observed on the B200 box). */
	clear_stale_error("HPR/IOGPU destructors");
}

void ref_hpr_use_sse(void* p)
{
	RefHPR* r = (RefHPR*)p;
	if (r->g) r->g->use_sse_filter(); else r->c->use_sse_filter();
}

void ref_hpr_use_soft(void* p)
{
	RefHPR* r = (RefHPR*)p;
	if (r->g) r->g->use_soft_mask(); else r->c->use_soft_mask();
}

void ref_hpr_reset(void* p)
{
	RefHPR* r = (RefHPR*)p;
	if (r->g) r->g->reset_buffers(); else r->c->reset_buffers();
}

void ref_hpr_geometry(void* p, double* g10)
{
	RefHPR* r = (RefHPR*)p;
	if (r->g) get_geom(r->g, g10); else get_geom(r->c, g10);
}

/* one hop exactly as the callers drive it: host samples -> IOGPU.host_in ->
 * process_next_hop(io.device_in)   (zen/fakert.h:225-229) */
void ref_hpr_process_next_hop(void* p, const float* hop_in)
{
	RefHPR* r = (RefHPR*)p;
	if (r->g) {
		std::copy(hop_in, hop_in + r->g->hop, r->io->host_in);
		r->g->process_next_hop(r->io->device_in);
	}
	else {
		r->c->process_next_hop(const_cast<float*>(hop_in));
	}
}

int ref_hpr_get(void* p, int which, float* out)
{
	RefHPR* r = (RefHPR*)p;
	return r->g ? get_field(r->g, which, out) : get_field(r->c, which, out);
}

/* run n_hops and collect the first `hop` samples of each output per hop
 * (what copy_harmonic/percussive/residual hand out, hps.cu:341-390).
 * outs may be NULL individually. */
void ref_hpr_run(void* p, const float* audio, int n_hops, float* h_out, float* p_out, float* r_out)
{
	RefHPR* r = (RefHPR*)p;
	size_t hop = r->g ? r->g->hop : r->c->hop;
	std::vector<float> tmp(2 * hop);
	for (int i = 0; i < n_hops; ++i) {
		ref_hpr_process_next_hop(p, audio + (size_t)i * hop);
		if (h_out) { ref_hpr_get(p, 6, tmp.data()); std::memcpy(h_out + (size_t)i * hop, tmp.data(), hop * sizeof(float)); }
		if (p_out) { ref_hpr_get(p, 7, tmp.data()); std::memcpy(p_out + (size_t)i * hop, tmp.data(), hop * sizeof(float)); }
		if (r_out) { ref_hpr_get(p, 8, tmp.data()); std::memcpy(r_out + (size_t)i * hop, tmp.data(), hop * sizeof(float)); }
	}
}

/* the region zen/fakert.h:221-247 times, per hop, in microseconds:
 * host copy-in -> process_next_hop -> copy_percussive -> host copy-out.
 * flags: bit0 sse, bit1 soft mask.  Returns 0, or -1 on ZgException. */
int ref_fakert_latency(int backend, float fs, int hop, float beta, int nocopybord, int flags,
                       const float* audio, int n_hops, int warm, float* perc_out, double* us_per_hop)
{
	try {
		if (backend == 0) {
			auto hpss = zen::hps::HPRRealtime<Backend::GPU>(fs, hop, beta, zen::hps::OUTPUT_PERCUSSIVE, nocopybord != 0);
			zen::io::IOGPU io(hop);
			if (flags & 1) hpss.use_sse_filter();
			if (flags & 2) hpss.use_soft_mask();
			if (warm) hpss.warmup(io);
			for (int i = 0; i < n_hops; ++i) {
				auto t1 = std::chrono::high_resolution_clock::now();
				std::copy(audio + (size_t)i * hop, audio + (size_t)(i + 1) * hop, io.host_in);
				hpss.process_next_hop(io.device_in);
				hpss.copy_percussive(io.device_out);
				std::copy(io.host_out, io.host_out + hop, perc_out + (size_t)i * hop);
				auto t2 = std::chrono::high_resolution_clock::now();
				us_per_hop[i] = std::chrono::duration<double, std::micro>(t2 - t1).count();
			}
		}
		else {
			auto hpss = zen::hps::HPRRealtime<Backend::CPU>(fs, hop, beta, zen::hps::OUTPUT_PERCUSSIVE);
			if (flags & 1) hpss.use_sse_filter();
			if (flags & 2) hpss.use_soft_mask();
			if (warm) hpss.warmup();
			for (int i = 0; i < n_hops; ++i) {
				auto t1 = std::chrono::high_resolution_clock::now();
				hpss.process_next_hop(const_cast<float*>(audio + (size_t)i * hop));
				hpss.copy_percussive(perc_out + (size_t)i * hop);
				auto t2 = std::chrono::high_resolution_clock::now();
				us_per_hop[i] = std::chrono::duration<double, std::micro>(t2 - t1).count();
			}
		}
	}
	catch (const zen::ZgException&) {
		return -1;
	}
	clear_stale_error("HPRRealtime/IOGPU destructors");
	return 0;
}

/* HPRIOffline<B>::process (hps.cu:128-280).  flags: bit0 sse, bit1 soft.
 * Returns elapsed ms of process() (the region zen/offline.h:165-167 times),
 * or -1 on ZgException. */
double ref_offline_process(int backend, float fs, int hop_h, int hop_p, float beta_h, float beta_p,
                           int nocopybord, int flags, const float* audio, long n,
                           float* h_out, float* p_out, float* r_out)
{
	try {
		std::vector<float> in(audio, audio + n);
		std::array<std::vector<float>, 3> res;
		double ms;
		clear_stale_error("before HPRIOffline");
		if (backend == 0) {
			auto hpss = zen::hps::HPRIOffline<Backend::GPU>(fs, hop_h, hop_p, beta_h, beta_p, nocopybord != 0);
			if (flags & 1) hpss.use_sse_filter();
			if (flags & 2) hpss.use_soft_mask();
			auto t1 = std::chrono::high_resolution_clock::now();
			res = hpss.process(in);
			auto t2 = std::chrono::high_resolution_clock::now();
			ms = std::chrono::duration<double, std::milli>(t2 - t1).count();
		}
		else {
			auto hpss = zen::hps::HPRIOffline<Backend::CPU>(fs, hop_h, hop_p, beta_h, beta_p, nocopybord != 0);
			if (flags & 1) hpss.use_sse_filter();
			if (flags & 2) hpss.use_soft_mask();
			auto t1 = std::chrono::high_resolution_clock::now();
			res = hpss.process(in);
			auto t2 = std::chrono::high_resolution_clock::now();
			ms = std::chrono::duration<double, std::milli>(t2 - t1).count();
		}
		std::memcpy(h_out, res[0].data(), n * sizeof(float));
		std::memcpy(p_out, res[1].data(), n * sizeof(float));
		std::memcpy(r_out, res[2].data(), n * sizeof(float));
		clear_stale_error("HPRIOffline destructors");
		return ms;
	}
	catch (const zen::ZgException&) {
		return -1.0;
	}
}

/* MedianFilter{GPU,CPU}::filter on a time x freq matrix; dst is in/out so the
 * caller can pre-fill a sentinel and see which cells NPP leaves untouched.
 * dir: 0 TimeCausal, 1 TimeAnticausal, 2 Frequency.  -1 on ZgException. */
int ref_median_filter(int backend, int time, int freq, int filter_len, int dir, int copy_bord,
                      const float* src, float* dst)
{
	try {
		size_t n = (size_t)time * freq;
		if (backend == 0) {
			mf::MedianFilterGPU f(time, freq, filter_len, (MedianFilterDirection)dir, copy_bord != 0);
			thrust::device_vector<float> s(src, src + n), d(dst, dst + n);
			f.filter(s, d);
			cudaDeviceSynchronize();
			thrust::copy(d.begin(), d.end(), dst);
			clear_stale_error("MedianFilterGPU");
		}
		else {
			mf::MedianFilterCPU f(time, freq, filter_len, (MedianFilterDirection)dir, copy_bord != 0);
			std::vector<float> s(src, src + n), d(dst, dst + n);
			f.filter(s, d);
			std::memcpy(dst, d.data(), n * sizeof(float));
		}
	}
	catch (const zen::ZgException&) {
		return -1;
	}
	return 0;
}

int ref_box_filter(int backend, int time, int freq, int filter_len, int dir, const float* src, float* dst)
{
	try {
		size_t n = (size_t)time * freq;
		if (backend == 0) {
			bx::BoxFilterGPU f(time, freq, filter_len, (MedianFilterDirection)dir);
			thrust::device_vector<float> s(src, src + n), d(dst, dst + n);
			f.filter(s, d);
			cudaDeviceSynchronize();
			thrust::copy(d.begin(), d.end(), dst);
			clear_stale_error("BoxFilterGPU");
		}
		else {
			bx::BoxFilterCPU f(time, freq, filter_len, (MedianFilterDirection)dir);
			std::vector<float> s(src, src + n), d(dst, dst + n);
			f.filter(s, d);
			std::memcpy(dst, d.data(), n * sizeof(float));
		}
	}
	catch (const zen::ZgException&) {
		return -1;
	}
	return 0;
}

/* timed median filter on a device-resident matrix (mfilt.bench.cu "NOMEM"
 * variant): returns mean microseconds per filter() call over `iters`. */
double ref_median_filter_time(int time, int freq, int filter_len, int dir, int copy_bord, int iters)
{
	size_t n = (size_t)time * freq;
	mf::MedianFilterGPU f(time, freq, filter_len, (MedianFilterDirection)dir, copy_bord != 0);
	std::vector<float> h(n);
	unsigned s = 12345u;
	for (size_t i = 0; i < n; ++i) { s = s * 1664525u + 1013904223u; h[i] = (float)(s >> 8) / 16777216.0f; }
	thrust::device_vector<float> src(h.begin(), h.end()), dst(n, 0.0f);
	f.filter(src, dst);
	cudaDeviceSynchronize();
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0);
	cudaEventCreate(&e1);
	cudaEventRecord(e0);
	for (int i = 0; i < iters; ++i) f.filter(src, dst);
	cudaEventRecord(e1);
	cudaEventSynchronize(e1);
	float ms = 0;
	cudaEventElapsedTime(&ms, e0, e1);
	cudaEventDestroy(e0);
	cudaEventDestroy(e1);
	return (double)ms * 1000.0 / iters;
}

/* FFTC2CWrapper{GPU,CPU}: in-place on nfft interleaved complex; dir 0 fwd, 1 inv */
void ref_fft(int backend, int nfft, float* inout, int dir)
{
	if (backend == 0) {
		zen::internal::fftw::FFTC2CWrapperGPU f(nfft);
		std::vector<thrust::complex<float>> h(nfft);
		std::memcpy(h.data(), inout, sizeof(float) * 2 * nfft);
		thrust::copy(h.begin(), h.end(), f.fft_vec.begin());
		if (dir == 0) f.forward(); else f.backward();
		cudaDeviceSynchronize();
		thrust::copy(f.fft_vec.begin(), f.fft_vec.end(), h.begin());
		std::memcpy(inout, h.data(), sizeof(float) * 2 * nfft);
		clear_stale_error("FFTC2CWrapperGPU");
	}
	else {
		zen::internal::fftw::FFTC2CWrapperCPU f(nfft);
		std::memcpy(f.fft_vec.data(), inout, sizeof(float) * 2 * nfft);
		if (dir == 0) f.forward(); else f.backward();
		std::memcpy(inout, f.fft_vec.data(), sizeof(float) * 2 * nfft);
	}
}

/* Window<std::vector<float>> (win.h:21-53): type 0 SqrtVonHann, 1 VonHann */
void ref_window(int type, int n, float* out)
{
	zen::internal::win::WindowCPU w((zen::internal::win::WindowType)type, n);
	std::memcpy(out, w.window.data(), n * sizeof(float));
}

} /* extern "C" */

/* bring-up aid: run one thrust kernel and report the CUDA error state */
extern "C" int ref_debug_thrust()
{
	int dev = -1, cnt = -1;
	cudaError_t e0 = cudaGetDeviceCount(&cnt);
	cudaError_t e1 = cudaGetDevice(&dev);
	std::printf("ref_debug: count=%d (%s) dev=%d (%s) last=%s\n", cnt, cudaGetErrorName(e0), dev, cudaGetErrorName(e1),
	            cudaGetErrorName(cudaPeekAtLastError()));
	try {
		thrust::device_vector<float> v(1000, 1.0f);
		thrust::fill(v.begin(), v.end(), 2.0f);
		float x = v[10];
		std::printf("ref_debug: fill ok, v[10]=%g last=%s\n", x, cudaGetErrorName(cudaPeekAtLastError()));
	}
	catch (const std::exception& e) {
		std::printf("ref_debug: thrust threw: %s ; last=%s\n", e.what(), cudaGetErrorName(cudaGetLastError()));
		return 1;
	}
	return 0;
}

#define REF_DBG_STEP(what) std::printf("ref_debug: after %-28s last=%s\n", what, cudaGetErrorName(cudaGetLastError()))
extern "C" int ref_debug_steps()
{
	REF_DBG_STEP("start");
	{
		zen::internal::win::WindowGPU w(zen::internal::win::WindowType::SqrtVonHann, 512);
		REF_DBG_STEP("WindowGPU");
	}
	{
		mf::MedianFilterGPU f(6, 4096, 3, MedianFilterDirection::TimeCausal, true);
		REF_DBG_STEP("MedianFilterGPU ctor");
		thrust::device_vector<float> s(6 * 4096, 1.0f), d(6 * 4096, 0.0f);
		REF_DBG_STEP("device_vector fill");
		f.filter(s, d);
		cudaDeviceSynchronize();
		REF_DBG_STEP("MedianFilterGPU filter");
	}
	REF_DBG_STEP("MedianFilterGPU dtor");
	{
		bx::BoxFilterGPU f(6, 4096, 3, MedianFilterDirection::TimeCausal);
		REF_DBG_STEP("BoxFilterGPU ctor");
	}
	{
		zen::internal::fftw::FFTC2CWrapperGPU f(4096);
		REF_DBG_STEP("FFTC2CWrapperGPU ctor");
		f.forward();
		cudaDeviceSynchronize();
		REF_DBG_STEP("FFT forward");
	}
	REF_DBG_STEP("FFT dtor");
	{
		zen::io::IOGPU io(1024);
		REF_DBG_STEP("IOGPU ctor");
	}
	REF_DBG_STEP("IOGPU dtor");
	{
		HPR<Backend::GPU> h(44100.0f, 1024, 2.5f, 7, MedianFilterDirection::TimeCausal, true);
		REF_DBG_STEP("HPR ctor");
	}
	REF_DBG_STEP("HPR dtor");
	return 0;
}

#define REF_TRY(label, ...)                                                                              \
	try {                                                                                               \
		__VA_ARGS__;                                                                                     \
		cudaError_t e_ = cudaDeviceSynchronize();                                                       \
		std::printf("ref_debug: %-34s ok   sync=%s last=%s\n", label, cudaGetErrorName(e_),              \
		            cudaGetErrorName(cudaGetLastError()));                                              \
	}                                                                                                   \
	catch (const std::exception& ex) {                                                                  \
		std::printf("ref_debug: %-34s THREW %s ; last=%s\n", label, ex.what(),                           \
		            cudaGetErrorName(cudaGetLastError()));                                              \
	}
extern "C" int ref_debug_hop()
{
	int hop = 1024;
	HPR<Backend::GPU> h(44100.0f, hop, 2.5f, 7, MedianFilterDirection::TimeCausal, true);
	zen::io::IOGPU io(hop);
	for (int i = 0; i < hop; ++i) io.host_in[i] = 0.001f * i;
	auto in_hop = io.device_in;
	std::fflush(stdout);
	REF_TRY("K1 copy out shift", thrust::copy(h.percussive_out.begin() + hop, h.percussive_out.end(), h.percussive_out.begin()));
	REF_TRY("K1 fill", thrust::fill(h.percussive_out.begin() + hop, h.percussive_out.end(), 0.0));
	REF_TRY("K2 copy input shift", thrust::copy(h.input.begin() + hop, h.input.end(), h.input.begin()));
	REF_TRY("K2 copy mapped->input", thrust::copy(in_hop, in_hop + hop, h.input.begin() + hop));
	REF_TRY("K3 window transform", thrust::transform(h.input.begin(), h.input.end(), h.window.window.begin(), h.fft.fft_vec.begin(), zen::internal::hps::window_functor()));
	REF_TRY("K4 fill complex", thrust::fill(h.fft.fft_vec.begin() + h.nwin, h.fft.fft_vec.end(), thrust::complex<float>{0.0, 0.0}));
	REF_TRY("K5 fft fwd", h.fft.forward());
	REF_TRY("K6 stft shift", thrust::copy(h.sliding_stft.begin() + h.nfft, h.sliding_stft.end(), h.sliding_stft.begin()));
	REF_TRY("K7 stft append", thrust::copy(h.fft.fft_vec.begin(), h.fft.fft_vec.end(), h.sliding_stft.end() - h.nfft));
	REF_TRY("K8 abs", thrust::transform(h.sliding_stft.begin(), h.sliding_stft.end(), h.s_mag.begin(), zen::internal::hps::complex_abs_functor()));
	REF_TRY("K10 time median", h.time.filter(h.s_mag, h.harmonic_matrix));
	REF_TRY("K12 freq median", h.frequency.filter(h.s_mag, h.percussive_matrix));
	REF_TRY("whole process_next_hop", h.process_next_hop(in_hop));
	REF_TRY("whole process_next_hop (2)", h.process_next_hop(in_hop));
	std::fflush(stdout);
	return 0;
}
