/* zen_b200 — C ABI of the B200-native HPR hot path (libzen_b200.so).
 *
 * The reference (sevagh/Zen) has no C ABI: its boundary is the C++ API of
 * libzen.  The C++ headers under zen_b200/include/
 * reproduce that API by name and forward to the entry points below; each entry
 * point cites the reference interface it replaces.  Plain pointers and sizes
 * only — no thrust / torch types.
 *
 * Pointers named d_* must be device-accessible (cudaMalloc memory, or the
 * device alias of mapped pinned host memory such as zen_io_alloc() hands out,
 * which is what the reference's callers pass: zen/fakert.h:229-230).
 * Pointers named h_* are ordinary host memory.
 *
 * All functions return ZEN_OK or a negative error code; nothing throws.
 * ZEN_ERR_GEOMETRY is returned wherever the reference throws zen::ZgException.
 * There is no CPU fallback: without a usable CUDA device every call that
 * computes returns ZEN_ERR_CUDA.
 */
#ifndef ZEN_B200_H
#define ZEN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
	ZEN_OK = 0,
	ZEN_ERR_GEOMETRY = -1, /* zen::ZgException in the reference */
	ZEN_ERR_CUDA = -2,
	ZEN_ERR_UNSUPPORTED = -3, /* size outside what the kernels are built for */
	ZEN_ERR_ARG = -4
};

/* libzen/mfilt.h:27-31 MedianFilterDirection */
enum { ZEN_TIME_CAUSAL = 0, ZEN_TIME_ANTICAUSAL = 1, ZEN_FREQUENCY = 2 };
/* libzen/libzen/hps.h:25-27 */
enum { ZEN_OUTPUT_HARMONIC = 1, ZEN_OUTPUT_PERCUSSIVE = 2, ZEN_OUTPUT_RESIDUAL = 4 };
/* libzen/win.h:13-16 */
enum { ZEN_WIN_SQRT_VON_HANN = 0, ZEN_WIN_VON_HANN = 1 };
/* option bits for the batched / offline entry points */
enum { ZEN_OPT_SSE = 1, ZEN_OPT_SOFT_MASK = 2, ZEN_OPT_NOCOPYBORD = 4 };

const char* zen_b200_version(void);
/* number of usable CUDA devices (0 => every compute call fails with ZEN_ERR_CUDA) */
int zen_device_count(void);

/* page-locked host memory (cudaHostAlloc / cudaFreeHost); NULL on failure */
void* zen_host_alloc(size_t bytes);
void zen_host_free(void* p);
/* synchronous cudaMemcpy wrappers for hosts without their own CUDA runtime binding */
int zen_copy_to_host(void* h_dst, const void* d_src, size_t bytes);
int zen_copy_to_device(void* d_dst, const void* h_src, size_t bytes);

/* ---- derived sizes: HPR<B>::HPR member initialisers, libzen/hps.h:216-285 ---- */
typedef struct {
	int hop, nwin, nfft, l_harm, l_perc, lag, stft_width;
	float cola_factor;
} zen_geometry;
int zen_hpr_geometry(float fs, int hop, int causal, zen_geometry* out);

/* ---- Window<T>, libzen/win.h:21-53 (computed on the host, as the reference does) ---- */
int zen_window(int type, int n, float* h_out);

/* ---- IOGPU, libzen/libzen/io.h:16-81: mapped pinned buffers + device aliases ---- */
typedef struct {
	float* host_in;
	float* host_out;
	float* device_in;
	float* device_out;
	size_t size;
} zen_io;
int zen_io_alloc(zen_io* io, size_t size);
void zen_io_free(zen_io* io);

/* ---- MedianFilterGPU, libzen/mfilt.h:33-268 ----
 * One filter pass over a time x freq row-major float matrix on the device.
 * Window / border rules are NPP's as driven by the reference (bit-exact):
 * copy_bord != 0 -> centred circular window along the filtered axis;
 * copy_bord == 0 -> shrunken ROI, cells outside it are left untouched.
 * filter_len is made odd as the reference does; ZEN_ERR_GEOMETRY if it exceeds
 * the filtered axis (checked before it is made odd). */
int zen_median_filter(int time, int freq, int filter_len, int direction, int copy_bord,
                      const float* d_src, float* d_dst, void* cuda_stream);
/* ---- BoxFilterGPU, libzen/box.h:30-215: wrap-padded moving average ---- */
int zen_box_filter(int time, int freq, int filter_len, int direction,
                   const float* d_src, float* d_dst, void* cuda_stream);

/* ---- FFTC2CWrapperGPU, libzen/fftw.h:20-49 ----
 * In-place 1-D complex-to-complex FFT on interleaved float pairs, unnormalised
 * in both directions; nfft a power of two in [2, 65536] (one CTA up to 4096,
 * a two-kernel four-step transform from 8192). */
int zen_fft_c2c(int nfft, float* d_inout, int inverse, void* cuda_stream);

/* ---- HPR<Backend::GPU>, libzen/hps.h:152-322 + libzen/hps.cu:429-652 ----
 * Streaming harmonic/percussive/residual separation, one hop per call.
 *
 * Limits of this build (everything else the reference accepts is accepted):
 *   - hop must be a power of two in [32, 4096] (nfft = 4 hop in [128, 16384], one transform per CTA);
 *     the reference's GPU path takes any hop cuFFT does, its IPP path powers of two only (libzen/fftw.h:57-60).
 *     zen_hpr_create / zen_offline_process* / zen_hpr_batch_* return ZEN_ERR_UNSUPPORTED otherwise.
 *   - the time-axis median reads at most ZEN_MAX_TAPS = 128 frames (csrc/hpr_core.cuh): l_harm, roughly
 *     0.2 fs / (3 hop) frames (libzen/hps.h:227), must not exceed 128, i.e. hop >= fs / 1920
 *     (fs 44.1 / 48 kHz: every supported hop; fs 192 kHz: hop >= 128).  ZEN_ERR_UNSUPPORTED otherwise.
 *   - the frequency-axis median window l_perc (500 nfft / fs bins, libzen/hps.h:229) must not exceed 255
 *     (fs >= 8 kHz at hop 1024).  ZEN_ERR_UNSUPPORTED otherwise. */
typedef struct zen_hpr zen_hpr;
int zen_hpr_create(zen_hpr** out, float fs, int hop, float beta, unsigned output_flags,
                   int causality /* ZEN_TIME_CAUSAL | ZEN_TIME_ANTICAUSAL */, int copy_bord);
void zen_hpr_destroy(zen_hpr* h);
int zen_hpr_use_sse_filter(zen_hpr* h); /* hps.h:287 */
int zen_hpr_use_soft_mask(zen_hpr* h);  /* hps.h:289 */
int zen_hpr_reset_buffers(zen_hpr* h);  /* hps.h:296-321 */
int zen_hpr_get_geometry(const zen_hpr* h, zen_geometry* out);
/* HPR<B>::process_next_hop (hps.cu:429-486): consumes hop samples at d_in_hop.
 * Asynchronous on the object's stream; results are ordered before any later
 * call on the same object. */
int zen_hpr_process_next_hop(zen_hpr* h, const float* d_in_hop);
/* HPRRealtime<GPU>::copy_{harmonic,percussive,residual} (hps.cu:341-363): the
 * first hop samples of the overlap-add buffer -> d_out_hop, then waits, so the
 * destination is readable by the host on return (zen/fakert.h:229-234). */
int zen_hpr_copy_harmonic(zen_hpr* h, float* d_out_hop);
int zen_hpr_copy_percussive(zen_hpr* h, float* d_out_hop);
int zen_hpr_copy_residual(zen_hpr* h, float* d_out_hop);
/* process_next_hop + copy_* in ONE launch: outputs whose pointer is NULL are
 * skipped.  This is the latency path (zen fakert's timed region). */
int zen_hpr_process_hop_io(zen_hpr* h, const float* d_in_hop, float* d_out_h, float* d_out_p, float* d_out_r);
int zen_hpr_synchronize(zen_hpr* h);
/* Returns once the hop passed to the last zen_hpr_process_next_hop has been read, i.e. the caller may refill the
 * buffer (the reference's process_next_hop has copied in_hop when it returns, hps.cu:452-453).  Free for a resident
 * session fed from mapped host memory (the hop was copied at submission); a stream synchronisation otherwise. */
int zen_hpr_wait_input_consumed(zen_hpr* h);
/* Resident real-time session for a causal stream: a persistent kernel keeps the
 * stream's state (|X| ring, overlap-add tails, previous hop, window and twiddle
 * tables) in shared memory and serves zen_hpr_process_next_hop /
 * zen_hpr_process_hop_io / zen_hpr_copy_* without a kernel launch, a stream
 * synchronisation or a system-wide fence per hop.  When the hop pointer and the
 * output pointers are host-visible (mapped pinned memory, zen_io_alloc) the hop
 * is PUSHED to the kernel as tagged 16-byte groups {x0, x1, x2, tag} and the
 * outputs come back the same way; device-memory pointers are read / written by
 * the kernel itself behind a completion flag.  With the default plan (hard mask,
 * copy-border) the hop is split over an 8-CTA thread-block cluster
 * (ZEN_B200_RT_CLUSTER=1|2|4|8).  Calls stay synchronous: on return the outputs
 * are readable by the host.  The kernel leaves by itself after
 * ZEN_B200_RT_IDLE_MS (default 250) without a hop and is brought back
 * transparently by the next call; every other entry point pauses it first.
 * While it is resident, device-wide synchronisation (cudaDeviceSynchronize)
 * waits for that idle time-out.  ZEN_B200_RT_PUSH=0 disables the tagged
 * transfers (the kernel then reads the hop itself: one more PCIe round trip). */
int zen_hpr_realtime_begin(zen_hpr* h);
int zen_hpr_realtime_end(zen_hpr* h);
/* The on-disk format either side of the path, batched on the device (device pointers; rows = independent streams).
 * zen_pcm16_decode_mono: PCM16 mono or interleaved stereo -> float32 mono exactly as libnyquist decodes it for
 * zen offline / fakert: (float)s / 32767.f (vendor/libnyquist/include/libnyquist/Common.h:296-302), stereo folded
 * with (l + r) / 2.0f (Common.h:669-675; zen/offline.h:104-117, zen/fakert.h:117-130).  Strides in elements.
 * zen_pcm16_encode_normalized: what the command line writes: x / max(-min, max) (zen/offline.h:180-192,
 * zen/fakert.h:259-268), then (int16_t)lroundf(x * 32767.f), no dither (vendor/libnyquist/src/Common.cpp:332-337);
 * d_peaks[n_streams] receives the peaks.  A silent stream (peak 0, where the reference divides by zero) stays silent. */
int zen_pcm16_decode_mono(const int16_t* d_pcm, long pcm_stride, int channels, int n_streams, long n_frames, float* d_out, long out_stride);
int zen_pcm16_encode_normalized(const float* d_in, long in_stride, int n_streams, long n, int16_t* d_out, long out_stride, float* d_peaks);
/* The same on a CUDA stream, without waiting (the two entry points above run on the legacy default stream and
 * synchronise).  zen_pcm16_peaks_async: d_peaks[s] = max |x| of row s; zen_pcm16_encode_with_peaks_async: the encode
 * pass alone, given the peaks; zen_pcm16_encode_normalized_async = both.  Rows whose pointers and strides are 16-byte
 * aligned take the vector path (8 samples per thread and access). */
int zen_pcm16_decode_mono_async(const int16_t* d_pcm, long pcm_stride, int channels, int n_streams, long n_frames, float* d_out,
                                long out_stride, void* cuda_stream);
int zen_pcm16_peaks_async(const float* d_in, long in_stride, int n_streams, long n, float* d_peaks, void* cuda_stream);
int zen_pcm16_encode_with_peaks_async(const float* d_in, long in_stride, int n_streams, long n, const float* d_peaks, int16_t* d_out,
                                      long out_stride, void* cuda_stream);
int zen_pcm16_encode_normalized_async(const float* d_in, long in_stride, int n_streams, long n, int16_t* d_out, long out_stride,
                                      float* d_peaks, void* cuda_stream);
/* host-only test hook: which half-spectrum bins CTA `rank` of a `cluster`-CTA resident kernel owns
 * (out6 = {k0, k1, a0, a1, b0, b1}: bin pairs (k, nfft/2 - k) for k in [k0, k1), i.e. bins [a0, a1) and [b0, b1)) */
int zen_rt_split_ranges(int nfft, int rank, int cluster, int* out6);
/* host-only test hooks for the tagged 16-byte groups {x[3g], x[3g+1], x[3g+2], tag} the resident session exchanges with
 * its kernel (no device involved).  groups: 16-byte aligned, ceil(hop / 3) * 16 bytes.  zen_rt_unpack_groups returns
 * the number of leading groups that carried `tag` and were unpacked. */
int zen_rt_pack_groups(const float* src, int hop, unsigned tag, void* groups);
int zen_rt_unpack_groups(const void* groups, int hop, unsigned tag, float* dst);
/* diagnostics (ZEN_B200_RT_STAMPS=1): SM cycle counter at the phase boundaries of the last hop the resident kernel
 * served, [9] / [12] = %globaltimer (ns) at its start / end */
int zen_hpr_realtime_stamps(zen_hpr* h, unsigned long long* out16);
int zen_hpr_realtime_stamps_rank1(zen_hpr* h, unsigned long long* out16); /* CTA 1 of the cluster; [9] = when it saw the command */
/* Make caller-owned device buffers (nwin floats each, 8-byte aligned) the object's
 * streaming state, so that e.g. thrust::device_vector members named like the
 * reference's (hps.h:182-197) ARE the state; zeroes them (reset_buffers). */
int zen_hpr_bind_state(zen_hpr* h, float* d_input, float* d_harmonic_out, float* d_percussive_out, float* d_residual_out);
/* Device pointers to the streaming state that the reference exposes as public
 * members of HPR<GPU> (hps.h:182-197): which = 0 input[nwin], 1 harmonic_out[nwin],
 * 2 percussive_out[nwin], 3 residual_out[nwin], 4 window[nwin]. */
float* zen_hpr_state_ptr(zen_hpr* h, int which);
/* Materialise the reference's full stft_width x nfft debug matrices from the
 * ring state (s_mag, harmonic_matrix, percussive_matrix, masks, sliding_stft as
 * interleaved complex), each into caller-provided device memory (NULL = skip).
 * Not on the per-hop path. */
int zen_hpr_materialize(zen_hpr* h, float* d_sliding_stft, float* d_s_mag, float* d_harmonic_matrix,
                        float* d_percussive_matrix, float* d_harmonic_mask, float* d_percussive_mask,
                        float* d_residual_mask);

/* ---- the real-time loop of `zen fakert` (zen/fakert.h:197-256) ----
 * HPRRealtime<GPU>(fs, hop, beta, OUTPUT_PERCUSSIVE) + IOGPU(hop); warmup_iters
 * hops of iota data then reset (hps.cu:392-409); then per hop the region the
 * reference times: host copy-in -> process_next_hop -> copy_percussive -> host
 * copy-out.  fused == 1 replaces the two calls by zen_hpr_process_hop_io;
 * fused == 2 additionally serves it from the resident kernel (zen_hpr_realtime_begin);
 * fused == 3 is the resident kernel behind the reference's own two calls (process_next_hop
 * is then only submitted and copy_percussive unpacks the output on the host).
 * h_us_per_hop (optional) receives the wall time of each hop in microseconds. */
int zen_fakert_run(float fs, int hop, float beta, int options, const float* h_audio, long n_hops,
                   int warmup_iters, int fused, float* h_perc_out, double* h_us_per_hop);

/* ---- batched streams: many independent HPRRealtime<GPU> streams at once ----
 * Equivalent to running, for every stream s, n_hops calls of
 * process_next_hop(in + s*in_stride + i*hop) on a fresh HPR object and
 * collecting the first hop samples of each enabled output after every call.
 * d_in: n_streams rows of n_hops*hop samples (row stride in_stride floats);
 * d_out_*: same layout (row stride out_stride), NULL for outputs not wanted.
 * causality ZEN_TIME_CAUSAL reproduces HPRRealtime, ZEN_TIME_ANTICAUSAL the
 * objects HPRIOffline drives.  options: ZEN_OPT_* bits.  max_streams / max_hops_per_stream are sizing hints
 * (1 is fine): scratch belongs to the resident CTAs and staging buffers grow on demand. */
typedef struct zen_hpr_batch zen_hpr_batch;
int zen_hpr_batch_create(zen_hpr_batch** out, float fs, int hop, float beta, unsigned output_flags,
                         int causality, int options, int max_streams, long max_hops_per_stream);
void zen_hpr_batch_destroy(zen_hpr_batch* b);
int zen_hpr_batch_process(zen_hpr_batch* b, const float* d_in, long in_stride, int n_streams, long n_hops,
                          float* d_out_h, float* d_out_p, float* d_out_r, long out_stride, void* cuda_stream);
/* same, host buffers (pinned or pageable): streams are cut into chunks that are
 * copied in, processed and copied out on alternating CUDA streams */
int zen_hpr_batch_process_host(zen_hpr_batch* b, const float* h_in, long in_stride, int n_streams, long n_hops,
                               float* h_out_h, float* h_out_p, float* h_out_r, long out_stride);
/* The same with the command line's on-disk sample format on both sides of the PCIe link (zen/offline.h:88-117,
 * 180-223; zen/fakert.h:101-130, 259-287): h_in holds mono PCM16 rows, decoded on the device exactly as libnyquist
 * does ((float)s / 32767.f, vendor/libnyquist/include/libnyquist/Common.h:296-302); every requested output comes
 * back peak-normalised per stream (x / max|x|) and converted with (int16_t)lroundf(x * 32767.f)
 * (vendor/libnyquist/src/Common.cpp:332-337) - what `zen offline|fakert` write to their wav files.  2 bytes per
 * sample cross the link in each direction instead of 4.  h_peaks_* (optional, n_streams floats each) receive the
 * peaks the outputs were divided by.  A silent stream stays silent (the reference divides by zero there). */
int zen_hpr_batch_process_host_pcm16(zen_hpr_batch* b, const int16_t* h_in, long in_stride, int n_streams, long n_hops,
                                     int16_t* h_out_h, int16_t* h_out_p, int16_t* h_out_r, long out_stride,
                                     float* h_peaks_h, float* h_peaks_p, float* h_peaks_r);
/* number of kernels launched by the last zen_hpr_batch_process* call */
long zen_hpr_batch_last_launches(const zen_hpr_batch* b);
/* device time of the fused kernel(s) of the last zen_hpr_batch_process call, ms (CUDA events) */
float zen_hpr_batch_last_kernel_ms(const zen_hpr_batch* b);

/* ---- HPRIOffline<GPU>::process, libzen/hps.cu:21-221 ----
 * Two-pass iterative HPR on a whole signal in host memory; returns
 * {harmonic (pass 1), percussive (pass 2), residual (all zero — the reference's
 * pass-2 object is built with OUTPUT_PERCUSSIVE only)} each of n samples.
 * ZEN_ERR_GEOMETRY if hop_h % hop_p != 0 (hps.cu:33-36). */
int zen_offline_process(float fs, int hop_h, int hop_p, float beta_h, float beta_p, int options,
                        const float* h_audio, long n, float* h_harmonic, float* h_percussive, float* h_residual);
/* same with device-resident input/outputs (no PCIe in the call) */
int zen_offline_process_device(float fs, int hop_h, int hop_p, float beta_h, float beta_p, int options,
                               const float* d_audio, long n, float* d_harmonic, float* d_percussive,
                               float* d_residual, void* cuda_stream);

/* ---- downstream consumer: McLeod pitch on the harmonic output (demos/pitch-tracking/pitch.cpp:40-135) ----
 * MPM(audio_buffer_size, sample_rate).pitch(buffer), batched and on the device: d_audio holds n_buffers buffers of n
 * samples (n = 256 ... 4096, a power of two; buffer b starts at d_audio + b * stride), e.g. the harmonic hops of
 * zen_hpr_batch_process / zen_hpr_copy_harmonic where the kernels left them (the reference's demo takes them through
 * io.host_out, main.cu:90-107).  d_pitch[b] = pitch in Hz, or -1 as the reference returns it; d_nsdf (optional,
 * n_buffers * n floats) receives real_autocorrelation's output. */
int zen_mpm_pitch(int n, float sample_rate, const float* d_audio, long stride, int n_buffers, float* d_pitch, float* d_nsdf,
                  void* cuda_stream);

/* ---- the consumer of the percussive output: BTrack's onset detection function, demos/beat-tracking/OnsetDetection.cpp:60-131
 * (complex spectral difference, half-wave rectified; 512-sample frames, 256-sample hops as OnsetDetection.h:15,27 fix them),
 * called per hop by BTrack::processHop (BTrack.cpp:93-98) on the hops main.cu:107-118 copies out of the GPU.  Here the hops
 * are read in device memory: row s of d_audio (stride in floats, >= 256 n_hops) is one stream from its start, row s of
 * d_odf (odf_stride >= n_hops) receives one sample per hop - what BTrack::processOnsetDetectionFunctionSample consumes. */
int zen_onset_csd(const float* d_audio, long stride, int n_streams, long n_hops, float* d_odf, long odf_stride, void* cuda_stream);

/* ---- the beat tracker behind that onset detection function: BTrack::processOnsetDetectionFunctionSample and what it
 * calls, demos/beat-tracking/BTrack.cpp:100-397 (cumulative score, beat prediction, tempo update once per beat).  Host
 * code, as in the reference: it consumes one float per 256-sample hop, e.g. a row of zen_onset_csd copied to the host.
 * zen_btrack_process feeds n consecutive samples; per sample it reports (each array optional) whether a beat is due in
 * that hop (beatDueInFrame), the tempo estimate in BPM (estimatedTempo) and the cumulative score
 * (latestCumulativeScoreValue).  The object keeps its state between calls (a stream can be fed in pieces). */
typedef struct zen_btrack zen_btrack;
int zen_btrack_create(zen_btrack** out, int sample_rate);
void zen_btrack_destroy(zen_btrack* b);
int zen_btrack_process(zen_btrack* b, const float* h_odf, long n, unsigned char* h_beat, float* h_tempo, float* h_cumscore);
/* its two lookup tables (BTrackPrecomputed.h, recomputed from their formulas): 128 and 41 x 41 floats */
int zen_btrack_tables(const zen_btrack* b, float* h_rayleigh128, float* h_transition41x41);

#ifdef __cplusplus
}
#endif
#endif /* ZEN_B200_H */
