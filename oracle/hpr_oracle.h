/* TEST INFRASTRUCTURE — CPU restatement (plain C) of the reference's HPR hot
 * path.  It is the checker for tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / reference arm.  Nothing under zen_b200/ may include, link or
 * load it; the product path fails loudly without its CUDA library instead.
 *
 * Pinned against: the reference's own known-answer tests (mfilt.test.cu,
 * box.test.cu, restated in tests/test_oracle_golden.py) and against outputs of
 * the UNMODIFIED reference compiled from /root/reference (oracle/_ref,
 * fixtures under tests/golden/ with the generating script
 * oracle/ref/probe_ref_gpu.py).
 *
 * geometry flag: ZO_GEOM_GPU restates the NPP path (wrap border / shrunken
 * ROI, libzen/mfilt.h:93-267, box.h:84-214); ZO_GEOM_CPU restates the IPP path
 * (centred window, replicated border, libzen/mfilt.h:285-341, box.h:232-287).
 */
#ifndef ZEN_HPR_ORACLE_H
#define ZEN_HPR_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

enum { ZO_GEOM_GPU = 0, ZO_GEOM_CPU = 1 };
enum { ZO_TIME_CAUSAL = 0, ZO_TIME_ANTICAUSAL = 1, ZO_FREQUENCY = 2 }; /* mfilt.h:27-31 */
enum { ZO_OUT_HARMONIC = 1, ZO_OUT_PERCUSSIVE = 2, ZO_OUT_RESIDUAL = 4 }; /* libzen/hps.h:25-27 */
enum { ZO_WIN_SQRT_HANN = 0, ZO_WIN_HANN = 1 }; /* win.h:13-16 */

typedef struct {
	int hop, nwin, nfft, l_harm, l_perc, lag, stft_width;
	float cola;
} zo_geom;

/* hps.h:216-285: derived sizes and COLA factor */
void zo_geometry(float fs, int hop, int causal, zo_geom* g);
/* win.h:21-53 */
void zo_window(int type, int n, float* out);

/* mfilt.h MedianFilterGPU/CPU::filter on a time x freq row-major matrix.
 * dst is in/out (cells the reference leaves untouched stay as they were).
 * Returns 0, or -1 where the reference throws ZgException. */
int zo_median_filter(int geom, int time, int freq, int filter_len, int dir, int copy_bord,
                     const float* src, float* dst);
/* box.h BoxFilterGPU/CPU::filter (arithmetic mean; GPU always wrap-padded) */
int zo_box_filter(int geom, int time, int freq, int filter_len, int dir, const float* src, float* dst);
/* fftw.h: in-place complex-to-complex, unnormalised both ways; dir 0 fwd 1 inv.
 * Evaluated in double and rounded once to float. */
void zo_fft(int nfft, float* interleaved, int dir);

/* hps.h HPR<B>: streaming state machine */
typedef struct zo_hpr zo_hpr;
zo_hpr* zo_hpr_create(int geom, float fs, int hop, float beta, unsigned flags, int causality, int copy_bord);
void zo_hpr_destroy(zo_hpr*);
void zo_hpr_use_sse_filter(zo_hpr*);
void zo_hpr_use_soft_mask(zo_hpr*);
void zo_hpr_reset_buffers(zo_hpr*);
void zo_hpr_geom(const zo_hpr*, zo_geom*);
/* hps.cu:429-652 */
void zo_hpr_process_next_hop(zo_hpr*, const float* in_hop);
/* which: 0 s_mag 1 harmonic_matrix 2 percussive_matrix 3 harmonic_mask 4 percussive_mask
 * 5 residual_mask 6 harmonic_out 7 percussive_out 8 residual_out 9 reciprocal 10 input
 * 11 window 12 sliding_stft (interleaved).  Returns element count (floats). */
int zo_hpr_get(const zo_hpr*, int which, float* out);
/* run n_hops, collecting the first hop samples of each enabled output per hop */
void zo_hpr_run(zo_hpr*, const float* audio, int n_hops, float* h_out, float* p_out, float* r_out);

/* hps.cu:109-280 HPRIOffline<B>::process; flags bit0 sse, bit1 soft mask.
 * geom GPU returns {harmonic(pass 1), percussive(pass 2), zeros};
 * geom CPU returns {percussive, percussive, percussive} as the reference does.
 * Returns 0, or -1 where the reference throws ZgException. */
int zo_offline_process(int geom, float fs, int hop_h, int hop_p, float beta_h, float beta_p,
                       int nocopybord, int flags, const float* audio, long n,
                       float* h_out, float* p_out, float* r_out);

/* demos/pitch-tracking/pitch.cpp:40-135 MPM::pitch (the consumer of the harmonic output, main.cu:90-107): n samples in,
 * pitch in Hz or -1; nsdf_out (optional, n floats) receives real_autocorrelation's output */
float zo_mpm_pitch(const float* audio, int n, float sample_rate, float* nsdf_out);

/* demos/beat-tracking/OnsetDetection.cpp:60-131 (complex spectral difference, half-wave rectified; the consumer of the
 * percussive output, main.cu:92-118): n_hops hops of 256 samples in, one ODF sample per hop out */
void zo_onset_csd(const float* audio, long n_hops, float* out);
void zo_onset_window(float* w512); /* Window.h:31-40, Hann, 512 */

/* zen/fakert.h:15-34 get_chunk_limits: number of hops fakert processes */
long zo_fakert_n_chunks(long size, long hop);

#ifdef __cplusplus
}
#endif
#endif
