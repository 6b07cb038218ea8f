/* TEST INFRASTRUCTURE — see hpr_oracle.h.  Plain C restatement of the
 * reference HPR path; every function cites the reference lines it follows.
 * Compile with -ffp-contract=off so float expressions are evaluated as written. */
#include "hpr_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define ZO_EPS FLT_EPSILON /* libzen/hps.h:22 */

/* ------------------------------------------------------------ geometry --- */

/* win.h:11 PI = 3.14159265359F; win.h:30-53 */
void zo_window(int type, int n, float* out)
{
	const float PI = 3.14159265359F;
	float N = (float)n;
	for (int i = 0; i < n; ++i) {
		float h = 0.5F * (1.0F - cosf(2.0F * PI * (float)i / N));
		out[i] = (type == ZO_WIN_SQRT_HANN) ? sqrtf(h) : h;
	}
}

/* hps.h:222-230 (member-initialiser expressions, float/double mix kept),
 * hps.h:265-274 (lag, COLA) */
void zo_geometry(float fs, int hop, int causal, zo_geom* g)
{
	g->hop = hop;
	g->nwin = 2 * hop;
	g->nfft = 4 * hop;
	g->l_harm = (int)roundf((float)(0.2 / ((float)(g->nfft - hop) / fs)));
	g->l_perc = (int)roundf(500 / (fs / (float)g->nfft));
	g->lag = causal ? 1 : g->l_harm;
	g->stft_width = 2 * g->l_harm;
	float* w = (float*)malloc(sizeof(float) * (size_t)g->nwin);
	zo_window(ZO_WIN_SQRT_HANN, g->nwin, w);
	float acc = 0.0f;
	for (int i = 0; i < g->nwin; ++i)
		acc += w[i] * w[i];
	g->cola = (float)g->nfft / acc;
	free(w);
}

/* ------------------------------------------------------- 1-D primitives --- */

static int cmp_float(const void* a, const void* b)
{
	float x = *(const float*)a, y = *(const float*)b;
	return (x > y) - (x < y);
}

/* out[k] = median(ext[k .. k+L)), L odd; window kept sorted while it slides */
static void sliding_median(const float* ext, int n_out, int L, float* out, float* sorted)
{
	if (n_out <= 0)
		return;
	if (L == 1) {
		memcpy(out, ext, sizeof(float) * (size_t)n_out);
		return;
	}
	memcpy(sorted, ext, sizeof(float) * (size_t)L);
	qsort(sorted, (size_t)L, sizeof(float), cmp_float);
	out[0] = sorted[L / 2];
	for (int k = 1; k < n_out; ++k) {
		float o = ext[k - 1], v = ext[k + L - 1];
		if (o != v) {
			int p = 0, q = 0, lo, hi;
			lo = 0; hi = L;
			while (lo < hi) { int m = (lo + hi) / 2; if (sorted[m] < o) lo = m + 1; else hi = m; }
			p = lo;
			lo = 0; hi = L;
			while (lo < hi) { int m = (lo + hi) / 2; if (sorted[m] < v) lo = m + 1; else hi = m; }
			q = lo;
			if (q > p) {
				memmove(sorted + p, sorted + p + 1, sizeof(float) * (size_t)(q - p - 1));
				sorted[q - 1] = v;
			}
			else {
				memmove(sorted + q + 1, sorted + q, sizeof(float) * (size_t)(p - q));
				sorted[q] = v;
			}
		}
		out[k] = sorted[L / 2];
	}
}

/* out[k] = mean(ext[k .. k+L)), each window summed in double */
static void sliding_mean(const float* ext, int n_out, int L, float* out)
{
	for (int k = 0; k < n_out; ++k) {
		double acc = 0.0;
		for (int j = 0; j < L; ++j)
			acc += (double)ext[k + j];
		out[k] = (float)(acc / (double)L);
	}
}

/* One filter pass over a time x freq matrix.  The window / border rules:
 *   GPU copy_bord  : nppiCopyWrapBorder (mfilt.h:246-255) pads `mid` wrapped
 *                    samples before and `mid+1` after along the filtered axis;
 *                    with the ROI offsets / anchors of mfilt.h:111-189 every
 *                    direction becomes the centred circular window
 *                    (i-mid .. i+mid) mod dim, all cells written;
 *   GPU !copy_bord : TimeCausal    rows r in [L,T)        window r-L .. r-1
 *                    TimeAnticausal rows r in [mid,mid+T-L) window r-mid .. r+mid
 *                    Frequency     cols c in [0,F-L)      window c .. c+L-1
 *                    (mfilt.h:111-158, 239-244), other cells untouched;
 *   CPU            : ippBorderRepl (mfilt.h:336-341): centred, clamped.
 * ZgException rule mfilt.h:80-87 uses the length before it is made odd. */
static int filter_pass(int geom, int is_box, int T, int F, int filter_len, int dir, int copy_bord,
                       const float* src, float* dst)
{
	if (((dir == ZO_TIME_CAUSAL || dir == ZO_TIME_ANTICAUSAL) && filter_len > T)
	    || (dir == ZO_FREQUENCY && filter_len > F))
		return -1;
	int L = filter_len + (1 - (filter_len % 2)); /* mfilt.h:89-91 */
	int mid = L / 2;
	int along_freq = (dir == ZO_FREQUENCY);
	int dim = along_freq ? F : T;
	int n_lines = along_freq ? T : F;
	int out0, n_out, ext0; /* outputs [out0,out0+n_out); window of output i = ext[i-out0 .. +L) */
	int mode;              /* 0 plain, 1 wrap, 2 clamp */
	if (geom == ZO_GEOM_CPU) {
		mode = 2; out0 = 0; n_out = dim; ext0 = -mid;
	}
	else if (copy_bord || is_box) {
		mode = 1; out0 = 0; n_out = dim; ext0 = -mid;
	}
	else if (dir == ZO_TIME_CAUSAL) {
		mode = 0; out0 = L; n_out = T - L; ext0 = 0;
	}
	else if (dir == ZO_TIME_ANTICAUSAL) {
		mode = 0; out0 = mid; n_out = T - L; ext0 = 0;
	}
	else {
		mode = 0; out0 = 0; n_out = F - L; ext0 = 0;
	}
	if (n_out <= 0)
		return 0;
	int n_ext = n_out + L - 1;
	float* ext = (float*)malloc(sizeof(float) * (size_t)(n_ext + 1));
	float* res = (float*)malloc(sizeof(float) * (size_t)n_out);
	float* sorted = (float*)malloc(sizeof(float) * (size_t)L);
	for (int line = 0; line < n_lines; ++line) {
		for (int j = 0; j < n_ext; ++j) {
			int idx = ext0 + j;
			if (mode == 1)
				idx = ((idx % dim) + dim) % dim;
			else if (mode == 2)
				idx = idx < 0 ? 0 : (idx > dim - 1 ? dim - 1 : idx);
			ext[j] = along_freq ? src[(size_t)line * F + idx] : src[(size_t)idx * F + line];
		}
		if (is_box)
			sliding_mean(ext, n_out, L, res);
		else
			sliding_median(ext, n_out, L, res, sorted);
		for (int k = 0; k < n_out; ++k) {
			int i = out0 + k;
			if (along_freq)
				dst[(size_t)line * F + i] = res[k];
			else
				dst[(size_t)i * F + line] = res[k];
		}
	}
	free(ext);
	free(res);
	free(sorted);
	return 0;
}

int zo_median_filter(int geom, int time, int freq, int filter_len, int dir, int copy_bord,
                     const float* src, float* dst)
{
	return filter_pass(geom, 0, time, freq, filter_len, dir, copy_bord, src, dst);
}

int zo_box_filter(int geom, int time, int freq, int filter_len, int dir, const float* src, float* dst)
{
	return filter_pass(geom, 1, time, freq, filter_len, dir, 1, src, dst);
}

/* twiddle table cos/sin(2*pi*j/n), j < n/2, built once per size */
static const double* twiddles(int n)
{
	static double* cache[32];
	int order = 0;
	while ((1 << order) < n)
		++order;
	if (!cache[order]) {
		const double two_pi = 6.283185307179586476925286766559;
		double* t = (double*)malloc(sizeof(double) * (size_t)(n > 1 ? n : 2));
		for (int j = 0; j < n / 2; ++j) {
			t[2 * j] = cos(two_pi * (double)j / (double)n);
			t[2 * j + 1] = sin(two_pi * (double)j / (double)n);
		}
		cache[order] = t;
	}
	return cache[order];
}

/* fftw.h:35-43 / 108-114: unnormalised C2C both ways.  Radix-2 DIT in double. */
static void fft_double(double* w, int n, int sign)
{
	int order = 0;
	while ((1 << order) < n)
		++order;
	const double* tw = twiddles(n);
	for (int i = 0; i < n; ++i) {
		unsigned r = 0;
		for (int b = 0; b < order; ++b)
			r |= ((unsigned)(i >> b) & 1u) << (order - 1 - b);
		if ((int)r > i) {
			double t0 = w[2 * i], t1 = w[2 * i + 1];
			w[2 * i] = w[2 * r];
			w[2 * i + 1] = w[2 * r + 1];
			w[2 * r] = t0;
			w[2 * r + 1] = t1;
		}
	}
	for (int len = 2; len <= n; len <<= 1) {
		int half = len >> 1, step = n / len;
		for (int base = 0; base < n; base += len) {
			for (int j = 0; j < half; ++j) {
				double c = tw[2 * j * step];
				double s = sign * tw[2 * j * step + 1];
				double* a = w + 2 * (base + j);
				double* b = w + 2 * (base + j + half);
				double tr = b[0] * c - b[1] * s;
				double ti = b[0] * s + b[1] * c;
				b[0] = a[0] - tr;
				b[1] = a[1] - ti;
				a[0] += tr;
				a[1] += ti;
			}
		}
	}
}

void zo_fft(int nfft, float* x, int dir)
{
	double* w = (double*)malloc(sizeof(double) * 2 * (size_t)nfft);
	for (int i = 0; i < 2 * nfft; ++i)
		w[i] = (double)x[i];
	fft_double(w, nfft, dir == 0 ? -1 : +1);
	for (int i = 0; i < 2 * nfft; ++i)
		x[i] = (float)w[i];
	free(w);
}

/* ----------------------------------------------------------------- HPR --- */

struct zo_hpr {
	int geom;
	float fs, beta;
	zo_geom g;
	int causality, copy_bord;
	int out_h, out_p, out_r, use_sse, soft_mask;
	float *input, *window;
	float* stft; /* interleaved complex, W*nfft */
	float *s_mag, *reciprocal, *hmat, *pmat, *pmask, *hmask, *rmask;
	float *pout, *hout, *rout;
	float* fft_vec; /* interleaved complex, nfft */
	double* fft_work;
};

static float* zalloc(size_t n) { return (float*)calloc(n ? n : 1, sizeof(float)); }

/* hps.h:216-285 */
zo_hpr* zo_hpr_create(int geom, float fs, int hop, float beta, unsigned flags, int causality, int copy_bord)
{
	zo_hpr* h = (zo_hpr*)calloc(1, sizeof(zo_hpr));
	h->geom = geom;
	h->fs = fs;
	h->beta = beta;
	h->causality = causality;
	h->copy_bord = copy_bord;
	zo_geometry(fs, hop, causality == ZO_TIME_CAUSAL, &h->g);
	int W = h->g.stft_width, nfft = h->g.nfft;
	/* the four filter constructors throw when the filter exceeds the axis
	 * (mfilt.h:80-87, box.h:71-78) */
	if (h->g.l_harm > W || h->g.l_perc > nfft || W <= 0) {
		free(h);
		return NULL;
	}
	size_t m = (size_t)W * nfft;
	h->input = zalloc(h->g.nwin);
	h->window = zalloc(h->g.nwin);
	zo_window(ZO_WIN_SQRT_HANN, h->g.nwin, h->window);
	h->stft = zalloc(2 * m);
	h->s_mag = zalloc(m);
	h->reciprocal = zalloc(m);
	h->hmat = zalloc(m);
	h->pmat = zalloc(m);
	h->pmask = zalloc(m);
	h->hmask = zalloc(m);
	h->rmask = zalloc(m);
	h->pout = zalloc(h->g.nwin);
	h->hout = zalloc(h->g.nwin);
	h->rout = zalloc(h->g.nwin);
	h->fft_vec = zalloc(2 * (size_t)nfft);
	h->fft_work = (double*)calloc(2 * (size_t)nfft, sizeof(double));
	h->out_h = (flags & ZO_OUT_HARMONIC) != 0;
	h->out_p = (flags & ZO_OUT_PERCUSSIVE) != 0;
	h->out_r = (flags & ZO_OUT_RESIDUAL) != 0;
	return h;
}

void zo_hpr_destroy(zo_hpr* h)
{
	if (!h)
		return;
	free(h->input); free(h->window); free(h->stft); free(h->s_mag); free(h->reciprocal);
	free(h->hmat); free(h->pmat); free(h->pmask); free(h->hmask); free(h->rmask);
	free(h->pout); free(h->hout); free(h->rout); free(h->fft_vec); free(h->fft_work);
	free(h);
}

void zo_hpr_use_sse_filter(zo_hpr* h) { h->use_sse = 1; } /* hps.h:287 */
void zo_hpr_use_soft_mask(zo_hpr* h) { h->soft_mask = 1; } /* hps.h:289 */
void zo_hpr_geom(const zo_hpr* h, zo_geom* g) { *g = h->g; }

/* hps.h:296-321 */
void zo_hpr_reset_buffers(zo_hpr* h)
{
	size_t m = (size_t)h->g.stft_width * h->g.nfft;
	memset(h->input, 0, sizeof(float) * h->g.nwin);
	memset(h->pout, 0, sizeof(float) * h->g.nwin);
	memset(h->hout, 0, sizeof(float) * h->g.nwin);
	memset(h->rout, 0, sizeof(float) * h->g.nwin);
	memset(h->fft_vec, 0, sizeof(float) * 2 * h->g.nfft);
	memset(h->stft, 0, sizeof(float) * 2 * m);
	memset(h->s_mag, 0, sizeof(float) * m);
	memset(h->reciprocal, 0, sizeof(float) * m);
	memset(h->hmat, 0, sizeof(float) * m);
	memset(h->pmat, 0, sizeof(float) * m);
	memset(h->hmask, 0, sizeof(float) * m);
	memset(h->pmask, 0, sizeof(float) * m);
	memset(h->rmask, 0, sizeof(float) * m);
}

static void hpr_fft(zo_hpr* h, int sign)
{
	int n = h->g.nfft;
	for (int i = 0; i < 2 * n; ++i)
		h->fft_work[i] = (double)h->fft_vec[i];
	fft_double(h->fft_work, n, sign);
	for (int i = 0; i < 2 * n; ++i)
		h->fft_vec[i] = (float)h->fft_work[i];
}

/* hps.cu:515-528: fft_vec = X[row]*mask[row]; inverse FFT; out += Re * COLA */
static void mask_ifft_ola(zo_hpr* h, const float* mask_row, float* out)
{
	int nfft = h->g.nfft, nwin = h->g.nwin;
	size_t off = (size_t)(h->g.stft_width - h->g.lag) * nfft;
	const float* X = h->stft + 2 * off;
	for (int k = 0; k < nfft; ++k) {
		h->fft_vec[2 * k] = X[2 * k] * mask_row[k];
		h->fft_vec[2 * k + 1] = X[2 * k + 1] * mask_row[k];
	}
	hpr_fft(h, +1);
	for (int i = 0; i < nwin; ++i)
		out[i] = out[i] + h->fft_vec[2 * i] * h->g.cola; /* hps.h:75-79 */
}

/* hps.h:100-113 */
static float hard_mask(float x, float y, float beta) { return (float)((x / (y + ZO_EPS)) >= beta); }
/* hps.h:116-129 (exponent truncated to int) */
static float soft_mask(float x, float y, int power)
{
	float px = powf(x, (float)power), py = powf(y, (float)power);
	return px / (px + py + ZO_EPS);
}
/* hps.h:132-140 */
static float sse_mask(float x, float y) { return x * x / (x * x + y * y + ZO_EPS); }

/* hps.cu:488-580 */
static void apply_median_filter(zo_hpr* h)
{
	int W = h->g.stft_width, nfft = h->g.nfft;
	size_t m = (size_t)W * nfft;
	size_t off = (size_t)(W - h->g.lag) * nfft;
	for (size_t i = 0; i < m; ++i)
		h->s_mag[i] = hypotf(h->stft[2 * i], h->stft[2 * i + 1]); /* hps.h:82-89 */
	filter_pass(h->geom, 0, W, nfft, h->g.l_harm, h->causality, h->copy_bord, h->s_mag, h->hmat);
	filter_pass(h->geom, 0, W, nfft, h->g.l_perc, ZO_FREQUENCY, h->copy_bord, h->s_mag, h->pmat);
	if (h->out_p) {
		for (int k = 0; k < nfft; ++k)
			h->pmask[off + k] = h->soft_mask ? soft_mask(h->pmat[off + k], h->hmat[off + k], (int)h->beta)
			                                 : hard_mask(h->pmat[off + k], h->hmat[off + k], h->beta);
		mask_ifft_ola(h, h->pmask + off, h->pout);
	}
	if (h->out_h) {
		for (int k = 0; k < nfft; ++k)
			h->hmask[off + k] = h->soft_mask ? soft_mask(h->hmat[off + k], h->pmat[off + k], (int)h->beta)
			                                 : hard_mask(h->hmat[off + k], h->pmat[off + k], h->beta - ZO_EPS);
		mask_ifft_ola(h, h->hmask + off, h->hout);
	}
	if (h->out_r && !h->soft_mask) {
		for (size_t i = 0; i < m; ++i)
			h->rmask[i] = 1 - (h->hmask[i] + h->pmask[i]); /* hps.h:35-43 */
		mask_ifft_ola(h, h->rmask + off, h->rout);
	}
}

/* hps.cu:582-652 */
static void apply_sse_filter(zo_hpr* h)
{
	int W = h->g.stft_width, nfft = h->g.nfft;
	size_t m = (size_t)W * nfft;
	size_t off = (size_t)(W - h->g.lag) * nfft;
	for (size_t i = 0; i < m; ++i) {
		h->s_mag[i] = powf(hypotf(h->stft[2 * i], h->stft[2 * i + 1]), 2.0f); /* hps.h:91-98 */
		h->reciprocal[i] = (1.0f / h->s_mag[i]) * 1.0F;                        /* hps.h:45-56 */
	}
	filter_pass(h->geom, 1, W, nfft, h->g.l_harm, h->causality, 1, h->reciprocal, h->hmat);
	filter_pass(h->geom, 1, W, nfft, h->g.l_perc, ZO_FREQUENCY, 1, h->reciprocal, h->pmat);
	for (size_t i = 0; i < m; ++i) {
		h->pmat[i] = (1.0f / h->pmat[i]) * ((float)h->g.l_perc + 1.0F);
		h->hmat[i] = (1.0f / h->hmat[i]) * ((float)h->g.l_harm + 1.0F);
	}
	if (h->out_p) {
		for (int k = 0; k < nfft; ++k)
			h->pmask[off + k] = sse_mask(h->pmat[off + k], h->hmat[off + k]);
		mask_ifft_ola(h, h->pmask + off, h->pout);
	}
	if (h->out_h) {
		for (int k = 0; k < nfft; ++k)
			h->hmask[off + k] = sse_mask(h->hmat[off + k], h->pmat[off + k]);
		mask_ifft_ola(h, h->hmask + off, h->hout);
	}
}

/* hps.cu:429-486 */
void zo_hpr_process_next_hop(zo_hpr* h, const float* in_hop)
{
	int hop = h->g.hop, nwin = h->g.nwin, nfft = h->g.nfft, W = h->g.stft_width;
	float* outs[3] = {h->out_p ? h->pout : NULL, h->out_h ? h->hout : NULL, h->out_r ? h->rout : NULL};
	for (int o = 0; o < 3; ++o) {
		if (!outs[o])
			continue;
		memmove(outs[o], outs[o] + hop, sizeof(float) * (size_t)(nwin - hop));
		memset(outs[o] + hop, 0, sizeof(float) * (size_t)(nwin - hop));
	}
	memmove(h->input, h->input + hop, sizeof(float) * (size_t)(nwin - hop));
	memcpy(h->input + hop, in_hop, sizeof(float) * (size_t)hop);
	for (int i = 0; i < nwin; ++i) {
		h->fft_vec[2 * i] = h->input[i] * h->window[i]; /* hps.h:25-33 */
		h->fft_vec[2 * i + 1] = 0.0f;
	}
	memset(h->fft_vec + 2 * nwin, 0, sizeof(float) * 2 * (size_t)(nfft - nwin));
	hpr_fft(h, -1);
	memmove(h->stft, h->stft + 2 * (size_t)nfft, sizeof(float) * 2 * (size_t)(W - 1) * nfft);
	memcpy(h->stft + 2 * (size_t)(W - 1) * nfft, h->fft_vec, sizeof(float) * 2 * (size_t)nfft);
	if (!h->use_sse)
		apply_median_filter(h);
	else
		apply_sse_filter(h);
}

int zo_hpr_get(const zo_hpr* h, int which, float* out)
{
	size_t m = (size_t)h->g.stft_width * h->g.nfft;
	const float* src = NULL;
	size_t n = m;
	switch (which) {
	case 0: src = h->s_mag; break;
	case 1: src = h->hmat; break;
	case 2: src = h->pmat; break;
	case 3: src = h->hmask; break;
	case 4: src = h->pmask; break;
	case 5: src = h->rmask; break;
	case 6: src = h->hout; n = h->g.nwin; break;
	case 7: src = h->pout; n = h->g.nwin; break;
	case 8: src = h->rout; n = h->g.nwin; break;
	case 9: src = h->reciprocal; break;
	case 10: src = h->input; n = h->g.nwin; break;
	case 11: src = h->window; n = h->g.nwin; break;
	case 12: src = h->stft; n = 2 * m; break;
	default: return -1;
	}
	memcpy(out, src, sizeof(float) * n);
	return (int)n;
}

void zo_hpr_run(zo_hpr* h, const float* audio, int n_hops, float* h_out, float* p_out, float* r_out)
{
	size_t hop = (size_t)h->g.hop;
	for (int i = 0; i < n_hops; ++i) {
		zo_hpr_process_next_hop(h, audio + (size_t)i * hop);
		if (h_out) memcpy(h_out + (size_t)i * hop, h->hout, sizeof(float) * hop);
		if (p_out) memcpy(p_out + (size_t)i * hop, h->pout, sizeof(float) * hop);
		if (r_out) memcpy(r_out + (size_t)i * hop, h->rout, sizeof(float) * hop);
	}
}

/* ------------------------------------------------------------- offline --- */

/* hps.cu:109-126 (chunk count evaluated in float, as written) */
static int chunk_padder(long size, int hop, int lag, long* padded)
{
	int n_chunks = (int)ceilf((float)size / (float)hop);
	long pad = (long)n_chunks * hop - size;
	pad += (long)lag * hop;
	n_chunks += lag;
	*padded = size + pad;
	return n_chunks;
}

int zo_offline_process(int geom, float fs, int hop_h, int hop_p, float beta_h, float beta_p,
                       int nocopybord, int flags, const float* audio, long n,
                       float* h_out, float* p_out, float* r_out)
{
	if (hop_h % hop_p != 0) /* hps.cu:33-36 */
		return -1;
	zo_hpr* ph = zo_hpr_create(geom, fs, hop_h, beta_h, ZO_OUT_HARMONIC | ZO_OUT_PERCUSSIVE | ZO_OUT_RESIDUAL,
	                           ZO_TIME_ANTICAUSAL, !nocopybord);
	zo_hpr* pp = zo_hpr_create(geom, fs, hop_p, beta_p, ZO_OUT_PERCUSSIVE, ZO_TIME_ANTICAUSAL, !nocopybord);
	if (!ph || !pp) {
		zo_hpr_destroy(ph);
		zo_hpr_destroy(pp);
		return -1;
	}
	if (flags & 1) { zo_hpr_use_sse_filter(ph); zo_hpr_use_sse_filter(pp); }
	if (flags & 2) { zo_hpr_use_soft_mask(ph); zo_hpr_use_soft_mask(pp); }

	/* pass 1, hps.cu:135-178 */
	long padded1;
	int n1 = chunk_padder(n, hop_h, ph->g.lag, &padded1);
	float* a1 = zalloc((size_t)padded1);
	memcpy(a1, audio, sizeof(float) * (size_t)n);
	float* inter = zalloc((size_t)padded1);
	float* harm = zalloc((size_t)padded1);
	for (int i = 0; i < n1; ++i) {
		zo_hpr_process_next_hop(ph, a1 + (size_t)i * hop_h);
		for (int j = 0; j < hop_h; ++j) {
			inter[(size_t)i * hop_h + j] = ph->pout[j] + ph->rout[j]; /* hps.h:142-150 */
			harm[(size_t)i * hop_h + j] = ph->hout[j];
		}
	}
	size_t shift1 = (size_t)ph->g.lag * hop_h;
	memmove(inter, inter + shift1, sizeof(float) * ((size_t)padded1 - shift1));
	memmove(harm, harm + shift1, sizeof(float) * ((size_t)padded1 - shift1));

	/* pass 2, hps.cu:180-217.  The signal handed to pass 2 is `intermediate`
	 * (see the note on its tail below). */
	long padded2;
	int n2 = chunk_padder(n, hop_p, pp->g.lag, &padded2);
	float* a2 = zalloc((size_t)padded2);
	/* The reference shrinks `intermediate` to n with vector::resize and then
	 * reads hops up to padded2 from it (hps.cu:176, 189-190): the samples in
	 * [n, padded2) are the not-yet-overwritten pass-1 outputs of the zero
	 * padding (same allocation, padded2 <= padded1), not zeros.  Restated. */
	long keep = padded2 < padded1 ? padded2 : padded1;
	memcpy(a2, inter, sizeof(float) * (size_t)keep);
	float* perc = zalloc((size_t)padded2);
	for (int i = 0; i < n2; ++i) {
		zo_hpr_process_next_hop(pp, a2 + (size_t)i * hop_p);
		memcpy(perc + (size_t)i * hop_p, pp->pout, sizeof(float) * (size_t)hop_p);
	}
	size_t shift2 = (size_t)pp->g.lag * hop_p;
	memmove(perc, perc + shift2, sizeof(float) * ((size_t)padded2 - shift2));

	if (geom == ZO_GEOM_GPU) {
		/* hps.cu:219-220; pass 2 has OUTPUT_PERCUSSIVE only, so its residual_out stays zero */
		memcpy(h_out, harm, sizeof(float) * (size_t)n);
		memcpy(p_out, perc, sizeof(float) * (size_t)n);
		memset(r_out, 0, sizeof(float) * (size_t)n);
	}
	else {
		/* hps.cu:278-279 */
		memcpy(h_out, perc, sizeof(float) * (size_t)n);
		memcpy(p_out, perc, sizeof(float) * (size_t)n);
		memcpy(r_out, perc, sizeof(float) * (size_t)n);
	}
	free(a1); free(inter); free(harm); free(a2); free(perc);
	zo_hpr_destroy(ph);
	zo_hpr_destroy(pp);
	return 0;
}

/* zen/fakert.h:15-34: full chunks strictly before size-k; the trailing
 * `i % k` branch never fires because i is a multiple of k */
long zo_fakert_n_chunks(long size, long hop)
{
	long cnt = 0;
	if (size > hop)
		for (long i = 0; i < size - hop; i += hop)
			++cnt;
	return cnt;
}

/* ------------------------------------------------------ downstream: pitch ---
 * McLeod pitch method as the reference's pitch-tracking demo runs it on the harmonic output of HPRRealtime<GPU>
 * (demos/pitch-tracking/pitch.cpp:40-135, pitch_detection.h:17-93, main.cu:90-107).  Restated as written, quirks
 * included: the power spectrum is only formed for the first N of the 2N bins (pitch.cpp:50-52), the autocorrelation is
 * not normalised into an NSDF, a missing estimate above the cut-off leaves period = 0 (pitch = +inf). */
static void mpm_parabolic(const float* a, int n, int x_, float* ox, float* oy)
{
	/* pitch.cpp:16-38 */
	float x = (float)x_;
	int xa;
	if (x < 1) {
		xa = (a[x_] <= a[x_ + 1]) ? x_ : x_ + 1;
	}
	else if (x > (float)(n - 1)) {
		xa = (a[x_] <= a[x_ - 1]) ? x_ : x_ - 1;
	}
	else {
		float den = a[x_ + 1] + a[x_ - 1] - 2 * a[x_];
		float delta = a[x_ - 1] - a[x_ + 1];
		if (!den) {
			*ox = x;
			*oy = a[x_];
		}
		else {
			*ox = x + delta / (2 * den);
			*oy = a[x_] - delta * delta / (8 * den);
		}
		return;
	}
	*ox = (float)xa;
	*oy = a[xa];
}

/* audio: n samples; nsdf_out (optional): n floats receiving real_autocorrelation's output.  Returns MPM::pitch. */
float zo_mpm_pitch(const float* audio, int n, float sample_rate, float* nsdf_out)
{
	const int n2 = 2 * n;
	float* z = (float*)calloc((size_t)2 * n2, sizeof(float)); /* out_im, interleaved, zero beyond n (pitch.cpp:42-45, clear()) */
	float* r = (float*)malloc(sizeof(float) * (size_t)n);
	for (int i = 0; i < n; ++i)
		z[2 * i] = audio[i];
	zo_fft(n2, z, 0); /* pitch.cpp:47-48 */
	{
		/* pitch.cpp:50-52: out_im[i] *= conj(out_im[i]) * scale for i < N only; complex products as (ac - bd, ad + bc) */
		const float sr = 1.0f / (float)(n * 2), si = 0.0f;
		for (int i = 0; i < n; ++i) {
			const float a = z[2 * i], b = z[2 * i + 1];
			const float cr = a * sr - (-b) * si, ci = a * si + (-b) * sr; /* conj(x) * scale */
			z[2 * i] = a * cr - b * ci;
			z[2 * i + 1] = a * ci + b * cr;
		}
	}
	zo_fft(n2, z, 1); /* pitch.cpp:54-55 */
	for (int i = 0; i < n; ++i)
		r[i] = z[2 * i]; /* pitch.cpp:57-60 */
	if (nsdf_out)
		memcpy(nsdf_out, r, sizeof(float) * (size_t)n);
	/* peak_picking, pitch.cpp:63-99 */
	int* maxpos = (int*)malloc(sizeof(int) * (size_t)(n / 2 + 2));
	int n_max = 0, pos = 0, cur = 0;
	const long size = n;
	while (pos < (size - 1) / 3 && r[pos] > 0)
		pos++;
	while (pos < size - 1 && r[pos] <= 0.0)
		pos++;
	if (pos == 0)
		pos = 1;
	while (pos < size - 1) {
		if (r[pos] > r[pos - 1] && r[pos] >= r[pos + 1] && (cur == 0 || r[pos] > r[cur]))
			cur = pos;
		pos++;
		if (pos < size - 1 && r[pos] <= 0) {
			if (cur > 0) {
				maxpos[n_max++] = cur;
				cur = 0;
			}
			while (pos < size - 1 && r[pos] <= 0.0)
				pos++;
		}
	}
	if (cur > 0)
		maxpos[n_max++] = cur;
	/* MPM::pitch, pitch.cpp:101-135 */
	float highest = -INFINITY; /* (float)-DBL_MAX */
	float* ex = (float*)malloc(sizeof(float) * (size_t)(n_max + 1));
	float* ey = (float*)malloc(sizeof(float) * (size_t)(n_max + 1));
	int n_est = 0;
	for (int m = 0; m < n_max; ++m) {
		const int i = maxpos[m];
		highest = r[i] > highest ? r[i] : highest;
		if (r[i] > 0.5) {
			mpm_parabolic(r, n, i, &ex[n_est], &ey[n_est]);
			highest = ey[n_est] > highest ? ey[n_est] : highest;
			n_est++;
		}
	}
	float result;
	if (n_est == 0) {
		result = -1.0f;
	}
	else {
		const float cutoff = (float)(0.93 * (double)highest);
		float period = 0;
		for (int m = 0; m < n_est; ++m)
			if (ey[m] >= cutoff) {
				period = ex[m];
				break;
			}
		const float est = sample_rate / period;
		result = (est > 80.0) ? est : -1.0f;
	}
	free(z);
	free(r);
	free(maxpos);
	free(ex);
	free(ey);
	return result;
}

/* ---- demos/beat-tracking/OnsetDetection.cpp:60-131: the onset detection function BTrack::processHop (BTrack.cpp:93-98)
 * computes from every 256-sample hop of the percussive output (main.cu:92-118).  Literal restatement, the in-place
 * window + swap of perform_FFT (OnsetDetection.cpp:72-85) included: the frame it shifts next time is the one it has just
 * windowed. ---- */
void zo_onset_window(float* w512)
{
	/* Window.h:31-40 calculate_hanning_window<512>; the reference evaluates cos with gcem at compile time */
	const float PI = 3.14159265359F, N = (float)(512 - 1);
	for (int n = 0; n < 512; ++n)
		w512[n] = 0.5F * (1.0F - cosf(2.0F * PI * ((float)n / N)));
}

void zo_onset_csd(const float* audio, long n_hops, float* out)
{
	enum { FS = 512, HS = 256 };
	float w[FS], frame[FS] = {0}, mag[FS], pmag[FS] = {0}, ph[FS], pph[FS] = {0}, pph2[FS] = {0};
	float z[2 * FS];
	zo_onset_window(w);
	for (long h = 0; h < n_hops; ++h) {
		/* calculate_sample, OnsetDetection.cpp:60-69 */
		memmove(frame, frame + HS, sizeof(float) * (FS - HS));
		memcpy(frame + FS - HS, audio + h * HS, sizeof(float) * HS);
		/* perform_FFT, :72-85 */
		for (int i = 0; i < HS; ++i) {
			const float t = frame[i];
			frame[i] = frame[i + HS];
			frame[i + HS] = t;
			frame[i] *= w[i + HS];
			frame[i + HS] *= w[i];
		}
		for (int i = 0; i < FS; ++i) {
			z[2 * i] = frame[i];
			z[2 * i + 1] = 0.0f; /* imIn stays zero */
		}
		zo_fft(FS, z, 0);
		/* complex_spectral_difference_hwr, :87-131 */
		float sum = 0;
		for (int i = 0; i < FS; ++i) {
			const float re = z[2 * i], im = z[2 * i + 1];
			ph[i] = atan2f(im, re);
			mag[i] = sqrtf(powf(re, 2) + powf(im, 2));
			const float dev = ph[i] - (2 * pph[i]) + pph2[i];
			const float md = mag[i] - pmag[i];
			if (md > 0) {
				const float csd = sqrtf(powf(mag[i], 2) + powf(pmag[i], 2) - 2 * mag[i] * pmag[i] * cosf(dev));
				sum = sum + csd;
			}
			pph2[i] = pph[i];
			pph[i] = ph[i];
			pmag[i] = mag[i];
		}
		out[h] = sum;
	}
}
