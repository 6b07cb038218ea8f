/* TEST INFRASTRUCTURE — stand-in for the 13 Intel IPP entry points that the
 * reference's CPU backend calls (/root/reference/libzen/mfilt.h:320-341,
 * box.h:266-287, fftw.h:69-114).  Intel IPP is closed source and absent from
 * this image, so the reference's CPU path can only be compiled against this
 * header.  Semantics restated from IPP's published documentation:
 *   - ippiFilterMedianBorder_32f_C1R / ippiFilterBoxBorder_32f_C1R with
 *     ippBorderRepl: window centred on the pixel, out-of-image taps replicate
 *     the nearest edge pixel, every ROI pixel is written;
 *   - ippsFFT{Fwd,Inv}_CToC_32fc_I with IPP_FFT_NODIV_BY_ANY: unnormalised in
 *     both directions.
 * Nothing under zen_b200/ includes this file. */
#ifndef ZEN_ORACLE_IPP_STANDIN_H
#define ZEN_ORACLE_IPP_STANDIN_H

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

typedef unsigned char Ipp8u;
typedef float Ipp32f;
typedef struct {
	Ipp32f re;
	Ipp32f im;
} Ipp32fc;
typedef struct {
	int width;
	int height;
} IppiSize;
typedef int IppStatus;
typedef int IppDataType;
typedef int IppiBorderType;
typedef int IppHintAlgorithm;

enum { ippStsNoErr = 0 };
enum { ipp32f = 13 };
enum { ippBorderRepl = 1 };
enum { ippAlgHintNone = 0 };
enum { IPP_FFT_NODIV_BY_ANY = 8 };

struct IppsFFTSpec_C_32fc {
	int order;
	int n;
	double* tw; /* cos/sin pairs, n/2 entries */
};

static inline const char* ippGetStatusString(IppStatus) { return "ipp stand-in"; }
static inline Ipp8u* ippsMalloc_8u(int len) { return (Ipp8u*)std::malloc(len > 0 ? len : 1); }
static inline void ippsFree(void* p) { std::free(p); }
static inline void* ippMalloc(int len) { return std::malloc(len > 0 ? len : 1); }
static inline void ippFree(void* p) { std::free(p); }

static inline IppStatus ippiFilterMedianBorderGetBufferSize(
    IppiSize, IppiSize mask, IppDataType, int, int* sz)
{
	*sz = (int)sizeof(float) * (mask.width * mask.height + 8);
	return ippStsNoErr;
}

static inline int ipp_standin_clamp(int v, int lo, int hi)
{
	return v < lo ? lo : (v > hi ? hi : v);
}

/* centred window, replicate border, nth_element selection */
static inline IppStatus ippiFilterMedianBorder_32f_C1R(const Ipp32f* src,
                                                       int srcStep,
                                                       Ipp32f* dst,
                                                       int dstStep,
                                                       IppiSize roi,
                                                       IppiSize mask,
                                                       IppiBorderType,
                                                       Ipp32f,
                                                       Ipp8u*)
{
	const int W = roi.width, H = roi.height;
	const int mw = mask.width, mh = mask.height;
	const int ax = mw / 2, ay = mh / 2;
	const int ss = srcStep / (int)sizeof(float), ds = dstStep / (int)sizeof(float);
	const int n = mw * mh, k = n / 2;
	std::vector<float> win(n);
	if (mh == 1) {
		/* row filter: keep the window sorted while it slides */
		std::vector<float> sorted(n);
		for (int y = 0; y < H; ++y) {
			const float* row = src + (size_t)y * ss;
			for (int t = 0; t < n; ++t)
				sorted[t] = row[ipp_standin_clamp(t - ax, 0, W - 1)];
			std::sort(sorted.begin(), sorted.end());
			dst[(size_t)y * ds] = sorted[k];
			for (int x = 1; x < W; ++x) {
				float out = row[ipp_standin_clamp(x - 1 - ax, 0, W - 1)];
				float in = row[ipp_standin_clamp(x + ax, 0, W - 1)];
				if (out != in) {
					float* p = std::lower_bound(sorted.data(), sorted.data() + n, out);
					float* q = std::lower_bound(sorted.data(), sorted.data() + n, in);
					if (q > p) {
						std::memmove(p, p + 1, (size_t)(q - p - 1) * sizeof(float));
						*(q - 1) = in;
					}
					else {
						std::memmove(q + 1, q, (size_t)(p - q) * sizeof(float));
						*q = in;
					}
				}
				dst[(size_t)y * ds + x] = sorted[k];
			}
		}
		return ippStsNoErr;
	}
	for (int y = 0; y < H; ++y) {
		for (int x = 0; x < W; ++x) {
			int t = 0;
			for (int j = 0; j < mh; ++j) {
				int yy = ipp_standin_clamp(y + j - ay, 0, H - 1);
				for (int i = 0; i < mw; ++i) {
					int xx = ipp_standin_clamp(x + i - ax, 0, W - 1);
					win[t++] = src[(size_t)yy * ss + xx];
				}
			}
			std::nth_element(win.begin(), win.begin() + k, win.end());
			dst[(size_t)y * ds + x] = win[k];
		}
	}
	return ippStsNoErr;
}

static inline IppStatus ippiFilterBoxBorderGetBufferSize(
    IppiSize, IppiSize mask, IppDataType, int, int* sz)
{
	*sz = (int)sizeof(float) * (mask.width * mask.height + 8);
	return ippStsNoErr;
}

/* centred window, replicate border, arithmetic mean accumulated in float in
 * raster order of the taps */
static inline IppStatus ippiFilterBoxBorder_32f_C1R(const Ipp32f* src,
                                                    int srcStep,
                                                    Ipp32f* dst,
                                                    int dstStep,
                                                    IppiSize roi,
                                                    IppiSize mask,
                                                    IppiBorderType,
                                                    const Ipp32f*,
                                                    Ipp8u*)
{
	const int W = roi.width, H = roi.height;
	const int mw = mask.width, mh = mask.height;
	const int ax = mw / 2, ay = mh / 2;
	const int ss = srcStep / (int)sizeof(float), ds = dstStep / (int)sizeof(float);
	const float inv = 1.0f / (float)(mw * mh);
	for (int y = 0; y < H; ++y) {
		for (int x = 0; x < W; ++x) {
			float acc = 0.0f;
			for (int j = 0; j < mh; ++j) {
				int yy = ipp_standin_clamp(y + j - ay, 0, H - 1);
				for (int i = 0; i < mw; ++i) {
					int xx = ipp_standin_clamp(x + i - ax, 0, W - 1);
					acc += src[(size_t)yy * ss + xx];
				}
			}
			dst[(size_t)y * ds + x] = acc * inv;
		}
	}
	return ippStsNoErr;
}

static inline IppStatus ippsFFTGetSize_C_32fc(
    int order, int, IppHintAlgorithm, int* specSize, int* initSize, int* bufSize)
{
	int n = 1 << order;
	*specSize = (int)sizeof(IppsFFTSpec_C_32fc) + (int)sizeof(double) * n + 64;
	*initSize = 0;
	*bufSize = (int)sizeof(double) * 2 * n + 64;
	return ippStsNoErr;
}

static inline IppStatus ippsFFTInit_C_32fc(IppsFFTSpec_C_32fc** pp,
                                           int order,
                                           int,
                                           IppHintAlgorithm,
                                           Ipp8u* specMem,
                                           Ipp8u*)
{
	IppsFFTSpec_C_32fc* s = (IppsFFTSpec_C_32fc*)specMem;
	s->order = order;
	s->n = 1 << order;
	size_t off = (sizeof(IppsFFTSpec_C_32fc) + 15) & ~(size_t)15;
	s->tw = (double*)(specMem + off);
	const double two_pi = 6.283185307179586476925286766559;
	for (int i = 0; i < s->n / 2; ++i) {
		s->tw[2 * i] = std::cos(two_pi * (double)i / (double)s->n);
		s->tw[2 * i + 1] = std::sin(two_pi * (double)i / (double)s->n);
	}
	*pp = s;
	return ippStsNoErr;
}

/* radix-2 decimation-in-time, evaluated in double, rounded to float once */
static inline void ipp_standin_fft(Ipp32fc* x, const IppsFFTSpec_C_32fc* s, Ipp8u* buf, int sign)
{
	const int n = s->n, order = s->order;
	double* w = (double*)(((size_t)buf + 15) & ~(size_t)15);
	for (int i = 0; i < n; ++i) {
		unsigned r = 0;
		for (int b = 0; b < order; ++b)
			r |= ((unsigned)(i >> b) & 1u) << (order - 1 - b);
		w[2 * r] = x[i].re;
		w[2 * r + 1] = x[i].im;
	}
	for (int len = 2; len <= n; len <<= 1) {
		int half = len >> 1, step = n / len;
		for (int base = 0; base < n; base += len) {
			for (int j = 0; j < half; ++j) {
				double c = s->tw[2 * j * step];
				double sn = sign * s->tw[2 * j * step + 1];
				double* a = w + 2 * (base + j);
				double* b = w + 2 * (base + j + half);
				double tr = b[0] * c - b[1] * sn;
				double ti = b[0] * sn + b[1] * c;
				b[0] = a[0] - tr;
				b[1] = a[1] - ti;
				a[0] += tr;
				a[1] += ti;
			}
		}
	}
	for (int i = 0; i < n; ++i) {
		x[i].re = (float)w[2 * i];
		x[i].im = (float)w[2 * i + 1];
	}
}

static inline IppStatus ippsFFTFwd_CToC_32fc_I(Ipp32fc* x, const IppsFFTSpec_C_32fc* s, Ipp8u* buf)
{
	ipp_standin_fft(x, s, buf, -1);
	return ippStsNoErr;
}

static inline IppStatus ippsFFTInv_CToC_32fc_I(Ipp32fc* x, const IppsFFTSpec_C_32fc* s, Ipp8u* buf)
{
	ipp_standin_fft(x, s, buf, +1);
	return ippStsNoErr;
}

/* split-array variant (demos/beat-tracking/OnsetDetection.cpp:16-47, 72-85): real and imaginary parts in separate
 * arrays, out of place */
typedef IppsFFTSpec_C_32fc IppsFFTSpec_C_32f;
static inline IppStatus ippsFFTGetSize_C_32f(int order, int flag, IppHintAlgorithm hint, int* specSize, int* initSize, int* bufSize)
{
	IppStatus st = ippsFFTGetSize_C_32fc(order, flag, hint, specSize, initSize, bufSize);
	*bufSize += (int)sizeof(Ipp32fc) * (1 << order);
	return st;
}
static inline IppStatus ippsFFTInit_C_32f(IppsFFTSpec_C_32f** pp, int order, int flag, IppHintAlgorithm hint, Ipp8u* specMem, Ipp8u* initMem)
{
	return ippsFFTInit_C_32fc(pp, order, flag, hint, specMem, initMem);
}
static inline IppStatus ippsFFTFwd_CToC_32f(const Ipp32f* srcRe, const Ipp32f* srcIm, Ipp32f* dstRe, Ipp32f* dstIm,
                                            const IppsFFTSpec_C_32f* s, Ipp8u* buf)
{
	const int n = s->n;
	Ipp32fc* x = (Ipp32fc*)(((size_t)buf + 15) & ~(size_t)15);
	Ipp8u* rest = (Ipp8u*)(x + n);
	for (int i = 0; i < n; ++i) {
		x[i].re = srcRe[i];
		x[i].im = srcIm[i];
	}
	ipp_standin_fft(x, s, rest, -1);
	for (int i = 0; i < n; ++i) {
		dstRe[i] = x[i].re;
		dstIm[i] = x[i].im;
	}
	return ippStsNoErr;
}

/* in-place inverse of the split-array variant (demos/beat-tracking/BTrack.cpp:289-290) */
static inline IppStatus ippsFFTInv_CToC_32f_I(Ipp32f* re, Ipp32f* im, const IppsFFTSpec_C_32f* s, Ipp8u* buf)
{
	const int n = s->n;
	Ipp32fc* x = (Ipp32fc*)(((size_t)buf + 15) & ~(size_t)15);
	Ipp8u* rest = (Ipp8u*)(x + n);
	for (int i = 0; i < n; ++i) {
		x[i].re = re[i];
		x[i].im = im[i];
	}
	ipp_standin_fft(x, s, rest, +1);
	for (int i = 0; i < n; ++i) {
		re[i] = x[i].re;
		im[i] = x[i].im;
	}
	return ippStsNoErr;
}

#endif /* ZEN_ORACLE_IPP_STANDIN_H */
