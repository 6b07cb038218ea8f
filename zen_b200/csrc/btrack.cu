// The beat tracker behind the onset detection function (reference:
// demos/beat-tracking/BTrack.cpp:100-397, called per hop by main.cu:107-121).  It consumes ONE float per 256-sample hop -
// the samples zen_onset_csd (csrc/onset.cu) leaves per stream - and is sequential control logic: a cumulative score with
// a log-Gaussian transition window, a beat prediction half a beat period ahead, and once per beat a tempo update
// (adaptive threshold, autocorrelation through a 512-point FFT, comb filter bank, 41-state Viterbi step).  It stays on
// the host, as in the reference; only the arithmetic is restated, operation for operation, so that the decisions are
// the reference's bit for bit for the same samples (tests/test_btrack.py against the reference compiled unmodified).
//
// Reproduced as they are (bugs as spec):
//   * the FFT behind the "balanced" autocorrelation is planned for 512 points although its buffers hold 1024
//     (BTrack.cpp:40: fft_order = log2(FrameSize)), so the autocorrelation is the circular one of the 512 samples;
//   * the adaptive threshold runs over all 1024 entries of the contiguous buffer, whose upper half is zero
//     (BTrack.cpp:201, 281), and its first eight means start at element 1 (BTrack.cpp:340-343);
//   * the lag of the slowest tempi (129 hops for 80 BPM at 44.1 kHz) indexes the comb filter output past its 128 entries,
//     i.e. into the tempo observation vector that follows it in the object (BTrack.cpp:214-221, BTrack.h:52-53);
//   * the two lookup tables "precomputed with numpy" (BTrackPrecomputed.h) are recomputed here from their formulas:
//     Rayleigh weighting n / 43^2 exp(-n^2 / (2 43^2)), and a Gaussian tempo transition matrix with sigma 5.
#include <cmath>
#include <cstddef>
#include <cstring>
#include <new>

#include "zen_common.cuh"

namespace {

constexpr int BUF = 512;        // OnsetDFBufferSize
constexpr int HOPSZ = 256;      // HopSize
constexpr int NTEMPO = 41;
constexpr float TIGHT = 5.0F, ALPHA = 0.9F, EPS_ODF = 0.0001F;

struct Ring {
	float v[BUF];
	unsigned w;
	float& at(size_t i) { return v[(i + w) & (BUF - 1)]; }
	void push(float x)
	{
		v[w] = x;
		w = (w + 1) & (BUF - 1);
	}
};

// 512-point complex FFT evaluated in double and rounded once (what the IPP stand-in of the test build does; Intel IPP
// itself differs from any of them in the last bits): radix 2, decimation in time
void fft512(const float* re_in, const float* im_in, float* re_out, float* im_out, int sign)
{
	constexpr int N = 512, ORDER = 9;
	static double tw[N];  // cos, sin interleaved for k < N / 2
	static bool have = false;
	if (!have) {
		const double two_pi = 6.283185307179586476925286766559;
		for (int k = 0; k < N / 2; ++k) {
			tw[2 * k] = std::cos(two_pi * (double)k / (double)N);
			tw[2 * k + 1] = std::sin(two_pi * (double)k / (double)N);
		}
		have = true;
	}
	double w[2 * N];
	for (int i = 0; i < N; ++i) {
		unsigned r = 0;
		for (int b = 0; b < ORDER; ++b)
			r |= ((unsigned)(i >> b) & 1u) << (ORDER - 1 - b);
		w[2 * r] = re_in[i];
		w[2 * r + 1] = im_in[i];
	}
	for (int len = 2; len <= N; len <<= 1) {
		const int half = len >> 1, step = N / len;
		for (int base = 0; base < N; base += len)
			for (int j = 0; j < half; ++j) {
				const double c = tw[2 * j * step], s = sign * tw[2 * j * step + 1];
				double* a = w + 2 * (base + j);
				double* b = w + 2 * (base + j + half);
				const double tr = b[0] * c - b[1] * s, ti = b[0] * s + b[1] * c;
				b[0] = a[0] - tr;
				b[1] = a[1] - ti;
				a[0] += tr;
				a[1] += ti;
			}
	}
	for (int i = 0; i < N; ++i) {
		re_out[i] = (float)w[2 * i];
		im_out[i] = (float)w[2 * i + 1];
	}
}

// std::max(a, b) of the reference: `a` survives when the comparison is false, which is what happens to a NaN sample
// (the onset detection function can return one, see csrc/onset.cu)
inline float max_first(float a, float b) { return (a < b) ? b : a; }

float mean_of(const float* x, size_t start, size_t end)  // BTrack.cpp:369-381
{
	float sum = 0;
	const size_t length = end - start;
	for (size_t i = start; i < end; ++i)
		sum = sum + x[i];
	return (length > 0) ? sum / length : 0;
}

// BTrack.cpp:324-367: moving-average threshold (8 back, 7 ahead), subtracted and clipped at zero
void adaptive_threshold(float* x, size_t N, float* thr)
{
	const size_t post = 7, pre = 8;
	const size_t t = N < post ? N : post;
	for (size_t i = 0; i <= t; ++i) {
		const size_t k = (i + pre) < N ? (i + pre) : N;
		thr[i] = mean_of(x, 1, k);
	}
	for (size_t i = t + 1; i < N - post; ++i)
		thr[i] = mean_of(x, i - pre, i + post);
	for (size_t i = N - post; i < N; ++i) {
		const size_t k = (i - post) > 1 ? (i - post) : 1;
		thr[i] = mean_of(x, k, N);
	}
	for (size_t i = 0; i < N; ++i) {
		x[i] = x[i] - thr[i];
		if (x[i] < 0) x[i] = 0;
	}
}

}  // namespace

struct zen_btrack {
	int sample_rate;
	float lag_factor;      // tempoToLagFactor
	float period;          // beatPeriod, in hops
	int m0, countdown;     // hops until the next prediction / until the predicted beat
	float tempo, latest;
	Ring odf, score;
	float w1[BUF];         // transition window of the last cumulative-score update (re-used by the prediction)
	float prev_delta[NTEMPO];
	float rayleigh[128];
	float transition[NTEMPO][NTEMPO];
	// work areas of the tempo update
	float contiguous[2 * BUF], thr[2 * BUF], acf[BUF];
	// comb filter bank output (128) FOLLOWED BY the tempo observation vector (41), as the reference's object lays them
	// out: its lag index for the slowest tempi exceeds 128 (80 BPM at 44.1 kHz: 129) and reads the observation vector
	// through the end of the comb array (BTrack.cpp:214-221) - previous call's values, or the ones just written
	float comb[128 + NTEMPO];
	float fre[2 * BUF], fim[2 * BUF], zero_im[2 * BUF];

	void update_score(float s);
	void predict();
	void update_tempo();
	bool step(float sample);
};

void zen_btrack::update_score(float s)  // BTrack.cpp:120-135, 383-397
{
	const size_t start = (size_t)(BUF - roundf(2.0F * period));
	const size_t end = (size_t)(BUF - roundf(period / 2.0F));
	{
		float v = -2.0F * period;
		const size_t n = end - start + 1;
		for (size_t i = 0; i < n; ++i) {
			w1[i] = expf((-1 * powf(TIGHT * logf(-v / period), 2.0F)) / 2.0F);
			v += 1.0F;
		}
	}
	float mx = 0.0F;
	for (size_t i = start; i <= end; ++i)
		mx = max_first(score.at(i) * w1[i - start], mx);
	latest = ((1.0F - ALPHA) * s) + (ALPHA * mx);
	score.push(latest);
}

void zen_btrack::predict()  // BTrack.cpp:137-191
{
	const size_t window = (size_t)period;
	// (the reference's arrays hold 128 / 512 + 128 entries and a beat period above 128 hops - 80 BPM at 44.1 kHz gives 129 -
	// runs over their end into the next member; sized here for the longest period the tempo grid can produce)
	float future[BUF + 512];
	float w2[512];
	for (size_t i = 0; i < BUF; ++i)
		future[i] = score.at(i);
	float v = 1.0F;
	for (size_t i = 0; i < window; ++i) {
		w2[i] = expf((-1.0F * powf((v - (period / 2.0F)), 2.0F)) / (2.0F * powf((period / 2.0F), 2.0F)));
		v += 1.0F;
	}
	for (size_t i = BUF; i < BUF + window; ++i) {
		const size_t start = (size_t)(i - roundf(2.0F * period));
		const size_t end = (size_t)(i - roundf(period / 2.0F));
		float mx = 0;
		int n = 0;
		for (size_t k = start; k <= end; ++k, ++n) {
			const float c = future[k] * w1[n];
			if (c > mx) mx = c;
		}
		future[i] = mx;
	}
	float mx = 0;
	int n = 0;
	for (size_t i = BUF; i < BUF + window; ++i, ++n) {
		const float c = future[i] * w2[n];
		if (c > mx) {
			mx = c;
			countdown = n;
		}
	}
	m0 = (int)(countdown + roundf(period / 2.0F));
}

void zen_btrack::update_tempo()  // BTrack.cpp:193-303
{
	for (size_t i = 0; i < BUF; ++i)
		contiguous[i] = odf.at(i);
	adaptive_threshold(contiguous, 2 * BUF, thr);
	// "balanced" autocorrelation: 512-point transform of the lower half, power spectrum, inverse, |.| / (512 - lag)
	std::memset(contiguous + BUF, 0, sizeof(float) * BUF);
	fft512(contiguous, zero_im, fre, fim, -1);
	for (int i = 0; i < 2 * BUF; ++i) {
		fre[i] = fre[i] * fre[i] + fim[i] * fim[i];
		fim[i] = 0.0F;
	}
	fft512(fre, fim, fre, fim, +1);
	for (size_t i = 0; i < BUF; ++i)
		acf[i] = sqrtf(fre[i] * fre[i] + fim[i] * fim[i]) / (float)(BUF - i);
	// comb filter bank over beat periods 2 .. 127, up to four comb teeth of growing width
	for (int i = 0; i < 128; ++i)
		comb[i] = 0.0F;
	for (int i = 2; i <= 127; ++i)
		for (int a = 1; a <= 4; ++a)
			for (int b = 1 - a; b <= a - 1; ++b)
				comb[i - 1] = comb[i - 1] + (acf[(a * i + b) - 1] * rayleigh[i - 1]) / (2 * a - 1);
	adaptive_threshold(comb, 128, thr);
	float* const obs = comb + 128;
	float delta[NTEMPO];
	for (size_t i = 0; i < NTEMPO; ++i) {
		const size_t t1 = (size_t)roundf(lag_factor / (((2.0F * i) + 80.0F)));
		const size_t t2 = t1 / 2;
		obs[i] = comb[t1 - 1] + comb[t2 - 1];
	}
	for (size_t j = 0; j < NTEMPO; ++j) {
		float mx = -1.0F;
		for (size_t i = 0; i < NTEMPO; ++i)
			mx = max_first(prev_delta[i] * transition[i][j], mx);
		delta[j] = mx * obs[j];
	}
	{
		float sum = 0.0F;
		for (size_t i = 0; i < NTEMPO; ++i)
			if (delta[i] > 0) sum += delta[i];
		if (sum > 0)
			for (size_t i = 0; i < NTEMPO; ++i)
				delta[i] /= sum;
	}
	float best = -1, best_i = -1;
	for (size_t j = 0; j < NTEMPO; ++j) {
		if (delta[j] > best) {
			best = delta[j];
			best_i = j;
		}
		prev_delta[j] = delta[j];
	}
	period = roundf((60.0F * ((float)sample_rate)) / (((2.0F * best_i) + 80.0F) * ((float)HOPSZ)));
	if (period > 0) tempo = 60.0F / ((((float)HOPSZ) / ((float)sample_rate)) * period);
}

bool zen_btrack::step(float sample)  // BTrack.cpp:100-118
{
	sample = fabsf(sample) + EPS_ODF;
	m0--;
	countdown--;
	bool beat = false;
	odf.push(sample);
	update_score(sample);
	if (m0 == 0) predict();
	if (countdown == 0) {
		beat = true;
		update_tempo();
	}
	return beat;
}

extern "C" {

int zen_btrack_create(zen_btrack** out, int sample_rate)
{
	if (!out || sample_rate < 1)
		return ZEN_ERR_ARG;
	zen_btrack* b = new (std::nothrow) zen_btrack();
	if (!b)
		return ZEN_ERR_ARG;
	std::memset(b, 0, sizeof(*b));
	b->sample_rate = sample_rate;
	b->lag_factor = 60.0F * ((float)sample_rate) / (float)HOPSZ;
	b->period = roundf(60.0F / ((((float)HOPSZ) / (float)sample_rate) * 120.0F));
	b->m0 = 10;
	b->countdown = -1;
	b->tempo = 120.0F;
	for (int i = 0; i < NTEMPO; ++i)
		b->prev_delta[i] = 1.0F;
	for (size_t i = 0; i < BUF; ++i)
		if ((i % ((size_t)round(b->period))) == 0) b->odf.v[i] = 1.0F;
	for (int n = 0; n < 128; ++n)
		b->rayleigh[n] = (float)(((double)n / (43.0 * 43.0)) * std::exp(-((double)n * (double)n) / (2.0 * 43.0 * 43.0)));
	for (int i = 0; i < NTEMPO; ++i)
		for (int j = 0; j < NTEMPO; ++j) {
			const double d = (double)(j + 1) - (double)(i + 1), sig = 5.0;
			b->transition[i][j] = (float)((1.0 / (sig * std::sqrt(2.0 * 3.141592653589793))) * std::exp(-(d * d) / (2.0 * sig * sig)));
		}
	*out = b;
	return ZEN_OK;
}

void zen_btrack_destroy(zen_btrack* b) { delete b; }

int zen_btrack_process(zen_btrack* b, const float* h_odf, long n, unsigned char* h_beat, float* h_tempo, float* h_cumscore)
{
	if (!b || (n > 0 && !h_odf) || n < 0)
		return ZEN_ERR_ARG;
	for (long i = 0; i < n; ++i) {
		const bool beat = b->step(h_odf[i]);
		if (h_beat) h_beat[i] = beat ? 1 : 0;
		if (h_tempo) h_tempo[i] = b->tempo;
		if (h_cumscore) h_cumscore[i] = b->latest;
	}
	return ZEN_OK;
}

// the two lookup tables (128 and 41 x 41 floats), for the tests that compare them with the reference's precomputed ones
int zen_btrack_tables(const zen_btrack* b, float* h_rayleigh128, float* h_transition41x41)
{
	if (!b || !h_rayleigh128 || !h_transition41x41)
		return ZEN_ERR_ARG;
	std::memcpy(h_rayleigh128, b->rayleigh, sizeof(b->rayleigh));
	std::memcpy(h_transition41x41, b->transition, sizeof(b->transition));
	return ZEN_OK;
}

}  // extern "C"
