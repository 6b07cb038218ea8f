// TEST INFRASTRUCTURE - flat entry points over the UNMODIFIED libnyquist conversion code the zen command line uses
// for its on-disk format (vendor/libnyquist/src/Common.cpp:300-337, include/libnyquist/Common.h:288-302, 669-675),
// compiled where it lies under /root/reference/vendor by oracle/Makefile (target ref_nyq) into
// oracle/_ref/libnyq_ref.so.  Used to pin oracle/np_model.py and csrc/pcm.cu (tests/test_pcm.py).
#include <algorithm>
#include <cstdint>
#include <vector>

#include "Common.h"

#define NYQ_API __attribute__((visibility("default")))

extern "C" {

// nqr::ConvertToFloat32(float*, const int16_t*, N, PCM_16): what WavDecoder does with a PCM16 payload
NYQ_API void nyq_pcm16_to_float(const int16_t* src, float* dst, long n) { nqr::ConvertToFloat32(dst, src, (size_t)n, nqr::PCM_16); }

// nqr::StereoToMono as zen/offline.h:104-117 / zen/fakert.h:117-130 call it (N = interleaved sample count)
NYQ_API void nyq_stereo_to_mono(const float* src, float* dst, long n_interleaved) { nqr::StereoToMono(src, dst, (size_t)n_interleaved); }

// nqr::ConvertFromFloat32(..., PCM_16, DITHER_NONE): what encode_wav_to_disk does with EncoderParams{1, PCM_16, DITHER_NONE}
NYQ_API void nyq_float_to_pcm16(const float* src, int16_t* dst, long n)
{
	nqr::ConvertFromFloat32(reinterpret_cast<uint8_t*>(dst), src, (size_t)n, nqr::PCM_16, nqr::DITHER_NONE);
}

// the command line's peak normalisation, restated from zen/offline.h:180-192 (zen/fakert.h:259-268 is the same code):
// minmax_element, real_max = max(-1 * min, max), every sample divided by it - then the libnyquist encode above
NYQ_API float zen_cli_normalize_encode(const float* src, int16_t* dst, long n)
{
	std::vector<float> x(src, src + n);
	auto limits = std::minmax_element(x.begin(), x.end());
	float real_max = std::max(-1 * (*limits.first), *limits.second);
	for (long j = 0; j < n; ++j)
		x[j] /= real_max;
	nyq_float_to_pcm16(x.data(), dst, n);
	return real_max;
}

}  // extern "C"
