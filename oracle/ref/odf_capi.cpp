// TEST INFRASTRUCTURE - flat entry point over the UNMODIFIED onset detection function of the reference's beat-tracking
// demo (demos/beat-tracking/OnsetDetection.cpp, Window.h; the consumer of the percussive output, main.cu:92-118),
// compiled where it lies by oracle/Makefile (target ref_odf) with its IPP FFT served by oracle/ref/ippstub/ipp.h and
// its compile-time window by the vendored gcem.  Pins oracle/hpr_oracle.c:zo_onset_csd (tests/test_onset.py).
#include "OnsetDetection.h"

extern "C" __attribute__((visibility("default"))) void ref_odf_run(const float* audio, long n_hops, float* out)
{
	OnsetDetectionFunction odf;
	for (long h = 0; h < n_hops; ++h)
		out[h] = odf.calculate_sample(audio + h * 256);   // HopSize (OnsetDetection.h:27) == chunk_size (main.cu:38)
}

extern "C" __attribute__((visibility("default"))) void ref_odf_window(float* out512)
{
	const Window<512> w = get_window<512, HanningWindow>();
	for (int i = 0; i < 512; ++i)
		out512[i] = w.data[i];
}
