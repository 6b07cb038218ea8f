// TEST INFRASTRUCTURE - flat entry point over the UNMODIFIED McLeod pitch consumer of the reference
// (demos/pitch-tracking/pitch.cpp, pitch_detection.h), compiled where it lies by oracle/Makefile (target ref_mpm) with
// its IPP FFT served by oracle/ref/ippstub/ipp.h.  Pins oracle/hpr_oracle.c:zo_mpm_pitch (tests/test_pitch.py).
#include "pitch_detection.h"

extern "C" __attribute__((visibility("default"))) float ref_mpm_pitch(const float* audio, long n, float sample_rate)
{
	MPM mpm(n, sample_rate);
	return mpm.pitch(audio);
}
