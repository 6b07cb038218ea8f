// TEST INFRASTRUCTURE - flat entry points over the UNMODIFIED beat tracker of the reference's beat-tracking demo
// (demos/beat-tracking/BTrack.cpp, OnsetDetection.cpp, BTrackPrecomputed.h), compiled where it lies by oracle/Makefile
// (target ref_btrack) with its IPP FFTs served by oracle/ref/ippstub/ipp.h.  Pins the host-side beat tracker of the
// product (zen_btrack_*, tests/test_btrack.py).
#include "BTrack.h"
#include "BTrackPrecomputed.h"

// one BTrack::processHop per 256-sample hop (main.cu:107-121); per hop: the onset detection function sample the tracker
// consumed (lastOnset), the cumulative score, whether a beat is due, and the tempo estimate
extern "C" __attribute__((visibility("default"))) void ref_btrack_run(int sample_rate, const float* audio, long n_hops, float* odf,
                                                                      float* cumscore, unsigned char* beat, float* tempo)
{
	BTrack bt(sample_rate);
	for (long h = 0; h < n_hops; ++h) {
		bt.processHop(audio + h * 256);
		odf[h] = bt.lastOnset;
		cumscore[h] = bt.latestCumulativeScoreValue;
		beat[h] = bt.beatDueInFrame ? 1 : 0;
		tempo[h] = bt.estimatedTempo;
	}
}

extern "C" __attribute__((visibility("default"))) void ref_btrack_tables(float* rayleigh128, float* transition41x41)
{
	for (int i = 0; i < 128; ++i)
		rayleigh128[i] = precomputed::RayleighWeightingVector128[i];
	for (int i = 0; i < 41; ++i)
		for (int j = 0; j < 41; ++j)
			transition41x41[i * 41 + j] = precomputed::TempoTransitionMatrix[i][j];
}
