// Compiles against zen_b200/include exactly the way a libzen caller would
// (zen/fakert.h, zen/offline.h, libzen/hps.test.cu call patterns) and checks the
// C++ surface on a GPU.  Built and run by tests/test_cpp_dropin.py.
#include <cstdio>
#include <cmath>
#include <random>
#include <vector>

#include <hps.h>
#include <libzen/hps.h>
#include <libzen/io.h>
#include <libzen/zen.h>
#include <mfilt.h>

using namespace zen::internal::hps;
using namespace zen::internal::hps::mfilt;
using namespace zen;

#define CHECK(cond)                                                              \
	do {                                                                         \
		if (!(cond)) {                                                           \
			std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond);        \
			return 1;                                                            \
		}                                                                        \
	} while (0)

int main()
{
	// --- mfilt.test.cu style: cross pattern, copy-border, exact equality, ZgException
	{
		int x = 9, y = 9;
		thrust::device_vector<float> testdata(x * y, 0.0F), result(x * y, 0.0F);
		for (int i = 0; i < x; ++i)
			for (int j = 0; j < y; ++j) {
				if (i == x / 2) testdata[i * y + j] = 5;
				if (j == y / 2) testdata[i * y + j] = 8;
			}
		MedianFilterGPU causal(x, y, 3, MedianFilterDirection::TimeCausal, true);
		causal.filter(testdata, result);
		for (int i = 0; i < x; ++i)
			for (int j = 0; j < y; ++j)
				CHECK(result[i * y + j] == (j == y / 2 ? 8 : 0));
		bool threw = false;
		try {
			MedianFilterGPU bad(x, y, 10, MedianFilterDirection::Frequency);
		}
		catch (const ZgException&) {
			threw = true;
		}
		CHECK(threw);
	}
	// --- hps.test.cu style: HPR<GPU> members, reset reproducibility, percussive-only
	{
		std::size_t hop = 256;
		std::uniform_real_distribution<float> dist(-1.0F, 1.0F);
		std::default_random_engine gen;
		std::vector<float> data(100 * hop);
		for (auto& v : data) v = dist(gen);
		zen::io::IOGPU io(hop);
		HPR<Backend::GPU> hpr(48000.0F, hop, 2.0, zen::hps::OUTPUT_PERCUSSIVE, MedianFilterDirection::TimeCausal, true);
		CHECK(hpr.nwin == 2 * hop && hpr.nfft == 4 * hop && hpr.l_harm == 12 && hpr.stft_width == 24 && hpr.lag == 1);
		std::vector<float> first(hop);
		for (int i = 0; i < 30; ++i) {
			std::copy(data.begin() + i * hop, data.begin() + (i + 1) * hop, io.host_in);
			hpr.process_next_hop(io.device_in);
			if (i == 0)
				for (std::size_t j = 0; j < hop; ++j) first[j] = hpr.percussive_out[j];
		}
		bool any = false;
		for (std::size_t j = 0; j < hop; ++j) {
			CHECK(hpr.harmonic_out[j] == 0.0F && hpr.residual_out[j] == 0.0F);
			any |= hpr.percussive_out[j] != 0.0F;
		}
		CHECK(any);
		hpr.reset_buffers();
		std::copy(data.begin(), data.begin() + hop, io.host_in);
		hpr.process_next_hop(io.device_in);
		for (std::size_t j = 0; j < hop; ++j)
			CHECK(hpr.percussive_out[j] == first[j]);
		hpr.refresh_matrices();
		float m = hpr.s_mag[(hpr.stft_width - 1) * hpr.nfft + 3];
		CHECK(m > 0.0F && std::isfinite(m));
	}
	// --- zen/fakert.h style loop
	{
		std::size_t hop = 1024;
		auto hpss = zen::hps::HPRRealtime<Backend::GPU>(44100.0F, hop, 2.5, zen::hps::OUTPUT_PERCUSSIVE, false);
		auto io = zen::io::IOGPU(hop);
		hpss.warmup(io);
		std::vector<float> out(8 * hop);
		for (int i = 0; i < 8; ++i) {
			for (std::size_t j = 0; j < hop; ++j)
				io.host_in[j] = std::sin(0.05f * (float)(i * hop + j)) + ((j % 512) == 0 ? 0.9f : 0.0f);
			hpss.process_next_hop(io.device_in);
			hpss.copy_percussive(io.device_out);
			std::copy(io.host_out, io.host_out + hop, out.begin() + i * hop);
		}
		float peak = 0;
		for (float v : out) peak = std::fmax(peak, std::fabs(v));
		CHECK(peak > 0.0F && std::isfinite(peak));
	}
	// --- zen/offline.h style
	{
		std::vector<float> audio(20 * 4096 + 11);
		for (std::size_t i = 0; i < audio.size(); ++i)
			audio[i] = 0.5f * std::sin(0.03f * (float)i) + ((i % 9000) < 40 ? 0.8f : 0.0f);
		auto hpss = zen::hps::HPRIOffline<Backend::GPU>(48000.0F, 4096, 256, 2.0, 2.0);
		auto all = hpss.process(audio);
		CHECK(all[0].size() == audio.size() && all[1].size() == audio.size() && all[2].size() == audio.size());
		bool differs = false, zero_res = true;
		for (std::size_t i = 0; i < audio.size(); ++i) {
			differs |= all[1][i] != audio[i];
			zero_res &= all[2][i] == 0.0F;
		}
		CHECK(differs && zero_res);
		bool threw = false;
		try {
			zen::hps::HPRIOffline<Backend::GPU> bad(48000.0F, 4096, 300, 2.0, 2.0);
		}
		catch (const ZgException&) {
			threw = true;
		}
		CHECK(threw);
	}
	std::printf("dropin_test OK\n");
	return 0;
}
