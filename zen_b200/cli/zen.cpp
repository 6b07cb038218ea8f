// `zen` command line for the B200 HPR path: same sub-commands, flags, defaults,
// output files and timing lines as the reference CLI (zen/main.cu:20-63,
// zen/offline.h, zen/fakert.h), on top of the C++ drop-in headers.
//
//   zen offline -i <wav> [--hps [hop-h [beta-h [hop-p [beta-p]]]]] [-o <prefix>] [--cpu] [--sse]
//               [--only-percussive] [--soft-mask] [--nocopybord]
//   zen fakert  -i <wav> [--hps [hop [beta]]] [-o <wav>] [--cpu] [--sse] [--soft-mask] [--nocopybord]
//   zen help | -h | --help          zen version | -v | --version
//
// The wav layer replaces libnyquist for the one format the reference's sample uses
// (RIFF/WAVE PCM16, also 32-bit float in): int16 -> float is s/32767, float -> int16 is
// lroundf(s*32767) (vendor/libnyquist Common.h:296-302, Common.cpp:332-337), stereo is folded
// to mono as (l+r)/2 (Common.h:669-675).  --cpu is refused: this build has no CPU backend.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include <libzen/hps.h>
#include <libzen/io.h>

namespace {

struct Wav {
	int sample_rate = 0;
	int channels = 0;
	std::vector<float> samples;  // interleaved
};

bool read_wav(const std::string& path, Wav& w, std::string& err)
{
	std::ifstream f(path, std::ios::binary);
	if (!f) {
		err = "cannot open " + path;
		return false;
	}
	std::vector<unsigned char> b((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
	auto u16 = [&](size_t o) { return (uint32_t)b[o] | ((uint32_t)b[o + 1] << 8); };
	auto u32 = [&](size_t o) { return u16(o) | (u16(o + 2) << 16); };
	if (b.size() < 44 || std::memcmp(b.data(), "RIFF", 4) != 0 || std::memcmp(b.data() + 8, "WAVE", 4) != 0) {
		err = "not a RIFF/WAVE file (a Git-LFS pointer?)";
		return false;
	}
	int fmt = 0, bits = 0;
	size_t pos = 12;
	bool have_fmt = false;
	while (pos + 8 <= b.size()) {
		uint32_t sz = u32(pos + 4);
		const unsigned char* id = b.data() + pos;
		size_t body = pos + 8;
		if (std::memcmp(id, "fmt ", 4) == 0 && body + 16 <= b.size()) {
			fmt = (int)u16(body);
			w.channels = (int)u16(body + 2);
			w.sample_rate = (int)u32(body + 4);
			bits = (int)u16(body + 14);
			if (fmt == 0xFFFE && sz >= 26) fmt = (int)u16(body + 24);  // WAVE_FORMAT_EXTENSIBLE sub-format
			have_fmt = true;
		}
		else if (std::memcmp(id, "data", 4) == 0) {
			if (!have_fmt) break;
			size_t n = std::min<size_t>(sz, b.size() - body);
			if (fmt == 1 && bits == 16) {
				w.samples.resize(n / 2);
				for (size_t i = 0; i < w.samples.size(); ++i)
					w.samples[i] = (float)(int16_t)u16(body + 2 * i) / 32767.f;
			}
			else if (fmt == 3 && bits == 32) {
				w.samples.resize(n / 4);
				std::memcpy(w.samples.data(), b.data() + body, w.samples.size() * 4);
			}
			else {
				err = "unsupported wav encoding (PCM16 or float32 only)";
				return false;
			}
			return w.channels >= 1;
		}
		pos = body + sz + (sz & 1);
	}
	err = "no data chunk";
	return false;
}

bool write_wav_pcm16_mono(const std::string& path, int sample_rate, const std::vector<float>& x)
{
	std::ofstream f(path, std::ios::binary);
	if (!f) return false;
	const uint32_t data_bytes = (uint32_t)(x.size() * 2);
	auto p16 = [&](uint32_t v) { char c[2] = {(char)(v & 255), (char)(v >> 8)}; f.write(c, 2); };
	auto p32 = [&](uint32_t v) { p16(v & 0xffff); p16(v >> 16); };
	f.write("RIFF", 4); p32(36 + data_bytes); f.write("WAVE", 4);
	f.write("fmt ", 4); p32(16); p16(1); p16(1); p32((uint32_t)sample_rate); p32((uint32_t)sample_rate * 2); p16(2); p16(16);
	f.write("data", 4); p32(data_bytes);
	for (float s : x) {
		long v = lroundf(s * 32767.f);
		p16((uint32_t)(uint16_t)(int16_t)std::max(-32768L, std::min(32767L, v)));
	}
	return (bool)f;
}

std::vector<float> to_mono(const Wav& w)
{
	if (w.channels == 2) {
		std::vector<float> m(w.samples.size() / 2);
		for (size_t i = 0, j = 0; i + 1 < w.samples.size(); i += 2, ++j)
			m[j] = (w.samples[i] + w.samples[i + 1]) / 2.0f;
		return m;
	}
	return w.samples;
}

// divide by max(|min|, max); an all-zero signal becomes NaN exactly as in the reference (offline.h:182-191)
void peak_normalize(std::vector<float>& x)
{
	if (x.empty()) return;
	auto lim = std::minmax_element(x.begin(), x.end());
	float real_max = std::max(-1 * (*lim.first), *lim.second);
	for (auto& v : x) v /= real_max;
}

struct Args {
	std::string cmd, infile, out;
	bool do_hps = false, cpu = false, sse = false, only_perc = false, soft = false, nocopybord = false;
	std::vector<std::string> hps_vals;
};

bool is_number(const std::string& s)
{
	if (s.empty()) return false;
	char* end = nullptr;
	std::strtod(s.c_str(), &end);
	return end && *end == 0;
}

bool parse(int argc, char** argv, Args& a, std::string& err)
{
	if (argc < 2) {
		err = "missing command";
		return false;
	}
	a.cmd = argv[1];
	for (int i = 2; i < argc; ++i) {
		std::string t = argv[i];
		if (t == "-i" || t == "--input") {
			if (++i >= argc) { err = "missing value for " + t; return false; }
			a.infile = argv[i];
		}
		else if (t == "-o" || t == "--out-prefix" || t == "--output") {
			if (++i >= argc) { err = "missing value for " + t; return false; }
			a.out = argv[i];
		}
		else if (t == "--hps") {
			a.do_hps = true;
			const size_t max_vals = a.cmd == "offline" ? 4 : 2;
			while (i + 1 < argc && a.hps_vals.size() < max_vals && is_number(argv[i + 1]))
				a.hps_vals.push_back(argv[++i]);
		}
		else if (t == "--cpu") a.cpu = true;
		else if (t == "--sse") a.sse = true;
		else if (t == "--only-percussive" && a.cmd == "offline") a.only_perc = true;
		else if (t == "--soft-mask") a.soft = true;
		else if (t == "--nocopybord") a.nocopybord = true;
		else {
			err = "unknown argument " + t;
			return false;
		}
	}
	return true;
}

void usage()
{
	std::cout << "SYNOPSIS\n"
	             "  zen offline (-i|--input) <infile> [--hps [<hop-h>] [<beta-h>] [<hop-p>] [<beta-p>]] [(-o|--out-prefix) <outfile_prefix>]\n"
	             "              [--cpu] [--sse] [--only-percussive] [--soft-mask] [--nocopybord]\n"
	             "  zen fakert (-i|--input) <infile> [--hps [<hop>] [<beta>]] [(-o|--output) <outfile>] [--cpu] [--sse] [--soft-mask] [--nocopybord]\n"
	             "  zen (help|-h|--help)\n"
	             "  zen (version|-v|--version)\n\n"
	             "OPTIONS\n"
	             "  offline (process entire songs at a time): 2-pass HPR-iterative, defaults: harmonic=4096,2.0 percussive=256,2.0\n"
	             "  fakert (use slim rt algorithms with wav files): 1-pass P-realtime, defaults: 256,2.0\n";
}

int refuse_cpu()
{
	std::cerr << "zen (zen_b200 build): --cpu is not available, this build provides the GPU path only" << std::endl;
	return 2;
}

void print_info(const Wav& w)
{
	const size_t frames = w.samples.size() / (size_t)std::max(1, w.channels);
	std::cout << "Audio file info:" << std::endl;
	std::cout << "\tsample rate: " << w.sample_rate << std::endl;
	std::cout << "\tlen samples: " << w.samples.size() << std::endl;
	std::cout << "\tframe size: " << w.channels * 2 << std::endl;
	std::cout << "\tseconds: " << (double)frames / w.sample_rate << std::endl;
	std::cout << "\tchannels: " << w.channels << std::endl;
}

int run_offline(const Args& a)
{
	std::size_t hop_h = 4096, hop_p = 256;  // zen/offline.h:28-31
	float beta_h = 2.0f, beta_p = 2.0f;
	if (a.hps_vals.size() > 0) hop_h = (std::size_t)std::atol(a.hps_vals[0].c_str());
	if (a.hps_vals.size() > 1) beta_h = (float)std::atof(a.hps_vals[1].c_str());
	if (a.hps_vals.size() > 2) hop_p = (std::size_t)std::atol(a.hps_vals[2].c_str());
	if (a.hps_vals.size() > 3) beta_p = (float)std::atof(a.hps_vals[3].c_str());
	std::cout << "Running zen-offline with the following params:"
	          << "\n\tinfile: " << a.infile << "\n\toutfile_prefix: " << a.out << "\n\tdo hps: " << (a.do_hps ? "yes" : "no")
	          << "\n\t\tharmonic hop: " << hop_h << "\n\t\tharmonic beta: " << beta_h << "\n\t\tpercussive hop: " << hop_p
	          << "\n\t\tpercussive beta: " << beta_p << "\n\tcpu: " << (a.cpu ? "yes" : "no") << "\n\tsse: " << (a.sse ? "yes" : "no")
	          << "\n\tsoft mask: " << (a.soft ? "yes" : "no") << "\n\tnocopybord: " << (a.nocopybord ? "yes" : "no") << std::endl;
	if (a.infile.empty()) {
		std::cerr << "offline params error" << std::endl;
		return 1;
	}
	if (a.cpu) return refuse_cpu();
	Wav w;
	std::string err;
	if (!read_wav(a.infile, w, err)) {
		std::cerr << "zen: " << err << std::endl;
		return 1;
	}
	print_info(w);
	std::vector<float> audio = to_mono(w);
	std::array<std::vector<float>, 3> all_out;
	if (a.do_hps) {
		std::cout << "Processing input signal of size " << audio.size() << " with HPR-I separation using harmonic params: " << hop_h << ","
		          << beta_h << ", percussive params: " << hop_p << "," << beta_p << std::endl;
		auto hpss = zen::hps::HPRIOffline<zen::Backend::GPU>((float)w.sample_rate, hop_h, hop_p, beta_h, beta_p, a.nocopybord);
		if (a.sse) hpss.use_sse_filter();
		if (a.soft) hpss.use_soft_mask();
		auto t1 = std::chrono::high_resolution_clock::now();
		all_out = hpss.process(audio);
		auto t2 = std::chrono::high_resolution_clock::now();
		auto dur = std::chrono::duration_cast<std::chrono::milliseconds>(t2 - t1).count();
		std::cout << "GPU/CUDA/thrust: 2-pass HPR-I-Offline took " << dur << " ms" << std::endl;
	}
	else {
		all_out = {audio, audio, audio};
	}
	if (!a.out.empty()) {
		const char* suffix[3] = {"_harm.wav", "_perc.wav", "_residual.wav"};
		for (int i = 0; i < 3; ++i) {
			if (a.only_perc && i != 1) continue;
			peak_normalize(all_out[i]);
			if (!write_wav_pcm16_mono(a.out + suffix[i], w.sample_rate, all_out[i])) {
				std::cerr << "zen: cannot write " << a.out + suffix[i] << std::endl;
				return 1;
			}
		}
	}
	return 0;
}

int run_fakert(const Args& a)
{
	std::size_t hop = 256;  // zen/fakert.h:47-48
	float beta = 2.0f;
	if (a.hps_vals.size() > 0) hop = (std::size_t)std::atol(a.hps_vals[0].c_str());
	if (a.hps_vals.size() > 1) beta = (float)std::atof(a.hps_vals[1].c_str());
	std::cout << "Running zen-fakert with the following params:"
	          << "\n\tinfile: " << a.infile << "\n\toutfile: " << a.out << "\n\tdo hps: " << (a.do_hps ? "yes" : "no") << "\n\t\thop: " << hop
	          << "\n\t\tbeta: " << beta << "\n\tcpu: " << (a.cpu ? "yes" : "no") << "\n\tsse: " << (a.sse ? "yes" : "no")
	          << "\n\tsoft mask: " << (a.soft ? "yes" : "no") << "\n\tnocopybord: " << (a.nocopybord ? "yes" : "no") << std::endl;
	if (a.infile.empty()) {
		std::cerr << "fakert params error" << std::endl;
		return 1;
	}
	if (a.cpu) return refuse_cpu();
	Wav w;
	std::string err;
	if (!read_wav(a.infile, w, err)) {
		std::cerr << "zen: " << err << std::endl;
		return 1;
	}
	print_info(w);
	std::vector<float> audio = to_mono(w);
	std::vector<float> percussive_out = audio;  // the unprocessed tail stays raw input (fakert.h:132)

	// get_chunk_limits (fakert.h:15-34): full hops strictly before size - hop; the last chunk is never produced
	std::vector<std::size_t> starts;
	if (audio.size() > hop)
		for (std::size_t i = 0; i < audio.size() - hop; i += hop)
			starts.push_back(i);
	std::cout << "Slicing buffer size " << audio.size() << " into " << starts.size() << " chunks of size " << hop << std::endl;
	float delta_t = 1000 * (float)hop / (float)w.sample_rate;

	auto hpss = zen::hps::HPRRealtime<zen::Backend::GPU>((float)w.sample_rate, hop, beta, zen::hps::OUTPUT_PERCUSSIVE, a.nocopybord);
	auto io = zen::io::IOGPU(hop);
	if (a.sse) hpss.use_sse_filter();
	if (a.soft) hpss.use_soft_mask();
	hpss.warmup(io);
	// the hops are served by the resident kernel (10 us instead of 28 us per hop at hop 1024); ZEN_RESIDENT=0: one launch per call
	const char* res_env = std::getenv("ZEN_RESIDENT");
	const bool resident = !(res_env && res_env[0] == '0');
	if (resident) hpss.start_resident();

	float iters = 0.0F;
	int time_tot = 0;
	std::size_t n = 0;
	for (std::size_t s : starts) {
		auto t1 = std::chrono::high_resolution_clock::now();
		if (a.do_hps) {
			std::copy(audio.begin() + s, audio.begin() + s + hop, io.host_in);
			hpss.process_next_hop(io.device_in);
			hpss.copy_percussive(io.device_out);
			std::copy(io.host_out, io.host_out + hop, percussive_out.begin() + n);
		}
		else {
			std::copy(audio.begin() + s, audio.begin() + s + hop, percussive_out.begin() + n);
		}
		auto t2 = std::chrono::high_resolution_clock::now();
		time_tot += (int)std::chrono::duration_cast<std::chrono::microseconds>(t2 - t1).count();
		n += hop;
		iters += 1.0F;
	}
	std::cout << "PRealtime GPU:  Δn = " << hop << ", Δt(ms) = " << delta_t << ", average processing duration(us) = " << (float)time_tot / iters
	          << std::endl;
	if (!a.out.empty()) {
		peak_normalize(percussive_out);
		if (!write_wav_pcm16_mono(a.out, w.sample_rate, percussive_out)) {
			std::cerr << "zen: cannot write " << a.out << std::endl;
			return 1;
		}
	}
	return 0;
}

}  // namespace

int main(int argc, char** argv)
{
	Args a;
	std::string err;
	if (!parse(argc, argv, a, err)) {
		std::cerr << "zen: " << err << "\n";
		usage();
		return 1;
	}
	try {
		if (a.cmd == "offline") return run_offline(a);
		if (a.cmd == "fakert") return run_fakert(a);
		if (a.cmd == "help" || a.cmd == "-h" || a.cmd == "--help") {
			usage();
			return 0;
		}
		if (a.cmd == "version" || a.cmd == "-v" || a.cmd == "--version") {
			std::cout << "version 1.0\n";
			return 0;
		}
	}
	catch (const zen::ZgException& e) {
		std::cerr << "zen: " << e.what() << std::endl;
		return 1;
	}
	catch (const std::exception& e) {
		std::cerr << "zen: " << e.what() << std::endl;
		return 1;
	}
	std::cerr << "zen: unknown command " << a.cmd << "\n";
	usage();
	return 1;
}
