// TEST INFRASTRUCTURE - Backend::CPU classes for compiling the reference's own libzen/*.test.cu unchanged.
//
// zen_b200 ships Backend::GPU only (north_star: no CPU path in the product).  The reference's test sources,
// however, instantiate the IPP-backed CPU classes next to the GPU ones in the same translation unit
// (libzen/mfilt.test.cu:375-591, libzen/hps.test.cu:55-149, libzen/fftw.test.cu:20-101), so they cannot even be
// compiled without them.  This header, force-included by oracle/Makefile (target reftests) and by nothing else,
// supplies them on top of the plain-C oracle (oracle/hpr_oracle.h, ZO_GEOM_CPU = the IPP semantics: centred
// window, replicated border).  Interfaces follow libzen/mfilt.h:270-342, libzen/box.h:217-288,
// libzen/fftw.h:51-129 and libzen/hps.h:152-322.
#ifndef ZEN_B200_TEST_CPU_BACKEND_H
#define ZEN_B200_TEST_CPU_BACKEND_H

#include <cmath>
#include <cstdlib>
#include <utility>
#include <vector>

#include <thrust/complex.h>

#include <hps.h>      // zen_b200/include: HPR<Backend::GPU>, functors, MedianFilterGPU ...
#include <hpr_oracle.h>

namespace zen {
namespace internal {
	namespace hps {
		namespace mfilt {
			class MedianFilterCPU {
			public:
				int time, frequency, filter_len;
				MedianFilterDirection mydir;
				MedianFilterCPU(int time, int frequency, int filter_len, MedianFilterDirection dir, bool copy_bord = false)
				    : time(time)
				    , frequency(frequency)
				    , filter_len(filter_len)
				    , mydir(dir)
				{
					(void)copy_bord;  // not used for CPU (mfilt.h:289)
					detail::check_len(time, frequency, filter_len, dir, "median filter bigger than matrix dimension");
				}
				void filter(std::vector<float>& src, std::vector<float>& dst)
				{
					zo_median_filter(ZO_GEOM_CPU, time, frequency, filter_len, (int)mydir, 0, src.data(), dst.data());
				}
			};
		}  // namespace mfilt
		namespace box {
			class BoxFilterCPU {
			public:
				int time, frequency, filter_len;
				mfilt::MedianFilterDirection mydir;
				BoxFilterCPU(int time, int frequency, int filter_len, mfilt::MedianFilterDirection dir)
				    : time(time)
				    , frequency(frequency)
				    , filter_len(filter_len)
				    , mydir(dir)
				{
					mfilt::detail::check_len(time, frequency, filter_len, dir, "box filter bigger than matrix dimension");
				}
				void filter(std::vector<float>& src, std::vector<float>& dst)
				{
					zo_box_filter(ZO_GEOM_CPU, time, frequency, filter_len, (int)mydir, src.data(), dst.data());
				}
			};
		}  // namespace box
	}  // namespace hps
	namespace fftw {
		// IPP's ippsFFT{Fwd,Inv}_CToC_32fc_I works in single precision (libzen/fftw.h:108-114); so does this stand-in
		// (radix-2, twiddles rounded from double), because fftw.test.cu:83-101 also compares which outputs overflow
		// on uniform(FLT_MIN, FLT_MAX) inputs - a double-precision evaluation would stay finite where any float FFT
		// does not.
		class FFTC2CWrapperCPU {
		public:
			std::size_t nfft;
			std::vector<thrust::complex<float>> fft_vec;
			explicit FFTC2CWrapperCPU(std::size_t nfft)
			    : nfft(nfft)
			    , fft_vec(nfft)
			    , tw_(nfft / 2 ? nfft / 2 : 1)
			{
				for (std::size_t k = 0; k < nfft / 2; ++k) {
					const double a = -6.283185307179586476925286766559 * (double)k / (double)nfft;
					tw_[k] = thrust::complex<float>((float)std::cos(a), (float)std::sin(a));
				}
			}
			void forward() { run(false); }
			void backward() { run(true); }

		private:
			std::vector<thrust::complex<float>> tw_;
			void run(bool inverse)
			{
				const std::size_t n = nfft;
				thrust::complex<float>* x = fft_vec.data();
				for (std::size_t i = 1, j = 0; i < n; ++i) {
					std::size_t bit = n >> 1;
					for (; j & bit; bit >>= 1) j ^= bit;
					j ^= bit;
					if (i < j) std::swap(x[i], x[j]);
				}
				for (std::size_t len = 2; len <= n; len <<= 1) {
					const std::size_t step = n / len;
					for (std::size_t i = 0; i < n; i += len)
						for (std::size_t k = 0; k < len / 2; ++k) {
							thrust::complex<float> w = tw_[k * step];
							if (inverse) w = thrust::conj(w);
							const thrust::complex<float> u = x[i + k];
							const float vr = x[i + k + len / 2].real() * w.real() - x[i + k + len / 2].imag() * w.imag();
							const float vi = x[i + k + len / 2].real() * w.imag() + x[i + k + len / 2].imag() * w.real();
							x[i + k] = thrust::complex<float>(u.real() + vr, u.imag() + vi);
							x[i + k + len / 2] = thrust::complex<float>(u.real() - vr, u.imag() - vi);
						}
				}
			}
		};
	}  // namespace fftw
	namespace hps {
		// HPR<Backend::CPU>: the members the reference's tests read (hps.test.cu:160-372), refreshed after every hop
		template <>
		class HPR<zen::Backend::CPU> {
		public:
			float fs;
			std::size_t hop, nwin, nfft;
			float beta;
			int l_harm, l_perc, lag;
			std::size_t stft_width;
			std::vector<float> input, percussive_out, harmonic_out, residual_out;
			float COLA_factor;
			bool output_percussive, output_harmonic, output_residual, use_sse, soft_mask;

			HPR(float fs, std::size_t hop, float beta, unsigned int output_flags, mfilt::MedianFilterDirection causality, bool copy_bord)
			    : fs(fs)
			    , hop(hop)
			    , beta(beta)
			    , output_percussive(output_flags & zen::hps::OUTPUT_PERCUSSIVE)
			    , output_harmonic(output_flags & zen::hps::OUTPUT_HARMONIC)
			    , output_residual(output_flags & zen::hps::OUTPUT_RESIDUAL)
			    , use_sse(false)
			    , soft_mask(false)
			{
				o_ = zo_hpr_create(ZO_GEOM_CPU, fs, (int)hop, beta, output_flags, (int)causality, copy_bord ? 1 : 0);
				if (!o_) throw zen::ZgException("invalid HPR geometry");
				zo_geom g;
				zo_hpr_geom(o_, &g);
				nwin = g.nwin;
				nfft = g.nfft;
				l_harm = g.l_harm;
				l_perc = g.l_perc;
				lag = g.lag;
				stft_width = g.stft_width;
				COLA_factor = g.cola;
				input.assign(nwin, 0.0f);
				percussive_out.assign(nwin, 0.0f);
				harmonic_out.assign(nwin, 0.0f);
				residual_out.assign(nwin, 0.0f);
			}
			~HPR() { zo_hpr_destroy(o_); }
			HPR(const HPR&) = delete;
			HPR& operator=(const HPR&) = delete;
			void use_sse_filter()
			{
				use_sse = true;
				zo_hpr_use_sse_filter(o_);
			}
			void use_soft_mask()
			{
				soft_mask = true;
				zo_hpr_use_soft_mask(o_);
			}
			void process_next_hop(float* in_hop)
			{
				zo_hpr_process_next_hop(o_, in_hop);
				pull();
			}
			void reset_buffers()
			{
				zo_hpr_reset_buffers(o_);
				pull();
			}

		private:
			void pull()
			{
				zo_hpr_get(o_, 6, harmonic_out.data());
				zo_hpr_get(o_, 7, percussive_out.data());
				zo_hpr_get(o_, 8, residual_out.data());
				zo_hpr_get(o_, 10, input.data());
			}
			zo_hpr* o_;
		};
	}  // namespace hps
}  // namespace internal
}  // namespace zen

#endif
