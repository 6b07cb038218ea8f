// zen_b200 drop-in for libzen's internal <fftw.h> (reference: libzen/fftw.h:20-49):
// in-place, unnormalised complex FFT on the wrapper's own fft_vec — no cuFFT plan.
#ifndef ZEN_B200_FFTW_H
#define ZEN_B200_FFTW_H

#include <cstddef>

#include <thrust/complex.h>
#include <thrust/device_vector.h>

#include <libzen/zen.h>

namespace zen {
namespace internal {
	namespace fftw {

		class FFTC2CWrapperGPU {
		public:
			std::size_t nfft;
			thrust::device_vector<thrust::complex<float>> fft_vec;

			explicit FFTC2CWrapperGPU(std::size_t nfft)
			    : nfft(nfft)
			    , fft_vec(nfft)
			{
			}

			void forward() { run(0); }
			void backward() { run(1); }

		private:
			void run(int inverse)
			{
				zen::b200_detail::check(
				    zen_fft_c2c((int)nfft, reinterpret_cast<float*>(thrust::raw_pointer_cast(fft_vec.data())), inverse, nullptr),
				    "FFTC2CWrapperGPU");
			}
		};

	}  // namespace fftw
}  // namespace internal
}  // namespace zen

#endif
