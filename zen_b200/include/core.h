// zen_b200 drop-in for libzen's internal <core.h> (reference: libzen/core.h:17-41):
// the backend policy table.  Only Backend::GPU is provided by this library.
#ifndef ZEN_B200_CORE_H
#define ZEN_B200_CORE_H

#include <box.h>
#include <fftw.h>
#include <libzen/zen.h>
#include <mfilt.h>
#include <thrust/complex.h>
#include <thrust/device_vector.h>
#include <win.h>

namespace zen {
namespace internal {
	namespace core {

		template <zen::Backend T>
		struct TypeTraits;

		template <>
		struct TypeTraits<zen::Backend::GPU> {
			using InputPointer = thrust::device_ptr<float>;
			using RealVector = thrust::device_vector<float>;
			using ComplexVector = thrust::device_vector<thrust::complex<float>>;
			using FFTC2CWrapper = zen::internal::fftw::FFTC2CWrapperGPU;
			using MedianFilter = zen::internal::hps::mfilt::MedianFilterGPU;
			using BoxFilter = zen::internal::hps::box::BoxFilterGPU;
			using Window = zen::internal::win::WindowGPU;
		};

	}  // namespace core
}  // namespace internal
}  // namespace zen

#endif
