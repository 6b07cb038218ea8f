// zen_b200 drop-in for libzen's internal <hps.h> (reference: libzen/hps.h).
// zen::internal::hps::HPR<Backend::GPU> keeps the reference's public members by
// name.  input and the three *_out vectors ARE the kernel's streaming state
// (bound with zen_hpr_bind_state), so they are current after every
// process_next_hop, as hps.test.cu expects.  The stft_width x nfft matrices
// (sliding_stft, s_mag, reciprocal, harmonic_matrix, percussive_matrix, masks)
// are no longer recomputed every hop — the fused kernel only computes the row
// the reference consumes — and are refreshed by refresh_matrices(), or after
// every hop when sync_matrices is set.
#ifndef ZEN_B200_HPS_INTERNAL_H
#define ZEN_B200_HPS_INTERNAL_H

#include <cmath>
#include <cstddef>
#include <limits>

#include <thrust/complex.h>
#include <thrust/device_vector.h>
#include <thrust/fill.h>

#include <box.h>
#include <core.h>
#include <fftw.h>
#include <libzen/hps.h>
#include <mfilt.h>
#include <win.h>

namespace zen {
namespace internal {
	namespace hps {

		static constexpr float Eps = std::numeric_limits<float>::epsilon();

		// Element-wise functors of the reference (hps.h:25-150), kept because its tests use
		// them directly (box.test.cu:128-134).  The kernels have their own fused forms.
		struct window_functor {
			__host__ __device__ thrust::complex<float> operator()(const float& sample, const float& w) const
			{
				return thrust::complex<float>{sample * w, 0.0F};
			}
		};
		struct residual_mask_functor {
			__host__ __device__ float operator()(const float& mh, const float& mp) const { return 1 - (mh + mp); }
		};
		struct reciprocal_functor {
			const float factor;
			explicit reciprocal_functor(float f)
			    : factor(f)
			{
			}
			__host__ __device__ float operator()(const float& x) const { return (1.0f / x) * factor; }
		};
		struct apply_mask_functor {
			__host__ __device__ thrust::complex<float> operator()(const thrust::complex<float>& bin, const float& m) const
			{
				return bin * m;
			}
		};
		struct overlap_add_functor {
			const float cola_factor;
			explicit overlap_add_functor(float c)
			    : cola_factor(c)
			{
			}
			__host__ __device__ float operator()(const thrust::complex<float>& y, const float& acc) const
			{
				return acc + y.real() * cola_factor;
			}
		};
		struct complex_abs_functor {
			template <typename V>
			__host__ __device__ V operator()(const thrust::complex<V>& z)
			{
				return thrust::abs(z);
			}
		};
		struct complex_abs_squared_functor {
			template <typename V>
			__host__ __device__ V operator()(const thrust::complex<V>& z)
			{
				return powf(thrust::abs(z), 2.0f);
			}
		};
		struct hard_mask_functor {
			const float beta;
			explicit hard_mask_functor(float b)
			    : beta(b)
			{
			}
			__host__ __device__ float operator()(const float& a, const float& b) const { return float((a / (b + Eps)) >= beta); }
		};
		struct soft_mask_functor {
			const int power;
			explicit soft_mask_functor(int p)
			    : power(p)
			{
			}
			__host__ __device__ float operator()(const float& a, const float& b) const
			{
				return float(powf(a, power) / (powf(a, power) + powf(b, power) + Eps));
			}
		};
		struct sse_mask_functor {
			__host__ __device__ float operator()(const float& a, const float& b) const { return float(a * a / (a * a + b * b + Eps)); }
		};
		struct sum_vectors_functor {
			__host__ __device__ float operator()(const float& a, const float& b) const { return a + b; }
		};

		template <zen::Backend B>
		class HPR;

		template <>
		class HPR<zen::Backend::GPU> {
			using Traits = zen::internal::core::TypeTraits<zen::Backend::GPU>;

		public:
			using InputPointer = Traits::InputPointer;
			using RealVector = Traits::RealVector;
			using ComplexVector = Traits::ComplexVector;

			float fs;
			std::size_t hop, nwin, nfft;
			float beta;
			int l_harm, l_perc, lag;
			std::size_t stft_width;

			RealVector input;
			Traits::Window window;
			ComplexVector sliding_stft;
			RealVector s_mag, reciprocal, harmonic_matrix, percussive_matrix;
			RealVector percussive_mask, harmonic_mask, residual_mask;
			RealVector percussive_out, harmonic_out, residual_out;
			float COLA_factor;
			Traits::MedianFilter time, frequency;
			Traits::BoxFilter time_sse, frequency_sse;
			Traits::FFTC2CWrapper fft;
			bool output_percussive, output_harmonic, output_residual;
			bool use_sse, soft_mask;
			bool sync_matrices = false;  // zen_b200 extension: refresh the debug matrices after every hop

			HPR(float fs, std::size_t hop, float beta, unsigned int output_flags, mfilt::MedianFilterDirection causality, bool copy_bord)
			    : HPR(geom_of(fs, hop, causality), fs, beta, output_flags, causality, copy_bord)
			{
			}

			~HPR() { zen_hpr_destroy(h_); }
			HPR(const HPR&) = delete;
			HPR& operator=(const HPR&) = delete;

			void use_sse_filter()
			{
				use_sse = true;
				zen::b200_detail::check(zen_hpr_use_sse_filter(h_), "use_sse_filter");
			}
			void use_soft_mask()
			{
				soft_mask = true;
				zen::b200_detail::check(zen_hpr_use_soft_mask(h_), "use_soft_mask");
			}

			// one hop; blocking, like the reference's chain of thrust calls
			void process_next_hop(InputPointer in_hop)
			{
				zen::b200_detail::check(zen_hpr_process_next_hop(h_, thrust::raw_pointer_cast(in_hop)), "process_next_hop");
				zen::b200_detail::check(zen_hpr_synchronize(h_), "process_next_hop");
				if (sync_matrices)
					refresh_matrices();
			}

			// the reference's two per-hop stages are fused into process_next_hop; standalone
			// they only bring the public matrices up to date
			void apply_median_filter() { refresh_matrices(); }
			void apply_sse_filter() { refresh_matrices(); }

			void refresh_matrices()
			{
				auto raw = [](RealVector& v) { return thrust::raw_pointer_cast(v.data()); };
				zen::b200_detail::check(
				    zen_hpr_materialize(h_, reinterpret_cast<float*>(thrust::raw_pointer_cast(sliding_stft.data())), raw(s_mag),
				                        raw(harmonic_matrix), raw(percussive_matrix), raw(harmonic_mask), raw(percussive_mask),
				                        raw(residual_mask)),
				    "refresh_matrices");
			}

			void reset_buffers()
			{
				zen::b200_detail::check(zen_hpr_reset_buffers(h_), "reset_buffers");
				thrust::fill(fft.fft_vec.begin(), fft.fft_vec.end(), thrust::complex<float>{0.0F, 0.0F});
				thrust::fill(sliding_stft.begin(), sliding_stft.end(), thrust::complex<float>{0.0F, 0.0F});
				for (RealVector* v : {&s_mag, &reciprocal, &harmonic_matrix, &percussive_matrix, &harmonic_mask, &percussive_mask,
				                      &residual_mask})
					thrust::fill(v->begin(), v->end(), 0.0F);
			}

			zen_hpr* handle() { return h_; }

		private:
			zen_hpr* h_ = nullptr;

			static zen_geometry geom_of(float fs, std::size_t hop, mfilt::MedianFilterDirection causality)
			{
				zen_geometry g;
				zen::b200_detail::check(zen_hpr_geometry(fs, (int)hop, causality == mfilt::TimeCausal, &g), "HPR geometry");
				return g;
			}

			HPR(const zen_geometry& g, float fs, float beta, unsigned int output_flags, mfilt::MedianFilterDirection causality,
			    bool copy_bord)
			    : fs(fs)
			    , hop(g.hop)
			    , nwin(g.nwin)
			    , nfft(g.nfft)
			    , beta(beta)
			    , l_harm(g.l_harm)
			    , l_perc(g.l_perc)
			    , lag(g.lag)
			    , stft_width(g.stft_width)
			    , input(g.nwin, 0.0F)
			    , window(win::WindowType::SqrtVonHann, g.nwin)
			    , sliding_stft((std::size_t)g.stft_width * g.nfft, thrust::complex<float>{0.0F, 0.0F})
			    , s_mag((std::size_t)g.stft_width * g.nfft, 0.0F)
			    , reciprocal((std::size_t)g.stft_width * g.nfft, 0.0F)
			    , harmonic_matrix((std::size_t)g.stft_width * g.nfft, 0.0F)
			    , percussive_matrix((std::size_t)g.stft_width * g.nfft, 0.0F)
			    , percussive_mask((std::size_t)g.stft_width * g.nfft, 0.0F)
			    , harmonic_mask((std::size_t)g.stft_width * g.nfft, 0.0F)
			    , residual_mask((std::size_t)g.stft_width * g.nfft, 0.0F)
			    , percussive_out(g.nwin, 0.0F)
			    , harmonic_out(g.nwin, 0.0F)
			    , residual_out(g.nwin, 0.0F)
			    , COLA_factor(g.cola_factor)
			    , time(g.stft_width, g.nfft, g.l_harm, causality, copy_bord)
			    , frequency(g.stft_width, g.nfft, g.l_perc, mfilt::MedianFilterDirection::Frequency, copy_bord)
			    , time_sse(g.stft_width, g.nfft, g.l_harm, causality)
			    , frequency_sse(g.stft_width, g.nfft, g.l_perc, mfilt::MedianFilterDirection::Frequency)
			    , fft(g.nfft)
			    , output_percussive((output_flags & zen::hps::OUTPUT_PERCUSSIVE) != 0)
			    , output_harmonic((output_flags & zen::hps::OUTPUT_HARMONIC) != 0)
			    , output_residual((output_flags & zen::hps::OUTPUT_RESIDUAL) != 0)
			    , use_sse(false)
			    , soft_mask(false)
			{
				zen::b200_detail::check(zen_hpr_create(&h_, fs, g.hop, beta, output_flags, (int)causality, copy_bord ? 1 : 0), "HPR");
				auto raw = [](RealVector& v) { return thrust::raw_pointer_cast(v.data()); };
				zen::b200_detail::check(zen_hpr_bind_state(h_, raw(input), raw(harmonic_out), raw(percussive_out), raw(residual_out)),
				                        "HPR state");
			}
		};

	}  // namespace hps
}  // namespace internal
}  // namespace zen

#endif
