// zen_b200 drop-in for libzen's public <libzen/hps.h> (reference:
// libzen/libzen/hps.h:25-118, method bodies libzen/hps.cu:21-427).
// Same constructors, defaults and call surface; Backend::GPU only — this
// library has no CPU path, so the Backend::CPU instantiations throw.
#ifndef ZEN_B200_PUB_HPS_H
#define ZEN_B200_PUB_HPS_H

#include <array>
#include <cstddef>
#include <numeric>
#include <vector>

#include <thrust/device_ptr.h>

#include <libzen/io.h>
#include <libzen/zen.h>

namespace zen {
namespace internal {
	namespace hps {
		template <zen::Backend B>
		class HPR;
	}
}  // namespace internal

namespace hps {

	const unsigned int OUTPUT_HARMONIC = 1;
	const unsigned int OUTPUT_PERCUSSIVE = 1 << 1;
	const unsigned int OUTPUT_RESIDUAL = 1 << 2;

	// Driedger's iterative two-pass HPR on a whole signal.
	template <zen::Backend B>
	class HPRIOffline {
	public:
		HPRIOffline(float fs, std::size_t hop_h, std::size_t hop_p, float beta_h, float beta_p, bool nocopybord)
		    : fs_(fs)
		    , hop_h_(hop_h)
		    , hop_p_(hop_p)
		    , beta_h_(beta_h)
		    , beta_p_(beta_p)
		    , options_(nocopybord ? ZEN_OPT_NOCOPYBORD : 0)
		{
			static_assert(B == zen::Backend::GPU, "zen_b200 provides Backend::GPU only");
			if (hop_p == 0 || hop_h % hop_p != 0)
				throw zen::ZgException("hop_h and hop_p should be evenly divisible");
		}
		HPRIOffline(float fs, std::size_t hop_h, std::size_t hop_p, float beta_h, float beta_p)
		    : HPRIOffline(fs, hop_h, hop_p, beta_h, beta_p, false)
		{
		}
		HPRIOffline(float fs, std::size_t hop_h, std::size_t hop_p)
		    : HPRIOffline(fs, hop_h, hop_p, 2.0, 2.0)
		{
		}
		explicit HPRIOffline(float fs)
		    : HPRIOffline(fs, 4096, 256, 2.0, 2.0)
		{
		}

		// whole signal in, {harmonic, percussive, residual} of the same length out
		std::array<std::vector<float>, 3> process(std::vector<float> audio)
		{
			std::array<std::vector<float>, 3> out;
			for (auto& v : out)
				v.assign(audio.size(), 0.0F);
			if (audio.empty())
				return out;
			zen::b200_detail::check(zen_offline_process(fs_, (int)hop_h_, (int)hop_p_, beta_h_, beta_p_, options_, audio.data(),
			                                            (long)audio.size(), out[0].data(), out[1].data(), out[2].data()),
			                        "HPRIOffline::process");
			return out;
		}

		void use_sse_filter() { options_ |= ZEN_OPT_SSE; }
		void use_soft_mask() { options_ |= ZEN_OPT_SOFT_MASK; }

	private:
		float fs_;
		std::size_t hop_h_, hop_p_;
		float beta_h_, beta_p_;
		int options_;
	};

	// One causal stream, hop by hop.
	template <zen::Backend B>
	class HPRRealtime {
	public:
		HPRRealtime(float fs, std::size_t hop, float beta, unsigned int output_flags, bool nocopybord);
		HPRRealtime(float fs, std::size_t hop, float beta, unsigned int output_flags)
		    : HPRRealtime(fs, hop, beta, output_flags, false)
		{
		}
		HPRRealtime(float fs, std::size_t hop, unsigned int output_flags)
		    : HPRRealtime(fs, hop, 2.0, output_flags)
		{
		}
		HPRRealtime(float fs, unsigned int output_flags)
		    : HPRRealtime(fs, 256, 2.0, output_flags)
		{
		}
		~HPRRealtime();
		HPRRealtime(const HPRRealtime&) = delete;
		HPRRealtime& operator=(const HPRRealtime&) = delete;
		HPRRealtime(HPRRealtime&& o) noexcept
		    : p_impl(o.p_impl)
		{
			o.p_impl = nullptr;
		}

		void process_next_hop(thrust::device_ptr<float> in);
		void copy_harmonic(thrust::device_ptr<float> out);
		void copy_percussive(thrust::device_ptr<float> out);
		void copy_residual(thrust::device_ptr<float> out);

		// zen_b200 extension: process one hop and write the enabled outputs (null = skip) in a single launch
		void process_next_hop(thrust::device_ptr<float> in, thrust::device_ptr<float> out_harmonic,
		                      thrust::device_ptr<float> out_percussive, thrust::device_ptr<float> out_residual);

		void warmup(zen::io::IOGPU& io);

		// zen_b200 extension: serve the following hops from a resident (persistent) kernel that keeps the stream
		// state in shared memory and is driven through a doorbell in mapped memory: no launch per hop
		void start_resident();
		void stop_resident();

		void use_sse_filter();
		void use_soft_mask();

	private:
		zen::internal::hps::HPR<B>* p_impl;
	};

}  // namespace hps
}  // namespace zen

#include <hps.h>

namespace zen {
namespace hps {

	template <zen::Backend B>
	HPRRealtime<B>::HPRRealtime(float fs, std::size_t hop, float beta, unsigned int output_flags, bool nocopybord)
	    : p_impl(new zen::internal::hps::HPR<B>(fs, hop, beta, output_flags,
	                                            zen::internal::hps::mfilt::MedianFilterDirection::TimeCausal, !nocopybord))
	{
	}

	template <zen::Backend B>
	HPRRealtime<B>::~HPRRealtime()
	{
		delete p_impl;
	}

	template <zen::Backend B>
	void HPRRealtime<B>::process_next_hop(thrust::device_ptr<float> in)
	{
		// the hop has been read when this returns, as in the reference (hps.cu:452-453): callers refill `in` right away
		// (zen/fakert.h:225-229).  The outputs are waited for by the copy_* call that follows.
		zen::b200_detail::check(zen_hpr_process_next_hop(p_impl->handle(), thrust::raw_pointer_cast(in)), "process_next_hop");
		zen::b200_detail::check(zen_hpr_wait_input_consumed(p_impl->handle()), "process_next_hop");
	}

	template <zen::Backend B>
	void HPRRealtime<B>::process_next_hop(thrust::device_ptr<float> in, thrust::device_ptr<float> oh, thrust::device_ptr<float> op,
	                                      thrust::device_ptr<float> orr)
	{
		zen::b200_detail::check(zen_hpr_process_hop_io(p_impl->handle(), thrust::raw_pointer_cast(in), thrust::raw_pointer_cast(oh),
		                                               thrust::raw_pointer_cast(op), thrust::raw_pointer_cast(orr)),
		                        "process_next_hop");
		zen::b200_detail::check(zen_hpr_synchronize(p_impl->handle()), "process_next_hop");
	}

	template <zen::Backend B>
	void HPRRealtime<B>::copy_harmonic(thrust::device_ptr<float> out)
	{
		zen::b200_detail::check(zen_hpr_copy_harmonic(p_impl->handle(), thrust::raw_pointer_cast(out)), "copy_harmonic");
	}

	template <zen::Backend B>
	void HPRRealtime<B>::copy_percussive(thrust::device_ptr<float> out)
	{
		zen::b200_detail::check(zen_hpr_copy_percussive(p_impl->handle(), thrust::raw_pointer_cast(out)), "copy_percussive");
	}

	template <zen::Backend B>
	void HPRRealtime<B>::copy_residual(thrust::device_ptr<float> out)
	{
		zen::b200_detail::check(zen_hpr_copy_residual(p_impl->handle(), thrust::raw_pointer_cast(out)), "copy_residual");
	}

	// 1000 hops of iota data, then reset (hps.cu:392-409)
	template <zen::Backend B>
	void HPRRealtime<B>::warmup(zen::io::IOGPU& io)
	{
		const int test_iters = 1000;
		const std::size_t hop = p_impl->hop;
		for (int i = 0; i < test_iters; ++i) {
			std::iota(io.host_in, io.host_in + hop, (float)(i * hop));
			p_impl->process_next_hop(io.device_in);
		}
		p_impl->reset_buffers();
	}

	template <zen::Backend B>
	void HPRRealtime<B>::start_resident()
	{
		zen::b200_detail::check(zen_hpr_realtime_begin(p_impl->handle()), "start_resident");
	}

	template <zen::Backend B>
	void HPRRealtime<B>::stop_resident()
	{
		zen::b200_detail::check(zen_hpr_realtime_end(p_impl->handle()), "stop_resident");
	}

	template <zen::Backend B>
	void HPRRealtime<B>::use_sse_filter()
	{
		p_impl->use_sse_filter();
	}

	template <zen::Backend B>
	void HPRRealtime<B>::use_soft_mask()
	{
		p_impl->use_soft_mask();
	}

}  // namespace hps
}  // namespace zen

#endif
