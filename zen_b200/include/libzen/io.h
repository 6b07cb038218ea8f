// zen_b200 drop-in for <libzen/io.h> (reference: libzen/libzen/io.h:16-81):
// mapped pinned host buffers and their device aliases, the way every caller of
// the real-time path hands audio in and out (zen/fakert.h:202, 225-234).
#ifndef ZEN_B200_PUB_IO_H
#define ZEN_B200_PUB_IO_H

#include <cstddef>

#include <thrust/device_ptr.h>

#include <libzen/zen.h>

namespace zen {
namespace io {

	class IOGPU {
	public:
		float* host_in = nullptr;
		float* host_out = nullptr;
		thrust::device_ptr<float> device_in;
		thrust::device_ptr<float> device_out;
		std::size_t size = 0;

		explicit IOGPU(std::size_t n)
		    : size(n)
		{
			zen::b200_detail::check(zen_io_alloc(&io_, n), "IOGPU");
			host_in = io_.host_in;
			host_out = io_.host_out;
			device_in = thrust::device_pointer_cast(io_.device_in);
			device_out = thrust::device_pointer_cast(io_.device_out);
		}
		~IOGPU() { zen_io_free(&io_); }
		IOGPU(const IOGPU&) = delete;
		IOGPU& operator=(const IOGPU&) = delete;
		IOGPU(IOGPU&& o) noexcept { steal(o); }

	private:
		zen_io io_{};
		void steal(IOGPU& o)
		{
			io_ = o.io_;
			host_in = o.host_in;
			host_out = o.host_out;
			device_in = o.device_in;
			device_out = o.device_out;
			size = o.size;
			o.io_ = zen_io{};
		}
	};

}  // namespace io
}  // namespace zen

#endif
