// zen_b200 drop-in for libzen's internal <box.h> (reference: libzen/box.h:30-215):
// the SSE moving-average filter, always wrap-padded like the reference's
// nppiCopyWrapBorder + nppiFilterBox pair.
#ifndef ZEN_B200_BOX_H
#define ZEN_B200_BOX_H

#include <mfilt.h>

// the reference's box.h pulls the direction enum into the global namespace (box.h:24)
using namespace zen::internal::hps::mfilt;

namespace zen {
namespace internal {
	namespace hps {
		namespace box {

			class BoxFilterGPU {
			public:
				MedianFilterDirection mydir;
				int time;
				int frequency;
				int filter_len;

				BoxFilterGPU(int time, int frequency, int filter_len, MedianFilterDirection dir)
				    : mydir(dir)
				    , time(time)
				    , frequency(frequency)
				    , filter_len(filter_len)
				{
					mfilt::detail::check_len(time, frequency, filter_len, dir, "box filter bigger than matrix dimension");
				}

				void filter(thrust::device_vector<float>& src, thrust::device_vector<float>& dst) { filter(src.data(), dst.data()); }

				void filter(thrust::device_ptr<float> src, thrust::device_ptr<float> dst)
				{
					zen::b200_detail::check(zen_box_filter(time, frequency, filter_len, (int)mydir, thrust::raw_pointer_cast(src),
					                                       thrust::raw_pointer_cast(dst), nullptr),
					                        "BoxFilterGPU::filter");
				}
			};

		}  // namespace box
	}  // namespace hps
}  // namespace internal
}  // namespace zen

#endif
