// zen_b200 drop-in for libzen's internal <mfilt.h> (reference: libzen/mfilt.h).
// MedianFilterGPU keeps the constructor signature, the ZgException rule and the
// NPP window / border semantics (bit-exact), but owns no scratch: the kernels in
// zen_b200/csrc/filters.cu do the wrap border by index arithmetic.
#ifndef ZEN_B200_MFILT_H
#define ZEN_B200_MFILT_H

#include <thrust/device_ptr.h>
#include <thrust/device_vector.h>

#include <libzen/zen.h>

namespace zen {
namespace internal {
	namespace hps {
		namespace mfilt {

			enum MedianFilterDirection {
				TimeCausal = ZEN_TIME_CAUSAL,
				TimeAnticausal = ZEN_TIME_ANTICAUSAL,
				Frequency = ZEN_FREQUENCY,
			};

			namespace detail {
				inline void check_len(int time, int frequency, int filter_len, MedianFilterDirection dir, const char* what)
				{
					const bool along_time = dir == TimeCausal || dir == TimeAnticausal;
					if ((along_time && filter_len > time) || (dir == Frequency && filter_len > frequency))
						throw zen::ZgException(what);
				}
			}  // namespace detail

			class MedianFilterGPU {
			public:
				MedianFilterDirection mydir;
				int time;
				int frequency;
				int filter_len;  // as passed (the reference also keeps the pre-odd value here)
				int filter_mid;
				bool copy_bord;

				MedianFilterGPU(int time, int frequency, int filter_len, MedianFilterDirection dir, bool copy_bord = false)
				    : mydir(dir)
				    , time(time)
				    , frequency(frequency)
				    , filter_len(filter_len)
				    , filter_mid((filter_len + (1 - filter_len % 2)) / 2)
				    , copy_bord(copy_bord)
				{
					detail::check_len(time, frequency, filter_len, dir, "median filter bigger than matrix dimension");
				}

				void filter(thrust::device_vector<float>& src, thrust::device_vector<float>& dst) { filter(src.data(), dst.data()); }

				void filter(thrust::device_ptr<float> src, thrust::device_ptr<float> dst)
				{
					zen::b200_detail::check(zen_median_filter(time, frequency, filter_len, (int)mydir, copy_bord ? 1 : 0,
					                                          thrust::raw_pointer_cast(src), thrust::raw_pointer_cast(dst), nullptr),
					                        "MedianFilterGPU::filter");
				}
			};

		}  // namespace mfilt
	}  // namespace hps
}  // namespace internal
}  // namespace zen

#endif
