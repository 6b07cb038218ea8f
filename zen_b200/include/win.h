// zen_b200 drop-in for libzen's internal <win.h> (reference: libzen/win.h:21-53).
// The table is computed on the host by zen_window() with the reference's float
// expression, so it is bit-identical.
#ifndef ZEN_B200_WIN_H
#define ZEN_B200_WIN_H

#include <cstddef>
#include <vector>

#include <thrust/device_vector.h>

#include <libzen/zen.h>

namespace zen {
namespace internal {
	namespace win {

		static constexpr float PI = 3.14159265359F;

		enum WindowType { SqrtVonHann, VonHann };

		template <typename T>
		class Window {
		public:
			T window;

			Window(WindowType window_type, std::size_t window_size)
			    : window(window_size, 0.0F)
			{
				std::vector<float> host(window_size);
				zen::b200_detail::check(
				    zen_window(window_type == SqrtVonHann ? ZEN_WIN_SQRT_VON_HANN : ZEN_WIN_VON_HANN, (int)window_size, host.data()),
				    "Window");
				window = T(host.begin(), host.end());
			}
		};

		using WindowGPU = Window<thrust::device_vector<float>>;
		using WindowCPU = Window<std::vector<float>>;

	}  // namespace win
}  // namespace internal
}  // namespace zen

#endif
