// zen_b200 drop-in for libzen's public <libzen/zen.h> (reference: libzen/libzen/zen.h:8-16).
// Same names so callers recompile unchanged; everything forwards to the C ABI in
// include/zen_b200.h.
#ifndef ZEN_B200_PUB_ZEN_H
#define ZEN_B200_PUB_ZEN_H

#include <limits>
#include <stdexcept>
#include <string>

#include "../../../include/zen_b200.h"

namespace zen {

enum Backend { GPU, CPU };

class ZgException : public std::runtime_error {
public:
	explicit ZgException(std::string msg)
	    : std::runtime_error(msg)
	{
	}
};

constexpr float Eps = std::numeric_limits<float>::epsilon();

namespace b200_detail {
	// C-ABI status -> the reference's error behaviour: geometry errors throw
	// ZgException, resource failures are fatal (the reference prints and exits).
	inline void check(int rc, const char* what)
	{
		if (rc == ZEN_OK)
			return;
		if (rc == ZEN_ERR_GEOMETRY)
			throw ZgException(what);
		throw std::runtime_error(std::string("zen_b200: ") + what + " failed (code " + std::to_string(rc)
		                         + "); there is no CPU fallback");
	}
}  // namespace b200_detail

}  // namespace zen

#endif
