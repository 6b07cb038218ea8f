// Shared definitions for the sm_100a HPR kernels.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <cstdio>
#include <cstdint>

#include "../../include/zen_b200.h"

#define ZEN_EPS 1.1920928955078125e-07f  // FLT_EPSILON, libzen/hps.h:22

#define ZEN_CUDA_CHECK(expr)                                                              \
	do {                                                                                  \
		cudaError_t e__ = (expr);                                                         \
		if (e__ != cudaSuccess) {                                                         \
			std::fprintf(stderr, "zen_b200: CUDA error %s at %s:%d: %s\n",                \
			             cudaGetErrorName(e__), __FILE__, __LINE__, cudaGetErrorString(e__)); \
			return ZEN_ERR_CUDA;                                                          \
		}                                                                                 \
	} while (0)

namespace zen_b200 {

static inline bool is_pow2(long v) { return v > 0 && (v & (v - 1)) == 0; }

// mfilt.h:89-91
static inline int odd_len(int filter_len) { return filter_len + (1 - (filter_len % 2)); }

}  // namespace zen_b200
