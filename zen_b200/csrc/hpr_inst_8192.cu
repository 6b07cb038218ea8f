// explicit instantiation of the fused HPR kernels for nfft = 8192
#define ZEN_HPR_INSTANTIATE 8192
#include "hpr_launch.cuh"
