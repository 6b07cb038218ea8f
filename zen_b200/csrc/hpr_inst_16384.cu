// explicit instantiation of the fused HPR kernels for nfft = 16384
#define ZEN_HPR_INSTANTIATE 16384
#include "hpr_launch.cuh"
