// explicit instantiation of the fused HPR kernels for nfft = 4096
#define ZEN_HPR_INSTANTIATE 4096
#include "hpr_launch.cuh"
