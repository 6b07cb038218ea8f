// explicit instantiation of the fused HPR kernels for nfft = 512
#define ZEN_HPR_INSTANTIATE 512
#include "hpr_launch.cuh"
