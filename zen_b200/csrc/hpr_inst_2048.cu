// explicit instantiation of the fused HPR kernels for nfft = 2048
#define ZEN_HPR_INSTANTIATE 2048
#include "hpr_launch.cuh"
