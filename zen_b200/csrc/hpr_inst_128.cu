// explicit instantiation of the fused HPR kernels for nfft = 128
#define ZEN_HPR_INSTANTIATE 128
#include "hpr_launch.cuh"
