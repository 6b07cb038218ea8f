// explicit instantiation of the fused HPR kernels for nfft = 1024
#define ZEN_HPR_INSTANTIATE 1024
#include "hpr_launch.cuh"
