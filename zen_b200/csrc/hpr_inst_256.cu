// explicit instantiation of the fused HPR kernels for nfft = 256
#define ZEN_HPR_INSTANTIATE 256
#include "hpr_launch.cuh"
