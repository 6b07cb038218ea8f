/* Include-path shim for /root/reference/libzen/hps.cu:2. Test infrastructure only. */
#include <cuda_runtime.h>
