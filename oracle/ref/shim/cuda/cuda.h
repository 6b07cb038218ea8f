/* Include-path shim: the reference's hps.cu includes <cuda/cuda.h> (a Fedora
 * packaging path, /root/reference/libzen/hps.cu:1). Test infrastructure only. */
#include <cuda.h>
