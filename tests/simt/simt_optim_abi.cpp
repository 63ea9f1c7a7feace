// TEST INFRASTRUCTURE: cb200_optim.cu (L-BFGS step and line-search kernels + their C ABI) as a translation unit of the emulated
// product library (tests/simt/cuda_runtime.h).
#define CB200_SIMT_EMULATION 1
#include "cuda_runtime.h"

#include "../../curobo_b200/csrc/cb200_optim.cu"
