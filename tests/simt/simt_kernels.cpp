// TEST INFRASTRUCTURE: the main translation unit of the product (curobo_b200/csrc/cb200_kernels.cu: the fused rollout kernels, the
// per-operator drop-in kernels, the blob packer and the whole C ABI with its launch logic) compiled as ordinary C++ and executed by
// std::threads (tests/simt/cuda_runtime.h).  The exported symbols are the product's own cb200_* entry points: they take HOST
// pointers here.
#define CB200_SIMT_EMULATION 1
#define CB200_SIMT_SMEM_DECL alignas(128) unsigned char smem[232448]; alignas(16) float fsm[58112];
#include "cuda_runtime.h"

#include "../../curobo_b200/csrc/cb200_kernels.cu"
