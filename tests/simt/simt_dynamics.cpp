// TEST INFRASTRUCTURE: the RNEA CTA kernels of curobo_b200/csrc/cb200_dynamics.cu compiled as ordinary C++ and executed by
// std::threads (tests/simt/cuda_runtime.h), so every rows-per-CTA instantiation can be value-checked without a GPU.
#define CB200_SIMT_EMULATION 1
#include "cuda_runtime.h"

#include "../../curobo_b200/csrc/cb200_dynamics.cu"

namespace {
Model make_model(const float *fixed, const float *mc, const float *inertia, const int8_t *jtype, const int16_t *jmap,
                 const int16_t *lmap, const float *joff, const float *gravity, const int16_t *lstarts, const int16_t *llinks,
                 int nl, int D, int n_levels) {
  return Model{fixed, mc, inertia, jtype, jmap, lmap, joff, gravity, lstarts, llinks, nl, D, n_levels};
}
}  // namespace

extern "C" {
int em_rnea_forward(int R, int grid, float *tau, const float *q, const float *qd, const float *qdd, const float *fixed,
                    const float *mc, const float *inertia, const int8_t *jtype, const int16_t *jmap, const int16_t *lmap,
                    const float *joff, const float *gravity, const int16_t *lstarts, const int16_t *llinks, float *cache, int B,
                    int nl, int D, int n_levels, const float *f_ext) {
  FwdArgs a{make_model(fixed, mc, inertia, jtype, jmap, lmap, joff, gravity, lstarts, llinks, nl, D, n_levels), tau, cache, q, qd,
            qdd, f_ext, B};
  if (R == 32) simt::launch(rnea_forward_cta<32>, grid, kThreads, a);
  else if (R == 16) simt::launch(rnea_forward_cta<16>, grid, kThreads, a);
  else if (R == 8) simt::launch(rnea_forward_cta<8>, grid, kThreads, a);
  else return 1;
  return 0;
}

int em_rnea_backward(int R, int grid, float *gq, float *gqd, float *gqdd, const float *grad_tau, const float *q, const float *qd,
                     const float *fixed, const float *mc, const float *inertia, const int8_t *jtype, const int16_t *jmap,
                     const int16_t *lmap, const float *joff, const float *gravity, const int16_t *lstarts, const int16_t *llinks,
                     const float *cache, int B, int nl, int D, int n_levels, float *grad_f_ext) {
  BwdArgs a{make_model(fixed, mc, inertia, jtype, jmap, lmap, joff, gravity, lstarts, llinks, nl, D, n_levels), gq, gqd, gqdd,
            grad_f_ext, grad_tau, q, qd, cache, B};
  if (R == 32) simt::launch(rnea_backward_cta<32>, grid, kThreads, a);
  else if (R == 16) simt::launch(rnea_backward_cta<16>, grid, kThreads, a);
  else if (R == 8) simt::launch(rnea_backward_cta<8>, grid, kThreads, a);
  else return 1;
  return 0;
}
}
