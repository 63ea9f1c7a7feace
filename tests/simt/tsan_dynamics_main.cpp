// TEST INFRASTRUCTURE: race detector run of the RNEA CTA kernels.  The kernels (compiled as C++ through tests/simt/cuda_runtime.h,
// CTA threads = std::threads, __syncthreads() = std::barrier) are built with -fsanitize=thread and executed on a case read from a
// flat file written by tests/test_simt_emulation_cpu.py; ThreadSanitizer reports any pair of conflicting shared-memory / global
// accesses that no barrier orders -- i.e. a missing __syncthreads() in the kernel source.
#include <cstdio>

#include "simt_dynamics.cpp"

static std::vector<char> blob;
template <class T>
static const T *take(size_t &off, size_t n) {
  const T *p = reinterpret_cast<const T *>(blob.data() + off);
  off += ((n * sizeof(T) + 15) / 16) * 16;
  return p;
}

int main(int argc, char **argv) {
  if (argc < 2) return 2;
  FILE *f = fopen(argv[1], "rb");
  if (!f) return 2;
  fseek(f, 0, SEEK_END);
  blob.resize((size_t)ftell(f));
  fseek(f, 0, SEEK_SET);
  if (fread(blob.data(), 1, blob.size(), f) != blob.size()) return 2;
  fclose(f);
  size_t off = 0;
  const int *hdr = take<int>(off, 4);
  const int B = hdr[0], nl = hdr[1], D = hdr[2], nlev = hdr[3];
  const float *q = take<float>(off, (size_t)B * D), *qd = take<float>(off, (size_t)B * D), *qdd = take<float>(off, (size_t)B * D);
  const float *gt = take<float>(off, (size_t)B * D);
  const float *fixed = take<float>(off, (size_t)nl * 12), *mc = take<float>(off, (size_t)nl * 4), *inn = take<float>(off, (size_t)nl * 8);
  const int8_t *jt = take<int8_t>(off, nl);
  const int16_t *jm = take<int16_t>(off, nl), *lm = take<int16_t>(off, nl);
  const float *joff = take<float>(off, (size_t)nl * 2), *grav = take<float>(off, 6);
  const int16_t *ls = take<int16_t>(off, nlev + 1), *ll = take<int16_t>(off, nl);
  std::vector<float> tau((size_t)B * D), cache((size_t)B * nl * 20), gq((size_t)B * D), gqd((size_t)B * D), gqdd((size_t)B * D);
  double sum = 0;
  for (int R : {8, 16, 32}) {
    const int grid = (B + R - 1) / R > 1 ? (B + R - 1) / R - 1 : 1;
    if (em_rnea_forward(R, grid, tau.data(), q, qd, qdd, fixed, mc, inn, jt, jm, lm, joff, grav, ls, ll, cache.data(), B, nl, D, nlev,
                        nullptr))
      return 3;
    if (em_rnea_backward(R, grid, gq.data(), gqd.data(), gqdd.data(), gt, q, qd, fixed, mc, inn, jt, jm, lm, joff, grav, ls, ll,
                         cache.data(), B, nl, D, nlev, nullptr))
      return 3;
    for (float v : tau) sum += v;
    for (float v : gq) sum += v;
  }
  printf("ok %g\n", sum);
  return 0;
}
