// cb200_launch.h -- one spelling for a kernel launch, so that the host emulation build (tests/simt: the translation unit compiled
// as C++ with CB200_SIMT_EMULATION, CTA threads played by std::threads) runs the SAME launchers -- argument checks, shared-memory
// sizing, variant selection, grid sizing -- on the CPU.  Under nvcc the macro is exactly the triple-chevron launch.
#pragma once
// CB200_EXTERN_SHARED declares a kernel's dynamic shared-memory array.  Under nvcc it is exactly `extern __shared__`; in the host
// emulation `__shared__` alone means "static" (a CTA-wide array inside a kernel), so the extern declaration needs its own spelling.
#ifdef CB200_SIMT_EMULATION
#define CB200_EXTERN_SHARED extern
#else
#define CB200_EXTERN_SHARED extern __shared__
#endif
#ifdef CB200_SIMT_EMULATION
#define CB200_LAUNCH(kernel, grid, block, smem_bytes, stream, ...) simt::launch(kernel, (int)(grid), (int)(block), __VA_ARGS__)
#else
#define CB200_LAUNCH(kernel, grid, block, smem_bytes, stream, ...) kernel<<<(grid), (block), (smem_bytes), (stream)>>>(__VA_ARGS__)
#endif

// A barrier among a SUBSET of a CTA's warps (PTX named barrier): the warps of a team that shares one row meet without the other
// teams of the CTA.  `nthreads` must be a multiple of 32; ids 1..15 (0 is __syncthreads).
#ifdef CB200_SIMT_EMULATION
#define CB200_NAMED_BARRIER(id, nthreads) simt::named_barrier((id), (nthreads))
#else
#define CB200_NAMED_BARRIER(id, nthreads) asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory")
#endif

// Every launching entry point runs on the device that OWNS its output buffer, whatever the caller's current device is:
// occupancy queries, cudaFuncSetAttribute(MaxDynamicSharedMemorySize) and the launch itself are per device, and the host
// layer passes raw pointers + the tensor's stream without entering a device context.  CB200_DEVICE_GUARD(ptr) looks the
// pointer's device up (cudaPointerGetAttributes: no synchronisation, legal during graph capture), switches to it when it
// differs from the current one and switches back on scope exit.
#ifdef CB200_SIMT_EMULATION
#define CB200_DEVICE_GUARD(ptr) (void)(ptr)
#else
namespace cb200 {
struct DeviceGuard {
  int prev = -1;
  bool switched = false;
  explicit DeviceGuard(const void *p) {
    if (p == nullptr) return;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
      (void)cudaGetLastError();
      return;
    }
    if (at.type != cudaMemoryTypeDevice && at.type != cudaMemoryTypeManaged) return;
    if (cudaGetDevice(&prev) != cudaSuccess) return;
    if (prev != at.device && cudaSetDevice(at.device) == cudaSuccess) switched = true;
  }
  ~DeviceGuard() {
    if (switched) (void)cudaSetDevice(prev);
  }
  DeviceGuard(const DeviceGuard &) = delete;
  DeviceGuard &operator=(const DeviceGuard &) = delete;
};
}  // namespace cb200
#define CB200_DEVICE_GUARD(ptr) ::cb200::DeviceGuard cb200_device_guard_(ptr)
#endif
