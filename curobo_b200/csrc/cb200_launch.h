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
