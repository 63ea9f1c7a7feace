// cb200_launch.h -- one spelling for a kernel launch, so that the host emulation build (tests/simt: the translation unit compiled
// as C++ with CB200_SIMT_EMULATION, CTA threads played by std::threads) runs the SAME launchers -- argument checks, shared-memory
// sizing, variant selection, grid sizing -- on the CPU.  Under nvcc the macro is exactly the triple-chevron launch.
#pragma once
#ifdef CB200_SIMT_EMULATION
#define CB200_LAUNCH(kernel, grid, block, smem_bytes, stream, ...) simt::launch(kernel, (int)(grid), (int)(block), __VA_ARGS__)
#else
#define CB200_LAUNCH(kernel, grid, block, smem_bytes, stream, ...) kernel<<<(grid), (block), (smem_bytes), (stream)>>>(__VA_ARGS__)
#endif
