#!/bin/bash
# Race detection over every kernel the emulated GPU suite reaches: the product's translation units compiled for the host SIMT
# emulation WITH ThreadSanitizer, loaded into pytest through LD_PRELOAD=libtsan.so.  CTA threads and warp lanes are real threads,
# so a shared-memory exchange that lacks its __syncthreads() / __syncwarp() is a reported data race (the process then exits 66).
# ~15-25 min on 8 cores.
TSAN=$(gcc -print-file-name=libtsan.so)
ROOT=$(cd "$(dirname "$0")/.." && pwd)
# BLAS / OpenMP worker pools of numpy and torch are not instrumented: keep them single-threaded and suppress what is left of them
OPENBLAS_NUM_THREADS=1 OMP_NUM_THREADS=1 MKL_NUM_THREADS=1 LD_PRELOAD=$TSAN CB200_SIMT_TSAN=1 \
TSAN_OPTIONS="halt_on_error=1 exitcode=66 report_signal_unsafe=0 suppressions=$ROOT/tests/simt/tsan_suppressions.txt" \
  python -m pytest tests/test_emulated_gpu_suite_cpu.py -q -p no:cacheprovider "$@"
