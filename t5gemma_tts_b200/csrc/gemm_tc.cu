// tcgen05 / TMEM / TMA GEMM (placeholder until the kernel lands; never silently falls back).
#include "kernels.h"
cudaError_t launch_gemm_tc(const GemmArgs&, cudaStream_t) { return cudaErrorNotSupported; }
