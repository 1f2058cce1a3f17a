// tcgen05 / TMEM / TMA GEMM engine (placeholder until the tensor-core engine lands: every query answers
// "shape not taken", so the SIMT fp32 engine serves all problems).
#include "sgemm.cuh"

namespace cvar {
int tc_gemm_try(const cvar_gemm_args*, cudaStream_t) { return 0; }
int tc_qkv_try(const float*, const float*, const QkvEpilogue&, int, int, cudaStream_t) { return 0; }
int tc_conv_try(const cvar_conv_args*, cudaStream_t) { return 0; }
}  // namespace cvar
