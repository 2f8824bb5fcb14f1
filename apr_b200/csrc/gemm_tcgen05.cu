// gemm_tcgen05.cu — TF32 GEMM on 5th-gen tensor cores (tcgen05.mma, TMEM accumulators, TMA operand staging).
// Placeholder until the tcgen05 path lands: reports "unsupported" so callers take the fp32 CUDA-core path.
#include "common.cuh"

namespace aprb {

bool gemm_tf32_supported(int M, int N, int K) { (void)M; (void)N; (void)K; return false; }

int gemm_tf32_rowscale(const float*, const float*, int, int, int, const float*, float*, cudaStream_t) {
    set_error("gemm_tf32_rowscale: tcgen05 path not built");
    return APRB_ERR_UNSUPPORTED;
}

}  // namespace aprb

extern "C" int aprb_linear_tf32(const float* d_x, const float* d_W, int N, int Cin, int Cout, float* d_y, void* stream) {
    (void)d_x; (void)d_W; (void)N; (void)Cin; (void)Cout; (void)d_y; (void)stream;
    aprb::set_error("aprb_linear_tf32: tcgen05 path not built");
    return APRB_ERR_UNSUPPORTED;
}
