// gemm_tc.cu -- tcgen05 tensor-core engines (placeholder until the TMA/TMEM kernels land in this file).
#include "common.cuh"

namespace cofi {
struct Epilogue;
bool gemm_tc_supported(int64_t, int64_t, int64_t, int64_t, int, int, const void*, const void*, const void*) {
    return false;
}
int gemm_tc_launch(const float*, int64_t, const float*, int64_t, float*, int64_t, int64_t, int, int, const Epilogue&,
                   int, cudaStream_t) {
    set_error("tensor-core GEMM engine not built");
    return COFI_EUNSUPPORTED;
}
bool attention_tc_supported(int64_t, int64_t, int, int) { return false; }
int attention_tc_launch(const float*, const float*, const float*, int64_t, int64_t, int, int, int, float, float*,
                        cudaStream_t) {
    set_error("tensor-core attention engine not built");
    return COFI_EUNSUPPORTED;
}
bool sim_argmin_tc_supported(int64_t, int64_t, int64_t, int64_t, int) { return false; }
int sim_argmin_tc_launch(const float*, int64_t, const float*, int64_t, int64_t, int64_t, int, int, int64_t*, float*,
                         int, cudaStream_t) {
    set_error("tensor-core similarity engine not built");
    return COFI_EUNSUPPORTED;
}
}  // namespace cofi
