// tcgen05 / TMEM path of the grouped MLPs (placeholder until the kernels land: reports "unsupported").
#include "common.cuh"

int gnan_mlp_tc_supported(const gnan_mlp_params *, int) { return 0; }
size_t gnan_mlp_tc_workspace_bytes(int64_t, const gnan_mlp_params *, int, int) { return 0; }
int gnan_mlp_tc_fwd(const float *, int64_t, int64_t, const gnan_mlp_params *, float, uint64_t, int, float *, void *, size_t,
                    cudaStream_t)
{
    gnan_set_error("tcgen05 mlp path not built");
    return GNAN_ERR_UNSUPPORTED;
}
int gnan_mlp_tc_bwd(const float *, int64_t, int64_t, const gnan_mlp_params *, float, uint64_t, int, const float *,
                    const gnan_mlp_grads *, void *, size_t, cudaStream_t)
{
    gnan_set_error("tcgen05 mlp path not built");
    return GNAN_ERR_UNSUPPORTED;
}
