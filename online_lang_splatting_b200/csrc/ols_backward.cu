// placeholder until the backward kernels land
#include "ols_common.cuh"
int ols_launch_backward(const ols_raster_args*, const ols_bwd_args*, const ols::WsLayout&, cudaStream_t) {
    ols_set_error("backward not built yet");
    return OLS_ERR_UNSUPPORTED;
}
