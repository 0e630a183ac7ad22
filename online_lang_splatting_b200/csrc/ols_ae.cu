// placeholder until the autoencoder kernels land
#include "ols_common.cuh"
extern "C" {
int ols_ae_plan_create(const ols_ae_chain*, ols_ae_plan**, void*) { ols_set_error("AE not built yet"); return OLS_ERR_UNSUPPORTED; }
void ols_ae_plan_destroy(ols_ae_plan*) {}
int ols_ae_forward(const ols_ae_plan*, const float*, float*, int64_t, void*) { ols_set_error("AE not built yet"); return OLS_ERR_UNSUPPORTED; }
}
