// on-chip sampler kernels for family/link code 3 (see hmc_fused.cuh)
#include "hmc_fused.cuh"
int gmb_fused_launch_fl3(gmb_ctx* ctx, const FusedParams& p, size_t smem, int cs) { return launch_fused_ks<3>(ctx, p, smem, cs); }
