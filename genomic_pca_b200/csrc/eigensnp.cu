// eigensnp.cu -- EigenSNP driver (placeholder; implemented after the rfit slice is verified on the GPU).
#include "kernels.cuh"
extern "C" int gpca_eigensnp(gpca_ctx* c, const gpca_eigensnp_cfg*, const uint64_t*, uint64_t, const uint64_t*, float*,
                             double*, float*, uint32_t*) {
  if (!c) return GPCA_ERR_INVALID;
  c->set_error("gpca_eigensnp: not implemented yet");
  return GPCA_ERR_INVALID;
}
