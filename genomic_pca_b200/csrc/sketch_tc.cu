// sketch_tc.cu -- tcgen05 / TMEM / TMA engine for the sketch pass (placeholder until the engine lands).
#include "sketch_tc.cuh"

bool sketch_tc_supported(gpca_ctx*, const SketchProblem&) { return false; }
int launch_sketch_tc(gpca_ctx* c, const SketchProblem&) {
  c->set_error("tcgen05 engine not built");
  return GPCA_ERR_INVALID;
}
