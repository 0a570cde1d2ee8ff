// sketch_tc.cuh -- tcgen05 / TMEM / TMA engine for the sketch pass (engine 1).
#pragma once
#include "kernels.cuh"

// true when the tcgen05 engine can take this problem (shape/alignment); otherwise the SIMT engine runs.
bool sketch_tc_supported(gpca_ctx* c, const SketchProblem& p);
int launch_sketch_tc(gpca_ctx* c, const SketchProblem& p);

// integer engine (engine 2): exact int32 accumulation of u8 codes x two s8 limbs of the dense operand, l <= 32
bool sketch_i8_supported(gpca_ctx* c, const SketchProblem& p);
int launch_sketch_i8(gpca_ctx* c, const SketchProblem& p);
