// m1x_env.h -- environment knobs of the PROFILING build only (nvcc -DM1_EXPERIMENTS, tools/build_experiments.sh
// -> build_variants/libm1cu_exp.so).  The product library (build.sh / make cuda) never includes this file:
// it reads no environment variable.
//   M1_DEBUG_SKIP  bit 0: skip the colour phase, bit 1: return after it (phase-split timing; garbage output)
//   M1_PAD_SMEM    extra dynamic shared memory per CTA in bytes (occupancy sweeps)
#pragma once
#include <stdlib.h>
static inline int m1x_env_int(const char *name) { const char *v = getenv(name); return v ? atoi(v) : 0; }
