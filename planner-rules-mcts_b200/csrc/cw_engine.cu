// Continuous-world host API (placeholder until cw_kernels.cu lands).
#include "common.cuh"
extern "C" void brm_cw_release(brm_engine* e) { (void)e; }
