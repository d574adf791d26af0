// one translation unit per kernel instantiation (parallel build)
#include "spg_inst.cuh"
spg_status spg_launch_3_256l(spg_ctx *ctx, spg::KernelParams &kp) { return launch_bucket<3, 256, false, true>(ctx, kp); }
