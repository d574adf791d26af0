// one translation unit per kernel instantiation (parallel build)
#include "spg_fast_inst.cuh"
spg_status spg_launch_fast_6_g8(spg_ctx *ctx, spg::KernelParams &kp) { return launch_fast<6, 8, 8>(ctx, kp); }
