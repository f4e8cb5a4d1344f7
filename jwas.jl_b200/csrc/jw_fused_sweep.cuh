// jw_fused_sweep.cuh -- persistent fused sweep ("engine 1"); placeholder until the
// persistent kernel lands: engine 0 (multi-kernel) is the only engine.
#pragma once
#include "jw_common.cuh"
#include "jw_sweep_kernels.cuh"
static int jw_fused_prepare(jwas_handle*) { return 0; }
static void jw_fused_free(jwas_handle*) {}
static int jw_fused_sweep(jwas_handle*, const jw_chain_args&, float) {
    jw_set_error("engine 1 (persistent fused sweep) is not built into this library");
    return 2;
}
