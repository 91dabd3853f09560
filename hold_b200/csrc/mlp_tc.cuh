// tcgen05 (5th-gen tensor core) MLP path — HOLD_MLP_TC.  (placeholder: filled in below the fp32 path)
#pragma once
#include "common.cuh"
namespace hold {
struct TcMlp { int dummy; };
static int tc_init(hold_ctx*) { return HOLD_OK; }
static void tc_free(NodeState&) {}
static int tc_pack(hold_ctx*, NodeState&, const hold_mlp_weights*, const hold_mlp_weights*, cudaStream_t) { return HOLD_OK; }
static int tc_launch_sdf(hold_ctx*, NodeState&, int, const float*, const float*, float*, float*, float*, const SamplerState*, cudaStream_t) {
  set_error("HOLD_MLP_TC not built"); return HOLD_E_STATE; }
static int tc_launch_rgb(hold_ctx*, NodeState&, int, int, const float*, const float*, const float*, const float*, const float*, float*, cudaStream_t) {
  set_error("HOLD_MLP_TC not built"); return HOLD_E_STATE; }
}
