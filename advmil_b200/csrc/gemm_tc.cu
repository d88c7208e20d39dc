// tcgen05 engine — placeholder until the TMA/TMEM kernels land: reports "unsupported" so every shape
// runs on the fp32 FFMA engine.  (Replaced in the next milestone.)
#include "gemm_tc.cuh"

namespace advmil {

bool tc_linear_supported(int, int, int) { return false; }
int tc_linear_fwd(const float*, const float*, const float*, int, int, int, int, const Drop&, float*, int, cudaStream_t) {
  set_error("tc_linear_fwd: not built"); return ADVMIL_ERR_INVALID;
}
bool tc_gate_supported(int, int, int) { return false; }
int tc_gated_score_fwd(const float*, const float*, const float*, const float*, const float*, int, int, int, const Drop&,
                       const Drop&, float*, float*, int, cudaStream_t) {
  set_error("tc_gated_score_fwd: not built"); return ADVMIL_ERR_INVALID;
}
bool tc_embed_supported(int, int, int) { return false; }
int tc_region_embed_fwd(const float*, const float*, const float*, const float*, const float*, int, int, int, float,
                        float*, float*, int, cudaStream_t) {
  set_error("tc_region_embed_fwd: not built"); return ADVMIL_ERR_INVALID;
}
bool tc_bwd_data_supported(int, int, int) { return false; }
int tc_bwd_data(const float*, const float*, int, int, int, float*, const BwdDataExtras&, int, cudaStream_t) {
  set_error("tc_bwd_data: not built"); return ADVMIL_ERR_INVALID;
}
bool tc_bwd_weight_supported(int, int, int) { return false; }
int tc_bwd_weight(const float*, const float*, int, int, int, float*, int, float*, int, cudaStream_t) {
  set_error("tc_bwd_weight: not built"); return ADVMIL_ERR_INVALID;
}

}  // namespace advmil
