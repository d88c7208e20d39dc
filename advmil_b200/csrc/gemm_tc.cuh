// tcgen05 / TMEM / TMA engine for the N-row contractions (declarations; implementation in gemm_tc.cu).
#pragma once
#include "stages.cuh"

namespace advmil {

bool tc_linear_supported(int rows, int K, int N);
int tc_linear_fwd(const float* x, const float* W, const float* b, int rows, int K, int N, int relu, const Drop& drop,
                  float* y, int precision, cudaStream_t st);
bool tc_gate_supported(int rows, int L, int D);
int tc_gated_score_fwd(const float* v, const float* Wp, const float* bp, const float* wc, const float* bc, int rows,
                       int L, int D, const Drop& da, const Drop& db, float* ab, float* s, float* part, int precision,
                       cudaStream_t st);
bool tc_embed_supported(int rows, int C, int d);
int tc_region_embed_fwd(const float* x, const float* Wc, const float* bc, const float* gamma, const float* beta,
                        int rows, int C, int d, float eps, float* y_pre, float* emb, int precision, cudaStream_t st);
bool tc_bwd_data_supported(int rows, int Ny, int Nx);
int tc_bwd_data(const float* dY, const float* W, int rows, int Ny, int Nx, float* dX, const BwdDataExtras& ex,
                int precision, cudaStream_t st);
bool tc_bwd_weight_supported(int rows, int N1, int N2);
size_t tc_bwd_weight_ws_floats(int rows, int N1, int N2);
int tc_bwd_weight(const float* dY, const float* X, int rows, int N1, int N2, float* dW, int accumulate, float* ws,
                  int precision, cudaStream_t st);

}  // namespace advmil
