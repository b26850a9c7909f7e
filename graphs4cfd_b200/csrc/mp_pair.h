// mp_pair.h — internal argument blocks of the CTA-pair tensor-core kernels (hidden = 128, fp16x3).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/g4c.h"

namespace g4c {

using EdgeArgs = G4cEdgeDesc;   // mp_edge_pair.cu

int edge_pair_launch(const EdgeArgs& a, cudaStream_t st);
int edge_pair_profile(unsigned long long* out64);            // phase profile (needs -DG4C_PROFILE)
int row_pair_launch(const G4cRowTcDesc& d, cudaStream_t st);   // mp_row_pair.cu

// mp_edge_pair_tma.cu: experimental bulk-tensor (TMA) variants of the edge kernel, selected by G4C_EDGE_MODE / g4c_debug_set_edge_mode
// (0 = the v3 kernel, default; 1 = e' through shared memory + TMA stores; 2 = 1 + e / P_c tiles through TMA loads; 3 = 2 + P_r[src] through TMA gather4; 4 = 3 with 96 / 48 registers per epilogue / loader thread)
int edge_pair_mode();
void edge_pair_set_mode(int mode);
bool edge_pair_tma_supported(const EdgeArgs& a);
int edge_pair_tma_launch(const EdgeArgs& a, int mode, cudaStream_t st);
int edge_pair_tma_profile(unsigned long long* out64);
int tma_test_launch(int test, const float* src, int64_t rows, int k, const int32_t* idx, float* out, int c0, int j, int n0, cudaStream_t st);   // tma_test.cu         // phase profile of the TMA variants (needs -DG4C_PROFILE)

}  // namespace g4c
