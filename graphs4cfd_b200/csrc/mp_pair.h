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

// mp_edge_v5.cu: the kernel behind g4c_edge_aggr_fwd for launches with a fixed in-degree and no permutations (TMA data paths, packed fp32 epilogues)
bool edge_v5_supported(const EdgeArgs& a);
int edge_v5_launch(const EdgeArgs& a, cudaStream_t st);
int edge_v5_profile(unsigned long long* out64);              // phase profile (needs -DG4C_PROFILE)
int tma_test_launch(int test, const float* src, int64_t rows, int k, float* out, int c0, int j, int n0, cudaStream_t st);   // tma_test.cu

}  // namespace g4c
