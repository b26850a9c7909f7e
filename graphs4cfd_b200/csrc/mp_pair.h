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

}  // namespace g4c
