// gemm_tn_tc.cuh — tcgen05 route of gemm_tn (see gemm_tn_tc.cu).
#pragma once

#include "dense.cuh"

namespace b200q {

bool gemm_tn_tc_supported(const GemmTN& g);
int gemm_tn_tc(const GemmTN& g, cudaStream_t st);

}  // namespace b200q
