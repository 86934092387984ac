#pragma once
#include "epilogue.cuh"
namespace dbn {
inline bool umma_gemm_supported(int, int, const dbn_operands*, int) { return false; }
inline int umma_gemm_act(cudaStream_t, int, int, const dbn_operands*, int, const dbn_act_epilogue&) { return DBN_ERR_UNSUPPORTED; }
inline int umma_gemm_update(cudaStream_t, int, int, int, int, const dbn_operands*, int, const dbn_update_epilogue&) { return DBN_ERR_UNSUPPORTED; }
}
