// Host-side helpers shared by the translation units of libdecomp_b200.so.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/decomp_b200.h"

namespace dcp {

void set_error(const char* fmt, ...);
int check_cuda(cudaError_t e, const char* what);
int num_sms();

// Encodes a 2-D row-major FP64 tensor map with 128-byte swizzle.
//   inner: extent of the contiguous dimension (elements), outer: rows, ld: row pitch (elements)
int make_tensor_map(CUtensorMap* map, const double* base, uint64_t inner, uint64_t outer, uint64_t ld,
                    uint32_t box_inner, uint32_t box_outer);

// TN operand [K rows][W columns] (W % 16 == 0) seen as the 3-D tensor {16, K, W/16}: one box {16, 16, blocks} lands in
// shared memory as [block][k][16 doubles] with the 128-byte swizzle -- the TN tile layout -- in ONE instruction.
int make_tensor_map_tn3d(CUtensorMap* map, const double* base, uint64_t width, uint64_t rows, uint64_t ld,
                         uint32_t blocks);

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

}  // namespace dcp

#define DCP_CHECK_LAUNCH(what)                                                   \
  do {                                                                           \
    int rc_ = dcp::check_cuda(cudaGetLastError(), what);                         \
    if (rc_ != DECOMP_OK) return rc_;                                            \
  } while (0)
