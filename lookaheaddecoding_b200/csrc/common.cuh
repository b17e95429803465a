// Shared helpers for the lade sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include "../../include/lade_sm100.h"

namespace lade {

void set_cuda_error(cudaError_t e, const char* where);

#define LADE_CUDA_CHECK(expr)                                   \
  do {                                                          \
    cudaError_t _e = (expr);                                    \
    if (_e != cudaSuccess) {                                    \
      ::lade::set_cuda_error(_e, #expr);                        \
      return LADE_ECUDA;                                        \
    }                                                           \
  } while (0)

#define LADE_LAUNCH_CHECK(name)                                 \
  do {                                                          \
    cudaError_t _e = cudaGetLastError();                        \
    if (_e != cudaSuccess) {                                    \
      ::lade::set_cuda_error(_e, name);                         \
      return LADE_ECUDA;                                        \
    }                                                           \
  } while (0)

__host__ __device__ inline int rowdesc_make(int cls, int blk, int idx) {
  return (int)(((unsigned)cls << 30) | ((unsigned)blk << 15) | (unsigned)idx);
}
__host__ __device__ inline int rowdesc_cls(int rd) { return (int)((unsigned)rd >> 30); }
__host__ __device__ inline int rowdesc_blk(int rd) { return (int)(((unsigned)rd >> 15) & 0x7fffu); }
__host__ __device__ inline int rowdesc_idx(int rd) { return (int)((unsigned)rd & 0x7fffu); }

// Lookahead mask predicate over step-local (row r, column c); SURVEY.md App. B ==
// j_make_causal_mask_multilevel (lade/models/modeling_llama.py:115-207).  Cache columns are
// visible to every row and never reach this function.
__device__ __forceinline__ bool row_sees(int rd_r, int r, int rd_c, int c, int level_offset) {
  const int tr = rowdesc_cls(rd_r);
  const int tc = rowdesc_cls(rd_c);
  if (tr == LADE_ROW_PREFIX) return tc == LADE_ROW_PREFIX && c <= r;        // :124-130, :189-192
  if (tr == LADE_ROW_WINDOW) {
    if (tc == LADE_ROW_PREFIX) return true;                                 // :195
    if (tc != LADE_ROW_WINDOW) return false;
    const int bc = rowdesc_blk(rd_c), ic = rowdesc_idx(rd_c);
    const int br = rowdesc_blk(rd_r), ir = rowdesc_idx(rd_r);
    return bc == 0 ? (ic <= ir) : (bc <= br && ic == ir);                   // :201-203
  }
  if (tr == LADE_ROW_GUESS) {
    if (c <= level_offset) return true;                                     // :184
    return tc == LADE_ROW_GUESS && rowdesc_blk(rd_c) == rowdesc_blk(rd_r) &&
           rowdesc_idx(rd_c) <= rowdesc_idx(rd_r);                          // :141-181
  }
  return c == r;  // PAD rows only see themselves
}

__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

}  // namespace lade
