// Device-resident decode state shared by the state-machine kernels (state.cu) and the sampling kernel (sampling.cu).
#pragma once
#include "common.cuh"

namespace lade {

// header ints of the device state block
enum {
  S_FILL_LEVEL = 0, S_LST_TOKEN, S_KV_LEN, S_N_OUT, S_N_OLD, S_DONE, S_STEPS, S_MAX_LENGTH,
  S_N_PROMPT, S_N_GUESS_TOK, S_SKIP, S_HDR_INTS = 16
};

struct Dims {
  int W, N, G, GS, WCAP, V, cap, pool_from_prompt, n_eos;
  int D, rank;   // lookahead parallelism: DIST_WORKERS, LOCAL_RANK (1, 0 when off)
  int eos[4];
  int lm_cap;
  // offsets (in ints) into the state block
  int off_win, off_win_len, off_guess, off_out, off_old, off_cnt;
  long long off_tup;
  long long total_ints;
};

}  // namespace lade

struct LadeCtx {
  LadeConfig cfg;
  lade::Dims d;
  int32_t* state;   // device
};

namespace lade {

__device__ __forceinline__ int* st_win(int* st, const Dims& d, int level) { return st + d.off_win + level * d.WCAP; }
__device__ __forceinline__ int lp_rec_ints(const Dims& d) { return 3 + d.GS + d.WCAP; }

}  // namespace lade
