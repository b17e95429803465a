// Memory-bound glue kernels of the decoder layer, with the reference's bf16 rounding points.
//   RMSNorm (+ fused residual add)   lade/models/modeling_llama.py:222-227, :883-889
//   rotary embedding + KV append     lade/models/modeling_llama.py:342-346, :513-516
//   SwiGLU                            lade/models/modeling_llama.py:378
// All are HBM-bound byte movers: 16-byte vectorised, coalesced, one pass over the data.
#include "common.cuh"

#include <cstdlib>

namespace lade {

// Programmatic dependent launch on the glue kernels: every kernel orders itself behind its stream predecessor with
// griddepcontrol.wait as its first instruction (so it is correct whatever the predecessor does) and then lets ITS
// dependent start launching at once.  Whether a neighbouring library GEMM takes part is up to the library (a kernel
// launched without the attribute, or one that never triggers, simply behaves like a normal stream-ordered launch).
__device__ __forceinline__ void pdl_enter() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

static bool pdl_glue_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("LADE_PDL_GLUE");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v != 0;
}

// kernel launch with cudaLaunchAttributeProgrammaticStreamSerialization (when enabled)
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_glue_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// One CTA per row.  h = bf16(x + delta) (if delta), out = w * bf16(h_f32 * rsqrt(mean(h^2) + eps)).
template <bool GATHER>
__global__ void __launch_bounds__(1024) rmsnorm_kernel(const __nv_bfloat16* __restrict__ x,
                                                      const __nv_bfloat16* __restrict__ delta,
                                                      const __nv_bfloat16* __restrict__ w,
                                                      const int* __restrict__ rows_idx,
                                                      __nv_bfloat16* __restrict__ h_out,
                                                      __nv_bfloat16* __restrict__ out, int hidden, float eps) {
  extern __shared__ float s_row[];  // hidden floats
  __shared__ float s_part[32];
  pdl_enter();
  const int out_row = blockIdx.x;
  const int in_row = GATHER ? rows_idx[out_row] : out_row;
  const __nv_bfloat16* xr = x + (long long)in_row * hidden;
  const __nv_bfloat16* dr = delta ? delta + (long long)in_row * hidden : nullptr;
  const int nvec = hidden / 8;
  float ss = 0.f;
  for (int i = threadIdx.x; i < nvec; i += blockDim.x) {
    uint4 v = reinterpret_cast<const uint4*>(xr)[i];
    __nv_bfloat16* e = reinterpret_cast<__nv_bfloat16*>(&v);
    float f[8];
    if (dr) {
      uint4 dv = reinterpret_cast<const uint4*>(dr)[i];
      const __nv_bfloat16* de = reinterpret_cast<const __nv_bfloat16*>(&dv);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        e[j] = __float2bfloat16_rn(__bfloat162float(e[j]) + __bfloat162float(de[j]));   // residual add in bf16
        f[j] = __bfloat162float(e[j]);
      }
      if (h_out && !GATHER) reinterpret_cast<uint4*>(h_out + (long long)out_row * hidden)[i] = v;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = __bfloat162float(e[j]);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s_row[i * 8 + j] = f[j];
      ss += f[j] * f[j];
    }
  }
  ss = warp_sum(ss);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = ss;
  __syncthreads();
  float tot = 0.f;
  for (int k = 0; k < (int)(blockDim.x >> 5); ++k) tot += s_part[k];
  const float rstd = rsqrtf(tot / (float)hidden + eps);
  __nv_bfloat16* orow = out + (long long)out_row * hidden;
  for (int i = threadIdx.x; i < nvec; i += blockDim.x) {
    uint4 wv = reinterpret_cast<const uint4*>(w)[i];
    const __nv_bfloat16* we = reinterpret_cast<const __nv_bfloat16*>(&wv);
    uint4 ov;
    __nv_bfloat16* oe = reinterpret_cast<__nv_bfloat16*>(&ov);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float n = bf16_round(s_row[i * 8 + j] * rstd);                 // .to(input_dtype)
      oe[j] = __float2bfloat16_rn(__bfloat162float(we[j]) * n);           // weight * hidden (bf16 mul)
    }
    reinterpret_cast<uint4*>(orow)[i] = ov;
  }
}

// RoPE + append.  grid: rows ; block: 256 threads.  Work item = (head, 8-wide slice of the first half):
// the thread rotates elements [8i, 8i+8) of the first half against the same slice of the second half,
// all accesses 16 bytes.  (Hq + 2 Hkv) * D/16 items per row.
__global__ void __launch_bounds__(1024) rope_append_kernel(
    const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ cos_tab,
    const __nv_bfloat16* __restrict__ sin_tab, const int* __restrict__ pos, const int* __restrict__ meta,
    __nv_bfloat16* __restrict__ q_out, __nv_bfloat16* __restrict__ k_cache, __nv_bfloat16* __restrict__ v_cache,
    int q_pad, int n_heads, int n_kv_heads, int D, int kv_capacity, int max_pos) {
  // programmatic dependent launch: let the consumer (lade_attn_fwd) start its prologue and prefetch the cache tiles of
  // earlier steps while this grid runs; it orders itself with griddepcontrol.wait before touching what is written here
  pdl_enter();
  const int r = blockIdx.x;
  const int half = D >> 1;
  const int per_head = half >> 3;                        // 16-byte slices per half head
  const int n_items = (n_heads + 2 * n_kv_heads) * per_head;
  const int ld = (n_heads + 2 * n_kv_heads) * D;
  const int kv_len = meta[LADE_M_KV_LEN];
  const int cache_row = kv_len + r;
  int p = pos[r];
  p = p < 0 ? 0 : (p >= max_pos ? max_pos - 1 : p);
  const __nv_bfloat16* crow = cos_tab + (long long)p * D;
  const __nv_bfloat16* srow = sin_tab + (long long)p * D;
  for (int it = threadIdx.x; it < n_items; it += blockDim.x) {
    const int h = it / per_head, sl = it % per_head;
    const __nv_bfloat16* src = qkv + (long long)r * ld + (long long)h * D + sl * 8;
    const uint4 a = *reinterpret_cast<const uint4*>(src);
    const uint4 b = *reinterpret_cast<const uint4*>(src + half);
    __nv_bfloat16* dst;
    if (h < n_heads) {
      dst = q_out + ((long long)h * q_pad + r) * D + sl * 8;
    } else if (h < n_heads + n_kv_heads) {
      if (cache_row >= kv_capacity) continue;
      dst = k_cache + ((long long)(h - n_heads) * kv_capacity + cache_row) * D + sl * 8;
    } else {                                              // V: plain append
      if (cache_row >= kv_capacity) continue;
      dst = v_cache + ((long long)(h - n_heads - n_kv_heads) * kv_capacity + cache_row) * D + sl * 8;
      *reinterpret_cast<uint4*>(dst) = a;
      *reinterpret_cast<uint4*>(dst + half) = b;
      continue;
    }
    const uint4 c1 = *reinterpret_cast<const uint4*>(crow + sl * 8);
    const uint4 c2 = *reinterpret_cast<const uint4*>(crow + half + sl * 8);
    const uint4 s1 = *reinterpret_cast<const uint4*>(srow + sl * 8);
    const uint4 s2 = *reinterpret_cast<const uint4*>(srow + half + sl * 8);
    const __nv_bfloat16* x1 = reinterpret_cast<const __nv_bfloat16*>(&a);
    const __nv_bfloat16* x2 = reinterpret_cast<const __nv_bfloat16*>(&b);
    const __nv_bfloat16* pc1 = reinterpret_cast<const __nv_bfloat16*>(&c1);
    const __nv_bfloat16* pc2 = reinterpret_cast<const __nv_bfloat16*>(&c2);
    const __nv_bfloat16* ps1 = reinterpret_cast<const __nv_bfloat16*>(&s1);
    const __nv_bfloat16* ps2 = reinterpret_cast<const __nv_bfloat16*>(&s2);
    uint4 o1v, o2v;
    __nv_bfloat16* o1 = reinterpret_cast<__nv_bfloat16*>(&o1v);
    __nv_bfloat16* o2 = reinterpret_cast<__nv_bfloat16*>(&o2v);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float f1 = __bfloat162float(x1[j]), f2 = __bfloat162float(x2[j]);
      // (q * cos) + (rotate_half(q) * sin), every op rounded to bf16 (modeling_llama.py:344-345)
      o1[j] = __float2bfloat16_rn(bf16_round(f1 * __bfloat162float(pc1[j])) + bf16_round(-f2 * __bfloat162float(ps1[j])));
      o2[j] = __float2bfloat16_rn(bf16_round(f2 * __bfloat162float(pc2[j])) + bf16_round(f1 * __bfloat162float(ps2[j])));
    }
    *reinterpret_cast<uint4*>(dst) = o1v;
    *reinterpret_cast<uint4*>(dst + half) = o2v;
  }
}

__global__ void swiglu_kernel(const __nv_bfloat16* __restrict__ gate_up, __nv_bfloat16* __restrict__ out,
                              int rows, int inter) {
  pdl_enter();
  const int nvec = inter / 8;
  const long long total = (long long)rows * nvec;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / nvec), c = (int)(i % nvec);
    const uint4 gv = reinterpret_cast<const uint4*>(gate_up + (long long)r * 2 * inter)[c];
    const uint4 uv = reinterpret_cast<const uint4*>(gate_up + (long long)r * 2 * inter + inter)[c];
    const __nv_bfloat16* g = reinterpret_cast<const __nv_bfloat16*>(&gv);
    const __nv_bfloat16* u = reinterpret_cast<const __nv_bfloat16*>(&uv);
    uint4 ov;
    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(&ov);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float x = __bfloat162float(g[j]);
      const float s = bf16_round(x / (1.0f + expf(-x)));                    // silu in fp32, bf16 result
      o[j] = __float2bfloat16_rn(s * __bfloat162float(u[j]));
    }
    reinterpret_cast<uint4*>(out + (long long)r * inter)[c] = ov;
  }
}

// L2 prefetch of a weight range (fire-and-forget): one lane per CTA issues `cp.async.bulk.prefetch.L2` for chunks of
// the range in a grid-stride loop and exits; the TMA engines pull the lines into the 126 MB L2 while OTHER kernels
// (attention, norms, RoPE -- phases of the step in which HBM is otherwise idle) run.  No data reaches the SM.
__global__ void l2_prefetch_kernel(const char* __restrict__ base, long long bytes, int chunk) {
  // UBLKPF.L2 is a per-warp (uniform datapath) instruction: one lane per one-warp CTA issues, chunks interleaved over CTAs
  if (threadIdx.x != 0) return;
  const long long n_chunks = (bytes + chunk - 1) / chunk;
  for (long long i = blockIdx.x; i < n_chunks; i += gridDim.x) {
    const long long off = i * chunk;
    long long sz = bytes - off;
    if (sz > chunk) sz = chunk;
    sz &= ~15ll;                                           // size operand: multiple of 16 bytes
    if (sz > 0)
      asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(base + off), "r"((unsigned)sz) : "memory");
  }
}

}  // namespace lade

using namespace lade;

extern "C" {

static int norm_threads(int hidden) {
  int t = ((hidden / 8 + 31) / 32) * 32;
  return t < 128 ? 128 : (t > 1024 ? 1024 : t);
}

int lade_l2_prefetch(void* stream, const void* ptr, int64_t bytes, int32_t n_ctas, int32_t chunk_bytes) {
  if (!ptr || bytes < 0 || n_ctas < 1 || chunk_bytes < 16 || (chunk_bytes & 15) || (reinterpret_cast<uintptr_t>(ptr) & 15))
    return LADE_EINVAL;
  if (bytes == 0) return LADE_OK;
  l2_prefetch_kernel<<<n_ctas, 32, 0, (cudaStream_t)stream>>>((const char*)ptr, bytes, chunk_bytes);
  LADE_LAUNCH_CHECK("l2_prefetch_kernel");
  return LADE_OK;
}

int lade_rmsnorm(void* stream, const void* x, const void* delta, const void* weight, void* h_out, void* out,
                 int32_t rows, int32_t hidden, float eps) {
  if (!x || !weight || !out || rows < 1 || hidden < 8 || hidden % 8 != 0) return LADE_EINVAL;
  if (delta && !h_out) return LADE_EINVAL;
  const size_t smem = sizeof(float) * hidden;
  if (smem > 96 * 1024) return LADE_EUNSUPPORTED;
  if (smem > 48 * 1024)
    LADE_CUDA_CHECK(cudaFuncSetAttribute(rmsnorm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // one 16-byte vector per thread when the row fits (4096 / 8 = 512 threads): a single round of loads per phase
  const int threads = norm_threads(hidden);
  LADE_CUDA_CHECK(launch_pdl(rmsnorm_kernel<false>, dim3(rows), dim3(threads), smem, (cudaStream_t)stream,
                             (const __nv_bfloat16*)x, (const __nv_bfloat16*)delta, (const __nv_bfloat16*)weight,
                             (const int*)nullptr, (__nv_bfloat16*)h_out, (__nv_bfloat16*)out, hidden, eps));
  return LADE_OK;
}

int lade_rmsnorm_gather(void* stream, const void* x, const void* delta, const void* weight,
                        const int32_t* rows_idx, void* out, int32_t n_rows, int32_t hidden, float eps) {
  if (!x || !weight || !out || !rows_idx || n_rows < 1 || hidden < 8 || hidden % 8 != 0) return LADE_EINVAL;
  const size_t smem = sizeof(float) * hidden;
  if (smem > 96 * 1024) return LADE_EUNSUPPORTED;
  if (smem > 48 * 1024)
    LADE_CUDA_CHECK(cudaFuncSetAttribute(rmsnorm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int threads = norm_threads(hidden);
  LADE_CUDA_CHECK(launch_pdl(rmsnorm_kernel<true>, dim3(n_rows), dim3(threads), smem, (cudaStream_t)stream,
                             (const __nv_bfloat16*)x, (const __nv_bfloat16*)delta, (const __nv_bfloat16*)weight,
                             (const int*)rows_idx, (__nv_bfloat16*)nullptr, (__nv_bfloat16*)out, hidden, eps));
  return LADE_OK;
}

int lade_rope_append(void* stream, const void* qkv, const void* cos_tab, const void* sin_tab,
                     const int32_t* pos, const int32_t* meta, void* q_out, void* k_cache, void* v_cache,
                     int32_t rows, int32_t q_pad, int32_t n_heads, int32_t n_kv_heads, int32_t head_dim,
                     int32_t kv_capacity, int32_t max_pos) {
  if (!qkv || !cos_tab || !sin_tab || !pos || !meta || !q_out || !k_cache || !v_cache) return LADE_EINVAL;
  if (rows < 1 || rows > q_pad || head_dim % 16 != 0 || head_dim > 512 || n_heads < 1 || n_kv_heads < 1) return LADE_EINVAL;
  // one work item (head, 8-wide slice) per thread when they fit: (32 + 2*32) heads * 8 slices = 768 threads at 7B
  int rope_threads = (((n_heads + 2 * n_kv_heads) * (head_dim / 16) + 31) / 32) * 32;
  rope_threads = rope_threads < 128 ? 128 : (rope_threads > 1024 ? 1024 : rope_threads);
  LADE_CUDA_CHECK(launch_pdl(rope_append_kernel, dim3(rows), dim3(rope_threads), 0, (cudaStream_t)stream,
                             (const __nv_bfloat16*)qkv, (const __nv_bfloat16*)cos_tab, (const __nv_bfloat16*)sin_tab,
                             (const int*)pos, (const int*)meta, (__nv_bfloat16*)q_out, (__nv_bfloat16*)k_cache,
                             (__nv_bfloat16*)v_cache, q_pad, n_heads, n_kv_heads, head_dim, kv_capacity, max_pos));
  return LADE_OK;
}

int lade_swiglu(void* stream, const void* gate_up, void* out, int32_t rows, int32_t inter) {
  if (!gate_up || !out || rows < 1 || inter < 8 || inter % 8 != 0) return LADE_EINVAL;
  const long long total = (long long)rows * (inter / 8);
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  LADE_CUDA_CHECK(launch_pdl(swiglu_kernel, dim3(blocks), dim3(256), 0, (cudaStream_t)stream,
                             (const __nv_bfloat16*)gate_up, (__nv_bfloat16*)out, rows, inter));
  return LADE_OK;
}

}  // extern "C"
