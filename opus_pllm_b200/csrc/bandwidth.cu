// HBM-bound kernels of the OPUS-PLLM generation path: embedding gathers, LayerNorm / RMSNorm (+ residual, + split-K
// reduction), rotary embeddings (ESM and Llama flavours, the latter fused with the paged KV-cache append), masked
// mean-pool + L2-normalise, the soft-token splice gather and greedy argmax with EOS bookkeeping.
// All are coalesced 128-bit accesses with warp-shuffle reductions; none of them is reshaped into a GEMM.
#include "common.h"
#include "kernels.h"
#include "launch.cuh"
#include "ptx.cuh"

namespace opus {

namespace {

constexpr int WARPS_PER_BLOCK = 4;

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }

__device__ __forceinline__ void bf16x8_to_float(const uint4& q, float* f) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 float_to_bf16x8(const float* f) {
  uint4 q;
  q.x = pack_bf16x2(f[0], f[1]);
  q.y = pack_bf16x2(f[2], f[3]);
  q.z = pack_bf16x2(f[4], f[5]);
  q.w = pack_bf16x2(f[6], f[7]);
  return q;
}

// ------------------------------------------------------------------------------------------------
// E1: ESM token embedding with token-dropout rescale.   x[i,:] = E[tok[i],:] * scale[i]
// (fair-esm ESM2.forward, called at cstp_v3/modelling.py:48; scale = 0.88/(1-mask_ratio), 0 for <mask> tokens)
// ------------------------------------------------------------------------------------------------
__global__ void esm_embed_kernel(const int* __restrict__ tok, const float* __restrict__ scale,
                                 const float* __restrict__ table, float* __restrict__ x, int n_tok, int dim) {
  grid_dep_launch();
  grid_dep_wait();
  const int row = blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  if (row >= n_tok) return;
  const int lane = threadIdx.x & 31;
  const float s = scale[row];
  const float* src = table + (size_t)tok[row] * dim;
  float* dst = x + (size_t)row * dim;
  for (int c = lane * 4; c < dim; c += 128) {
    float4 v = ld4(src + c);
    v.x *= s; v.y *= s; v.z *= s; v.w *= s;
    *reinterpret_cast<float4*>(dst + c) = v;
  }
}

// ------------------------------------------------------------------------------------------------
// E2: LayerNorm over fp32 rows -> bf16 (the GEMM operand).  One warp per row, values kept in registers.
// ------------------------------------------------------------------------------------------------
// If `delta` != nullptr the pending bf16 branch output (out_proj / fc2 result) is first folded into the fp32 residual
// stream: x <- x + delta (written back), then normalised. `delta` may alias `y` (each lane reads its delta elements
// before it writes the same positions of y).
template <int MAX_V4>  // float4 chunks per lane
__global__ void layernorm_f32_bf16_kernel(float* x, const __nv_bfloat16* delta, const float* __restrict__ gamma,
                                          const float* __restrict__ beta, __nv_bfloat16* y, int rows,
                                          int cols, float eps) {
  grid_dep_launch();
  grid_dep_wait();
  const int row = blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  float* src = x + (size_t)row * cols;
  float4 v[MAX_V4];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < MAX_V4; ++i) {
    const int c = (i * 32 + lane) * 4;
    if (c < cols) {
      v[i] = ld4(src + c);
      if (delta != nullptr) {
        const uint2 dq = *reinterpret_cast<const uint2*>(delta + (size_t)row * cols + c);
        const __nv_bfloat162* d2 = reinterpret_cast<const __nv_bfloat162*>(&dq);
        const float2 da = __bfloat1622float2(d2[0]), db = __bfloat1622float2(d2[1]);
        v[i].x += da.x; v[i].y += da.y; v[i].z += db.x; v[i].w += db.y;
        *reinterpret_cast<float4*>(src + c) = v[i];
      }
      sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    } else {
      v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  const float mean = warp_sum(sum) / cols;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < MAX_V4; ++i) {
    const int c = (i * 32 + lane) * 4;
    if (c < cols) {
      const float a = v[i].x - mean, b = v[i].y - mean, cc = v[i].z - mean, d = v[i].w - mean;
      sq += (a * a + b * b) + (cc * cc + d * d);
    }
  }
  const float rstd = rsqrtf(warp_sum(sq) / cols + eps);
  __nv_bfloat16* dst = y + (size_t)row * cols;
#pragma unroll
  for (int i = 0; i < MAX_V4; ++i) {
    const int c = (i * 32 + lane) * 4;
    if (c < cols) {
      const float4 g = ld4(gamma + c);
      const float4 b = ld4(beta + c);
      uint2 o;
      o.x = pack_bf16x2((v[i].x - mean) * rstd * g.x + b.x, (v[i].y - mean) * rstd * g.y + b.y);
      o.y = pack_bf16x2((v[i].z - mean) * rstd * g.z + b.z, (v[i].w - mean) * rstd * g.w + b.w);
      *reinterpret_cast<uint2*>(dst + c) = o;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// L1: RMSNorm (HF LlamaRMSNorm, modeling_llama.py:62-67): y = w * bf16(x * rsqrt(mean(x^2) + eps)), fp32 inside.
// Optional fused residual add (h = bf16(x + r) written back, then normalised) and optional split-K reduction of fp32
// partial sums:  x_eff = bf16(residual + bf16(sum_s partial[s]))   -- the decode o_proj / down_proj tail.
// TPR threads cooperate on one row (RPC rows per CTA): TPR = 128 for prefill-sized inputs (many rows, few registers,
// high occupancy), TPR = 512 for decode-sized inputs (few rows: all partial-sum loads of a row are issued at once).
// ------------------------------------------------------------------------------------------------
template <int TPR, int RPC, int MAXC, bool PLAIN>
__global__ void __launch_bounds__(TPR* RPC, PLAIN ? 5 : 1)
rmsnorm_bf16_kernel(const __nv_bfloat16* x,                            // [rows, cols] or nullptr if partials
                    const float* __restrict__ partial, int n_partial,  // [n_partial][rows][cols]
                    const __nv_bfloat16* residual,                     // nullable (may alias h_out)
                    __nv_bfloat16* h_out,                              // nullable: x (+ residual) written back
                    const __nv_bfloat16* __restrict__ w, __nv_bfloat16* y, int rows, int cols, float eps) {
  grid_dep_launch();
  grid_dep_wait();
  constexpr int WPR = TPR / 32;  // warps per row
  __shared__ float red[RPC][WPR];
  const int r_in = threadIdx.x / TPR, t = threadIdx.x % TPR;
  const int row = blockIdx.x * RPC + r_in;
  const bool row_ok = row < rows;
  uint4 raw[MAXC];  // the row kept as packed bf16 (h is bf16-exact): 4 registers per 8 elements -> high occupancy
  float sq = 0.f;
  if constexpr (PLAIN) {
    // plain norm (no partial sums, no residual): issue every load of the row before consuming any (memory-level parallelism)
#pragma unroll
    for (int i = 0; i < MAXC; ++i) {
      const int c = (i * TPR + t) * 8;
      raw[i] = (row_ok && c < cols) ? *reinterpret_cast<const uint4*>(x + (size_t)row * cols + c) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int i = 0; i < MAXC; ++i) {
      const int c = (i * TPR + t) * 8;
      if (row_ok && c < cols) {
        float v[8];
        bf16x8_to_float(raw[i], v);
        if (h_out != nullptr) *reinterpret_cast<uint4*>(h_out + (size_t)row * cols + c) = raw[i];
#pragma unroll
        for (int j = 0; j < 8; ++j) sq += v[j] * v[j];
      }
    }
  } else {
#pragma unroll
    for (int i = 0; i < MAXC; ++i) {
      const int c = (i * TPR + t) * 8;
      raw[i] = make_uint4(0, 0, 0, 0);
      if (row_ok && c < cols) {
        const size_t off = (size_t)row * cols + c;
        float v[8];
        if (partial != nullptr) {
          float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
          // four slices per round: all eight loads are in flight before the first add (one L2 round trip instead of
          // n_partial); the adds keep the slice order, so the result is unchanged
          for (int s0 = 0; s0 < n_partial; s0 += 4) {
            float4 a[4], b[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const bool ok = s0 + u < n_partial;
              const float* pp = partial + (size_t)(ok ? s0 + u : s0) * rows * cols + off;
              a[u] = ld4(pp); b[u] = ld4(pp + 4);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              if (s0 + u < n_partial) {
                acc[0] += a[u].x; acc[1] += a[u].y; acc[2] += a[u].z; acc[3] += a[u].w;
                acc[4] += b[u].x; acc[5] += b[u].y; acc[6] += b[u].z; acc[7] += b[u].w;
              }
            }
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = bf16_round(acc[j]);
        } else {
          bf16x8_to_float(*reinterpret_cast<const uint4*>(x + off), v);
        }
        if (residual != nullptr) {
          float r[8];
          bf16x8_to_float(*reinterpret_cast<const uint4*>(residual + off), r);
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = bf16_round(v[j] + r[j]);
        }
        raw[i] = float_to_bf16x8(v);
        if (h_out != nullptr) *reinterpret_cast<uint4*>(h_out + off) = raw[i];
#pragma unroll
        for (int j = 0; j < 8; ++j) sq += v[j] * v[j];
      }
    }
  }
  if (y == nullptr) return;
  sq = warp_sum(sq);
  if ((threadIdx.x & 31) == 0) red[r_in][t >> 5] = sq;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int i = 0; i < WPR; ++i) tot += red[r_in][i];
  const float rstd = rsqrtf(tot / cols + eps);
#pragma unroll
  for (int i = 0; i < MAXC; ++i) {
    const int c = (i * TPR + t) * 8;
    if (row_ok && c < cols) {
      float v[8], wv[8], o[8];
      bf16x8_to_float(raw[i], v);
      bf16x8_to_float(*reinterpret_cast<const uint4*>(w + c), wv);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = wv[j] * bf16_round(v[j] * rstd);
      *reinterpret_cast<uint4*>(y + (size_t)row * cols + c) = float_to_bf16x8(o);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Generic split-K reduction: out bf16 [rows, cols] = sum_s partial[s] (+ bias[col])
// ------------------------------------------------------------------------------------------------
__global__ void splitk_reduce_bf16_kernel(const float* __restrict__ partial, int n_partial,
                                          const float* __restrict__ bias, __nv_bfloat16* __restrict__ out,
                                          size_t rows, int cols, int ldo, int gelu) {
  grid_dep_launch();
  grid_dep_wait();
  const size_t total = rows * (size_t)cols / 4;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t e = i * 4;
    float4 acc = ld4(partial + e);
    for (int s = 1; s < n_partial; ++s) {
      const float4 t = ld4(partial + (size_t)s * rows * cols + e);
      acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
    }
    const size_t r = e / cols;
    const int c = (int)(e - r * cols);
    if (bias != nullptr) {
      const float4 b = ld4(bias + c);
      acc.x += b.x; acc.y += b.y; acc.z += b.z; acc.w += b.w;
    }
    if (gelu) { acc.x = gelu_erf(acc.x); acc.y = gelu_erf(acc.y); acc.z = gelu_erf(acc.z); acc.w = gelu_erf(acc.w); }
    uint2 o;
    o.x = pack_bf16x2(acc.x, acc.y);
    o.y = pack_bf16x2(acc.z, acc.w);
    *reinterpret_cast<uint2*>(out + r * ldo + c) = o;
  }
}

// ------------------------------------------------------------------------------------------------
// E4: ESM rotary embedding on the fused QKV activations, in place.  q <- rope(q * hd^-0.5), k <- rope(k).
// Half-rotation convention: out[j] = x[j]*cos[j] - x[j+hd/2]*sin[j], out[j+hd/2] = x[j+hd/2]*cos[j] + x[j]*sin[j].
// (fair-esm RotaryEmbedding; cross-check HF modeling_esm.py:43-55,81-123,341-344.) cos/sin: fp32 [max_pos, hd/2].
// One thread handles 4 rotation pairs (8 values) of one head of one token.
// ------------------------------------------------------------------------------------------------
__global__ void rope_esm_kernel(__nv_bfloat16* __restrict__ qkv, const int* __restrict__ pos,
                                const float* __restrict__ cos_t, const float* __restrict__ sin_t, int n_tok,
                                int n_heads, int head_dim, int ld, float q_scale) {
  grid_dep_launch();
  grid_dep_wait();
  const int half = head_dim / 2;
  const int per_head = half / 4;                 // threads per head
  const int per_tok = 2 * n_heads * per_head;    // q heads then k heads
  const size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (idx >= (size_t)n_tok * per_tok) return;
  const int tok = (int)(idx / per_tok);
  int r = (int)(idx - (size_t)tok * per_tok);
  const int hsel = r / per_head;                 // 0 .. 2*n_heads-1
  const int j0 = (r - hsel * per_head) * 4;
  const bool is_q = hsel < n_heads;
  __nv_bfloat16* base = qkv + (size_t)tok * ld + (size_t)hsel * head_dim;  // k block follows q block contiguously
  const float sc = is_q ? q_scale : 1.0f;
  const int p = pos[tok];
  const float4 c = ld4(cos_t + (size_t)p * half + j0);
  const float4 s = ld4(sin_t + (size_t)p * half + j0);
  uint2 lo = *reinterpret_cast<const uint2*>(base + j0);
  uint2 hi = *reinterpret_cast<const uint2*>(base + half + j0);
  const __nv_bfloat162* l2 = reinterpret_cast<const __nv_bfloat162*>(&lo);
  const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&hi);
  const float2 la = __bfloat1622float2(l2[0]), lb = __bfloat1622float2(l2[1]);
  const float2 ha = __bfloat1622float2(h2[0]), hb = __bfloat1622float2(h2[1]);
  const float x1[4] = {la.x * sc, la.y * sc, lb.x * sc, lb.y * sc};
  const float x2[4] = {ha.x * sc, ha.y * sc, hb.x * sc, hb.y * sc};
  const float cc[4] = {c.x, c.y, c.z, c.w};
  const float ss[4] = {s.x, s.y, s.z, s.w};
  float o1[4], o2[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    o1[i] = x1[i] * cc[i] - x2[i] * ss[i];
    o2[i] = x2[i] * cc[i] + x1[i] * ss[i];
  }
  uint2 w1, w2;
  w1.x = pack_bf16x2(o1[0], o1[1]); w1.y = pack_bf16x2(o1[2], o1[3]);
  w2.x = pack_bf16x2(o2[0], o2[1]); w2.y = pack_bf16x2(o2[2], o2[3]);
  *reinterpret_cast<uint2*>(base + j0) = w1;
  *reinterpret_cast<uint2*>(base + half + j0) = w2;
}

// ------------------------------------------------------------------------------------------------
// L3+L4: Llama rotary embedding fused with the paged KV-cache append.
// HF apply_rotary_pos_emb (modeling_llama.py:124-168) runs in the activation dtype: cos/sin are cast to bf16 and each
// product and the sum round to bf16; the same rounding points are kept here so q/k match bit-for-bit.
//   q: rotated in place in qkv.   k: rotated, written back in place (prefill attention reads it) AND to the cache.
//   v: copied to the cache.       slot[tok] = physical_block * block_size + offset, or < 0 to skip the cache write.
// Cache layout: [num_blocks][n_kv_heads][block_size][head_dim] bf16 (one contiguous [block_size, head_dim] panel per
// (block, kv head) so decode attention streams 4 KB panels).
// Optional split-K input: when `partial` != nullptr the bf16 qkv row is first formed as bf16(sum_s partial[s]) (decode).
// ------------------------------------------------------------------------------------------------
__global__ void rope_llama_kvappend_kernel(__nv_bfloat16* __restrict__ qkv, const float* __restrict__ partial,
                                           int n_partial, const int* __restrict__ pos, const int* __restrict__ slot,
                                           const __nv_bfloat16* __restrict__ cos_t,
                                           const __nv_bfloat16* __restrict__ sin_t, __nv_bfloat16* __restrict__ kcache,
                                           __nv_bfloat16* __restrict__ vcache, int n_tok, int n_q_heads,
                                           int n_kv_heads, int head_dim, int ld, int block_size,
                                           const float* __restrict__ bias) {
  grid_dep_launch();
  grid_dep_wait();
  const int half = head_dim / 2;
  const int per_head = half / 8;  // threads per head (each: 8 rotation pairs, 16-byte accesses)
  const int heads_total = n_q_heads + 2 * n_kv_heads;
  const int per_tok = heads_total * per_head;
  const size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (idx >= (size_t)n_tok * per_tok) return;
  const int tok = (int)(idx / per_tok);
  const int r = (int)(idx - (size_t)tok * per_tok);
  const int hsel = r / per_head;
  const int j0 = (r - hsel * per_head) * 8;
  __nv_bfloat16* base = qkv + (size_t)tok * ld + (size_t)hsel * head_dim;

  float x1[8], x2[8];
  if (partial != nullptr) {
    const size_t off = (size_t)tok * ld + (size_t)hsel * head_dim;
    float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, b[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int s0 = 0; s0 < n_partial; s0 += 3) {   // three slices per round, loads first (see rmsnorm_bf16_kernel)
      float4 t1[3], t2[3], t3[3], t4[3];
#pragma unroll
      for (int u = 0; u < 3; ++u) {
        const bool ok = s0 + u < n_partial;
        const float* pp = partial + (size_t)(ok ? s0 + u : s0) * n_tok * ld + off;
        t1[u] = ld4(pp + j0); t2[u] = ld4(pp + j0 + 4); t3[u] = ld4(pp + half + j0); t4[u] = ld4(pp + half + j0 + 4);
      }
#pragma unroll
      for (int u = 0; u < 3; ++u) {
        if (s0 + u < n_partial) {
          a[0] += t1[u].x; a[1] += t1[u].y; a[2] += t1[u].z; a[3] += t1[u].w;
          a[4] += t2[u].x; a[5] += t2[u].y; a[6] += t2[u].z; a[7] += t2[u].w;
          b[0] += t3[u].x; b[1] += t3[u].y; b[2] += t3[u].z; b[3] += t3[u].w;
          b[4] += t4[u].x; b[5] += t4[u].y; b[6] += t4[u].z; b[7] += t4[u].w;
        }
      }
    }
    if (bias != nullptr) {   // q|k|v projection bias (Qwen2): added to the fp32 sum before the one bf16 rounding
      const float* bb = bias + (size_t)hsel * head_dim;
      const float4 b1 = ld4(bb + j0), b2 = ld4(bb + j0 + 4), b3 = ld4(bb + half + j0), b4 = ld4(bb + half + j0 + 4);
      a[0] += b1.x; a[1] += b1.y; a[2] += b1.z; a[3] += b1.w; a[4] += b2.x; a[5] += b2.y; a[6] += b2.z; a[7] += b2.w;
      b[0] += b3.x; b[1] += b3.y; b[2] += b3.z; b[3] += b3.w; b[4] += b4.x; b[5] += b4.y; b[6] += b4.z; b[7] += b4.w;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) { x1[i] = bf16_round(a[i]); x2[i] = bf16_round(b[i]); }
  } else {
    bf16x8_to_float(*reinterpret_cast<const uint4*>(base + j0), x1);
    bf16x8_to_float(*reinterpret_cast<const uint4*>(base + half + j0), x2);
  }

  const bool is_v = hsel >= n_q_heads + n_kv_heads;
  float o1[8], o2[8];
  if (!is_v) {
    const int p = pos[tok];
    // HF builds the tables as cat(freqs, freqs): column j + half holds the same angle as column j, so one cos / sin
    // load serves both halves (the table loads were 2/3 of this kernel's load instructions)
    float c1[8], s1[8];
    bf16x8_to_float(__ldg(reinterpret_cast<const uint4*>(cos_t + (size_t)p * head_dim + j0)), c1);
    bf16x8_to_float(__ldg(reinterpret_cast<const uint4*>(sin_t + (size_t)p * head_dim + j0)), s1);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      // first half: x1*cos + (-x2)*sin ; second half: x2*cos + x1*sin ; every op rounds to bf16 like torch
      o1[i] = bf16_round(bf16_round(x1[i] * c1[i]) + bf16_round(-x2[i] * s1[i]));
      o2[i] = bf16_round(bf16_round(x2[i] * c1[i]) + bf16_round(x1[i] * s1[i]));
    }
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) { o1[i] = x1[i]; o2[i] = x2[i]; }
  }
  const uint4 w1 = float_to_bf16x8(o1), w2 = float_to_bf16x8(o2);
  if (!is_v || partial != nullptr) {
    *reinterpret_cast<uint4*>(base + j0) = w1;
    *reinterpret_cast<uint4*>(base + half + j0) = w2;
  }
  if (hsel >= n_q_heads && slot != nullptr) {
    const int sl = slot[tok];
    if (sl >= 0) {
      const int kvh = is_v ? hsel - n_q_heads - n_kv_heads : hsel - n_q_heads;
      const int blk = sl / block_size, off = sl - blk * block_size;
      __nv_bfloat16* dst = (is_v ? vcache : kcache) + (((size_t)blk * n_kv_heads + kvh) * block_size + off) * head_dim;
      *reinterpret_cast<uint4*>(dst + j0) = w1;
      *reinterpret_cast<uint4*>(dst + half + j0) = w2;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// E2(final)+E10+P1 head: final LayerNorm of every residue, mean over residues [1, len-1) (drops <cls>/<eos>),
// then L2-normalise.  (cstp_v3/modelling.py:53-55 and F.normalize at :398.)  One CTA per sequence; each warp walks a
// strided subset of the residues, keeps a running fp32 sum of the normalised rows, CTA reduces through smem.
// Outputs: pooled fp32 [B, dim] (the reference's return value) and l2-normalised bf16 [B, dim] (projector GEMM input).
// ------------------------------------------------------------------------------------------------
template <int MAX_V4>
__global__ void final_ln_meanpool_kernel(const float* __restrict__ x, const __nv_bfloat16* __restrict__ delta,
                                         const int* __restrict__ cu_seqlens,
                                         const float* __restrict__ gamma, const float* __restrict__ beta,
                                         float* __restrict__ pooled, __nv_bfloat16* __restrict__ pooled_l2_bf16,
                                         float* __restrict__ hidden_out, int dim, float eps) {
  grid_dep_launch();
  grid_dep_wait();
  extern __shared__ float red[];  // [nwarps][dim]
  const int b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int start = cu_seqlens[b], end = cu_seqlens[b + 1];
  float4 acc[MAX_V4];
#pragma unroll
  for (int i = 0; i < MAX_V4; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);

  for (int t = start + warp; t < end; t += nwarps) {
    const float* src = x + (size_t)t * dim;
    float4 v[MAX_V4];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < MAX_V4; ++i) {
      const int c = (i * 32 + lane) * 4;
      v[i] = (c < dim) ? ld4(src + c) : make_float4(0.f, 0.f, 0.f, 0.f);
      if (delta != nullptr && c < dim) {
        const uint2 dq = *reinterpret_cast<const uint2*>(delta + (size_t)t * dim + c);
        const __nv_bfloat162* d2 = reinterpret_cast<const __nv_bfloat162*>(&dq);
        const float2 da = __bfloat1622float2(d2[0]), db = __bfloat1622float2(d2[1]);
        v[i].x += da.x; v[i].y += da.y; v[i].z += db.x; v[i].w += db.y;
      }
      sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
    const float mean = warp_sum(sum) / dim;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < MAX_V4; ++i) {
      const int c = (i * 32 + lane) * 4;
      if (c < dim) {
        const float a = v[i].x - mean, bb = v[i].y - mean, cc = v[i].z - mean, d = v[i].w - mean;
        sq += (a * a + bb * bb) + (cc * cc + d * d);
      }
    }
    const float rstd = rsqrtf(warp_sum(sq) / dim + eps);
    const bool pooled_row = (t > start) && (t < end - 1);
#pragma unroll
    for (int i = 0; i < MAX_V4; ++i) {
      const int c = (i * 32 + lane) * 4;
      if (c < dim) {
        const float4 g = ld4(gamma + c), be = ld4(beta + c);
        float4 o;
        o.x = (v[i].x - mean) * rstd * g.x + be.x;
        o.y = (v[i].y - mean) * rstd * g.y + be.y;
        o.z = (v[i].z - mean) * rstd * g.z + be.z;
        o.w = (v[i].w - mean) * rstd * g.w + be.w;
        if (hidden_out != nullptr) *reinterpret_cast<float4*>(hidden_out + (size_t)t * dim + c) = o;
        if (pooled_row) { acc[i].x += o.x; acc[i].y += o.y; acc[i].z += o.z; acc[i].w += o.w; }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < MAX_V4; ++i) {
    const int c = (i * 32 + lane) * 4;
    if (c < dim) *reinterpret_cast<float4*>(red + (size_t)warp * dim + c) = acc[i];
  }
  __syncthreads();
  const int n_res = end - start - 2;
  const float inv = n_res > 0 ? 1.0f / n_res : 0.f;  // torch: mean of an empty slice is NaN; we return 0 and flag on host
  float local_sq = 0.f;
  for (int c = threadIdx.x; c < dim; c += blockDim.x) {
    float s = 0.f;
    for (int w = 0; w < nwarps; ++w) s += red[(size_t)w * dim + c];
    s *= inv;
    red[c] = s;  // safe: column c of warp 0's slot is only touched by this thread
    pooled[(size_t)b * dim + c] = s;
    local_sq += s * s;
  }
  __shared__ float wsum[32];
  local_sq = warp_sum(local_sq);
  if (lane == 0) wsum[warp] = local_sq;
  __syncthreads();
  float tot = 0.f;
  for (int w = 0; w < nwarps; ++w) tot += wsum[w];
  const float denom = fmaxf(sqrtf(tot), 1e-12f);
  if (pooled_l2_bf16 != nullptr)
    for (int c = threadIdx.x; c < dim; c += blockDim.x)
      pooled_l2_bf16[(size_t)b * dim + c] = __float2bfloat16_rn(red[c] / denom);
}

// P1 alone: rows fp32 -> L2-normalised bf16 (pre-computed-embedding path, opus_arch.py:151-161).
__global__ void l2norm_f32_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, int rows, int dim) {
  grid_dep_launch();
  grid_dep_wait();
  const int row = blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  float sq = 0.f;
  for (int c = lane; c < dim; c += 32) { const float v = x[(size_t)row * dim + c]; sq += v * v; }
  const float denom = fmaxf(sqrtf(warp_sum(sq)), 1e-12f);
  for (int c = lane; c < dim; c += 32) y[(size_t)row * dim + c] = __float2bfloat16_rn(x[(size_t)row * dim + c] / denom);
}

// ------------------------------------------------------------------------------------------------
// S1: soft-token splice gather (opus_arch.py:176-270).  out[i,:] = src[i] >= 0 ? embed[src[i],:] :
//     src[i] == INT_MIN ? 0 (pad) : soft[-src[i]-1, :].   One warp per output row, 16-byte copies.
// ------------------------------------------------------------------------------------------------
__global__ void splice_gather_kernel(const int* __restrict__ src, const __nv_bfloat16* __restrict__ embed,
                                     const __nv_bfloat16* __restrict__ soft, __nv_bfloat16* __restrict__ out,
                                     int n_rows, int dim) {
  grid_dep_launch();
  grid_dep_wait();
  const int row = blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  if (row >= n_rows) return;
  const int lane = threadIdx.x & 31;
  const int s = src[row];
  uint4* dst = reinterpret_cast<uint4*>(out + (size_t)row * dim);
  const int n16 = dim / 8;
  if (s == INT_MIN) {
    for (int c = lane; c < n16; c += 32) dst[c] = make_uint4(0, 0, 0, 0);
    return;
  }
  const uint4* from = reinterpret_cast<const uint4*>(s >= 0 ? embed + (size_t)s * dim : soft + (size_t)(-s - 1) * dim);
  for (int c = lane; c < n16; c += 32) dst[c] = from[c];
}

// ------------------------------------------------------------------------------------------------
// L10: greedy selection + EOS bookkeeping (HF GenerationMixin._sample with do_sample=False, generation/utils.py).
// argmax over bf16 logits (ties -> lowest index, like torch.argmax); finished rows emit pad; a row finishes when it
// emits an EOS id. Also advances the per-row context length and records the token for the next step's embedding.
// ------------------------------------------------------------------------------------------------
__global__ void argmax_eos_kernel(const __nv_bfloat16* __restrict__ logits, int ld, int vocab,
                                  int* __restrict__ finished, const int* __restrict__ eos_ids, int n_eos, int pad_id,
                                  int* __restrict__ next_tok, int* __restrict__ out_ids, int out_ld, int step_imm,
                                  const int* __restrict__ step_ptr, int* __restrict__ n_unfinished) {
  grid_dep_launch();
  grid_dep_wait();
  const int b = blockIdx.x;
  const __nv_bfloat16* row = logits + (size_t)b * ld;
  float best = -INFINITY;
  int best_i = INT_MAX;
  for (int c = threadIdx.x * 8; c < vocab; c += blockDim.x * 8) {
    float f[8];
    bf16x8_to_float(*reinterpret_cast<const uint4*>(row + c), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (c + j < vocab && (f[j] > best || (f[j] == best && c + j < best_i))) { best = f[j]; best_i = c + j; }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
    if (ob > best || (ob == best && oi < best_i)) { best = ob; best_i = oi; }
  }
  __shared__ float sb[32];
  __shared__ int si[32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  if (lane == 0) { sb[warp] = best; si[warp] = best_i; }
  __syncthreads();
  if (warp == 0) {
    best = lane < nw ? sb[lane] : -INFINITY;
    best_i = lane < nw ? si[lane] : INT_MAX;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
      if (ob > best || (ob == best && oi < best_i)) { best = ob; best_i = oi; }
    }
    if (lane == 0) {
      int tok = best_i;
      const int step = step_ptr != nullptr ? *step_ptr : step_imm;
      const int fin = finished[b];
      if (fin) tok = pad_id;
      next_tok[b] = tok;
      out_ids[(size_t)b * out_ld + step] = tok;
      if (!fin) {
        bool is_eos = false;
        for (int e = 0; e < n_eos; ++e) is_eos |= (tok == eos_ids[e]);
        if (is_eos) {
          finished[b] = 1;
          if (n_unfinished != nullptr) atomicSub(n_unfinished, 1);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// L10b: temperature + nucleus (top-p) sampling with the same EOS bookkeeping as argmax_eos_kernel.
// HF semantics (GenerationMixin._sample, do_sample=True; TemperatureLogitsWarper then TopPLogitsWarper with
// min_tokens_to_keep = 1; the reference's default decode is temperature 0.1 / top_p 0.7, run_opus_ddp.py:126-128,156-157):
// scores = fp32(logits) / T; a token is KEPT iff the probability mass of the tokens ranked strictly above it is < top_p;
// the next token is drawn from the renormalised kept set.
// One CTA per row. The logits are bf16, so ranks are decided on their 16-bit order-preserving keys: a 16-step bisection
// finds the smallest key k* with mass(key > k*) < top_p * Z (every pass re-reads the row from L2, 256 KB). Ties at k* are
// kept lowest-index first (HF keeps them by its sort order; equal logits have equal probability). The draw uses a
// counter-based generator, u = hash(seed, row, step), so a (seed, prompt batch) pair always yields the same tokens; it
// cannot reproduce torch.multinomial's stream, tests compare the kept set and the sampling frequencies instead.
// ------------------------------------------------------------------------------------------------
constexpr int SAMPLE_THREADS = 1024;

__device__ __forceinline__ uint32_t bf16_key(uint16_t b) { return (b & 0x8000u) ? (uint16_t)~b : (uint16_t)(b | 0x8000u); }
__device__ __forceinline__ float key_value(uint32_t key) {
  const uint16_t b = (key & 0x8000u) ? (uint16_t)(key & 0x7fffu) : (uint16_t)~key;
  return __uint_as_float((uint32_t)b << 16);
}

// block-wide sum in a fixed order (deterministic); result broadcast to every thread. `red` = 33 floats of smem.
__device__ __forceinline__ float block_sum_1024(float v, float* red) {
  v = warp_sum(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  if (warp == 0) {
    float t = red[lane];
    t = warp_sum(t);
    if (lane == 0) red[32] = t;
  }
  __syncthreads();
  return red[32];
}

__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

__global__ void __launch_bounds__(SAMPLE_THREADS)
sample_top_p_kernel(const __nv_bfloat16* __restrict__ logits, int ld, int vocab, float inv_temp, float top_p,
                    unsigned long long seed, int* __restrict__ finished, const int* __restrict__ eos_ids, int n_eos,
                    int pad_id, int* __restrict__ next_tok, int* __restrict__ out_ids, int out_ld, int step_imm,
                    const int* __restrict__ step_ptr, int* __restrict__ n_unfinished, int* __restrict__ kept_count,
                    const unsigned long long* __restrict__ seed_ptr) {
  grid_dep_launch();
  grid_dep_wait();
  if (seed_ptr != nullptr) seed = *seed_ptr;   // seed in device memory: one captured graph serves every seed
  __shared__ float red[33];
  __shared__ float scan_m[SAMPLE_THREADS];
  __shared__ int scan_c[SAMPLE_THREADS];
  const int b = blockIdx.x;
  const uint16_t* row = reinterpret_cast<const uint16_t*>(logits) + (size_t)b * ld;
  const int tid = threadIdx.x;
  const float c = inv_temp * 1.4426950408889634f;

  // pass 1: maximum key
  uint32_t kmax = 0;
  for (int i = tid; i < vocab; i += SAMPLE_THREADS) kmax = max(kmax, bf16_key(row[i]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) kmax = max(kmax, __shfl_xor_sync(0xffffffffu, kmax, o));
  __shared__ uint32_t s_k[32];
  if ((tid & 31) == 0) s_k[tid >> 5] = kmax;
  __syncthreads();
  kmax = s_k[0];
#pragma unroll
  for (int w = 1; w < SAMPLE_THREADS / 32; ++w) kmax = max(kmax, s_k[w]);
  const float vmax = key_value(kmax);

  // mass of the tokens whose key is strictly greater than k (k = -1: everything = Z)
  auto mass_gt = [&](int k) {
    float acc = 0.f;
    for (int i = tid; i < vocab; i += SAMPLE_THREADS) {
      const uint32_t key = bf16_key(row[i]);
      if ((int)key > k) acc += exp2f((key_value(key) - vmax) * c);
    }
    return block_sum_1024(acc, red);
  };
  const float Z = mass_gt(-1);
  const float need = top_p * Z;
  // smallest key k* with mass_gt(k*) < need  (mass_gt(kmax) = 0 always qualifies)
  int lo = -1, hi = (int)kmax;          // invariant: mass_gt(lo) >= need (or lo = -1 and top_p >= 1), mass_gt(hi) < need
  float g_hi = 0.f;
  if (Z < need || !(top_p < 1.0f)) {     // top_p >= 1: keep everything
    hi = 0;
    g_hi = mass_gt(0);
    lo = -1;
    // all keys >= 0 kept: treat k* = 0 with every tie kept (handled below by n_keep = INT_MAX)
  } else {
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      const float g = mass_gt(mid);
      if (g < need) { hi = mid; g_hi = g; } else { lo = mid; }
    }
  }
  const int kstar = hi;
  float w_star = exp2f((key_value((uint32_t)kstar) - vmax) * c);
  if (!(w_star >= 0.f) || !(w_star <= 1.f)) w_star = 0.f;   // k* need not be a key that occurs (e.g. a NaN pattern)
  int n_keep = INT_MAX;
  if (top_p < 1.0f && Z >= need) {
    const float room = need - g_hi;                       // > 0
    const float q = w_star > 0.f ? ceilf(room / w_star) : 1.0f;
    n_keep = q < 1.f ? 1 : (q > 2.0e9f ? INT_MAX : (int)q);   // min_tokens_to_keep = 1
  }

  // final pass, index order: each thread owns a contiguous range
  const int per = (vocab + SAMPLE_THREADS - 1) / SAMPLE_THREADS;
  const int i0 = min(tid * per, vocab), i1 = min(i0 + per, vocab);
  float m1 = 0.f;
  int c1 = 0, g1 = 0;
  for (int i = i0; i < i1; ++i) {
    const int key = (int)bf16_key(row[i]);
    if (key > kstar) { m1 += exp2f((key_value((uint32_t)key) - vmax) * c); ++g1; }
    else if (key == kstar) ++c1;
  }
  scan_m[tid] = m1;
  scan_c[tid] = c1;
  const float n_above = block_sum_1024((float)g1, red);   // exact: counts <= 2^24
  __syncthreads();
  if (tid == 0) {
    // sequential exclusive scan by one thread: 1024 steps, deterministic, negligible next to the 18 passes above
    const int step = step_ptr != nullptr ? *step_ptr : step_imm;
    float M = 0.f;
    int Cn = 0;
    for (int t = 0; t < SAMPLE_THREADS; ++t) {
      const float mt = scan_m[t];
      const int ct = scan_c[t];
      scan_m[t] = M;
      scan_c[t] = Cn;
      M += mt;
      Cn += ct;
    }
    const int ties_kept = min(Cn, n_keep);
    const float K = M + (float)ties_kept * w_star;        // kept mass
    const uint64_t h = splitmix64(seed ^ splitmix64(((uint64_t)(uint32_t)b << 32) | (uint32_t)step));
    const float u = (float)(h >> 40) * (1.0f / 16777216.0f);   // [0, 1)
    const float target = u * K;
    // kept mass in front of range t: scan_m[t] + min(scan_c[t], n_keep) * w*. The draw falls into the last range whose
    // front mass is <= target; walk that range token by token.
    int tr = 0;
    for (int t = 1; t < SAMPLE_THREADS; ++t)
      if (scan_m[t] + (float)min(scan_c[t], n_keep) * w_star <= target) tr = t;
    int tok = -1, last_kept = -1;
    for (; tr < SAMPLE_THREADS && tok < 0; ++tr) {   // normally one range; continues only across rounding slack
      float cum = scan_m[tr] + (float)min(scan_c[tr], n_keep) * w_star;
      int ties_seen = scan_c[tr];
      const int j0 = min(tr * per, vocab), j1 = min(j0 + per, vocab);
      for (int i = j0; i < j1; ++i) {
        const int key = (int)bf16_key(row[i]);
        bool kept = false;
        float wi = 0.f;
        if (key > kstar) { kept = true; wi = exp2f((key_value((uint32_t)key) - vmax) * c); }
        else if (key == kstar) { kept = ties_seen < n_keep; wi = kept ? w_star : 0.f; ++ties_seen; }
        if (kept) {
          last_kept = i;
          cum += wi;
          if (cum > target) { tok = i; break; }
        }
      }
    }
    if (tok < 0) {
      // target sat in the rounding slack at the very end of the kept mass: take the last kept token
      if (last_kept < 0) {
        for (int i = vocab - 1; i >= 0 && last_kept < 0; --i)
          if ((int)bf16_key(row[i]) > kstar) last_kept = i;
      }
      tok = last_kept >= 0 ? last_kept : 0;
    }
    if (kept_count != nullptr) kept_count[b] = (int)n_above + ties_kept;
    const int fin = finished[b];
    if (fin) tok = pad_id;
    next_tok[b] = tok;
    out_ids[(size_t)b * out_ld + step] = tok;
    if (!fin) {
      bool is_eos = false;
      for (int e = 0; e < n_eos; ++e) is_eos |= (tok == eos_ids[e]);
      if (is_eos) {
        finished[b] = 1;
        if (n_unfinished != nullptr) atomicSub(n_unfinished, 1);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Teacher-forced scoring tail (HF LlamaForCausalLM.forward with labels=, reached from opus_llama.py:41-93): per-row
// cross entropy in fp32 over bf16 logits, loss[r] = logsumexp(logits[r, :]) - logits[r, target[r]]; rows whose target is
// negative (HF's ignore_index = -100) get 0. One CTA per row; the mean over the counted rows is the caller's.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512)
cross_entropy_rows_kernel(const __nv_bfloat16* __restrict__ logits, int ld, int vocab,
                          const int* __restrict__ target, float* __restrict__ loss) {
  const int r = blockIdx.x;
  const int tgt = target[r];
  if (tgt < 0 || tgt >= vocab) {
    if (threadIdx.x == 0) loss[r] = 0.f;
    return;
  }
  const __nv_bfloat16* row = logits + (size_t)r * ld;
  __shared__ float red[17];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float mx = -INFINITY;
  for (int c = threadIdx.x * 8; c < vocab; c += blockDim.x * 8) {
    float f[8];
    bf16x8_to_float(*reinterpret_cast<const uint4*>(row + c), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) if (c + j < vocab) mx = fmaxf(mx, f[j]);
  }
  mx = warp_max(mx);
  if (lane == 0) red[warp] = mx;
  __syncthreads();
  mx = red[0];
#pragma unroll
  for (int w = 1; w < 16; ++w) mx = fmaxf(mx, red[w]);
  __syncthreads();
  float sum = 0.f;
  for (int c = threadIdx.x * 8; c < vocab; c += blockDim.x * 8) {
    float f[8];
    bf16x8_to_float(*reinterpret_cast<const uint4*>(row + c), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) if (c + j < vocab) sum += expf(f[j] - mx);
  }
  sum = warp_sum(sum);
  if (lane == 0) red[warp] = sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
#pragma unroll
    for (int w = 0; w < 16; ++w) tot += red[w];
    loss[r] = logf(tot) + mx - __bfloat162float(row[tgt]);
  }
}

// decode input: x[b,:] = table[tok[b],:]; also advances positions / cache slots for the step (one launch per step).
__global__ void embed_gather_kernel(const int* __restrict__ tok, const __nv_bfloat16* __restrict__ table,
                                    __nv_bfloat16* __restrict__ x, int n_rows, int dim) {
  grid_dep_launch();
  grid_dep_wait();
  const int row = blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  if (row >= n_rows) return;
  const int lane = threadIdx.x & 31;
  const uint4* from = reinterpret_cast<const uint4*>(table + (size_t)tok[row] * dim);
  uint4* dst = reinterpret_cast<uint4*>(x + (size_t)row * dim);
  for (int c = lane; c < dim / 8; c += 32) dst[c] = from[c];
}

// decode input for the norm-fused decode path: x[b,:] = table[tok[b],:] and, per 32-feature slab, the sum of squares
// of the row (sumsq[slab][b]; the consuming GEMM derives the RMSNorm row scale from it). One CTA per row, one thread per
// slab (64 bytes = 4 x 16-byte loads).
__global__ void embed_gather_sumsq_kernel(const int* __restrict__ tok, const __nv_bfloat16* __restrict__ table,
                                          __nv_bfloat16* __restrict__ x, float* __restrict__ sumsq, int sumsq_ld,
                                          int dim) {
  grid_dep_launch();
  grid_dep_wait();
  const int row = blockIdx.x;
  const uint4* from = reinterpret_cast<const uint4*>(table + (size_t)tok[row] * dim);
  uint4* dst = reinterpret_cast<uint4*>(x + (size_t)row * dim);
  for (int slab = threadIdx.x; slab < dim / 32; slab += blockDim.x) {
    float sq = 0.f;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const uint4 q = from[slab * 4 + c];
      dst[slab * 4 + c] = q;
      const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __bfloat1622float2(h2[j]);
        sq += f.x * f.x;
        sq += f.y * f.y;
      }
    }
    sumsq[(size_t)slab * sumsq_ld + row] = sq;
  }
}

// per-step decode bookkeeping: pos[b] = ctx_len[b]; slot[b] = block_table[b][ctx/bs]*bs + ctx%bs; ctx_len[b]++.
__global__ void decode_advance_kernel(int* __restrict__ ctx_len, int* __restrict__ pos, int* __restrict__ slot,
                                      const int* __restrict__ block_table, int max_blocks, int block_size, int n,
                                      int* __restrict__ step) {
  grid_dep_launch();
  grid_dep_wait();
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b == 0 && step != nullptr) *step += 1;
  if (b >= n) return;
  const int c = ctx_len[b];
  pos[b] = c;
  slot[b] = block_table[(size_t)b * max_blocks + c / block_size] * block_size + c % block_size;
  ctx_len[b] = c + 1;
}

// LoRA merge (peft merge_and_unload semantics, model/builder.py:107-109): W += scale * (B @ A), fp32 math, bf16 weights.
// W [out, in], A [r, in], B [out, r].  One-off at load time; r is small (<= 64) so a direct kernel is enough.
__global__ void lora_merge_kernel(__nv_bfloat16* __restrict__ W, const __nv_bfloat16* __restrict__ A,
                                  const __nv_bfloat16* __restrict__ Bm, int out_f, int in_f, int r, float scale) {
  grid_dep_launch();
  grid_dep_wait();
  const size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (idx >= (size_t)out_f * in_f) return;
  const int o = (int)(idx / in_f), i = (int)(idx - (size_t)o * in_f);
  float acc = 0.f;
  for (int k = 0; k < r; ++k) acc += __bfloat162float(Bm[(size_t)o * r + k]) * __bfloat162float(A[(size_t)k * in_f + i]);
  W[idx] = __float2bfloat16_rn(__bfloat162float(W[idx]) + scale * acc);
}

inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }
inline int ok() {
  note_launch();
  return cudaGetLastError() == cudaSuccess ? OPUS_OK : OPUS_ERR_CUDA;
}

}  // namespace

int esm_embed(const int* tok, const float* scale, const float* table, float* x, int n_tok, int dim, cudaStream_t st) {
  if (dim % 4) return OPUS_ERR_ARG;
  if (n_tok == 0) return OPUS_OK;
  launch_pdl(n_tok <= 1024, esm_embed_kernel, dim3(cdiv(n_tok, WARPS_PER_BLOCK)), dim3(WARPS_PER_BLOCK * 32), 0, st, tok, scale, table, x, n_tok, dim);
  return ok();
}

int layernorm_f32_bf16(float* x, const __nv_bfloat16* delta, const float* gamma, const float* beta, __nv_bfloat16* y,
                       int rows, int cols, float eps, cudaStream_t st) {
  if (cols % 4 || cols > 128 * 16) return OPUS_ERR_ARG;
  if (rows == 0) return OPUS_OK;
  const int grid = cdiv(rows, WARPS_PER_BLOCK), blk = WARPS_PER_BLOCK * 32;
  if (cols <= 128 * 4) launch_pdl(rows <= 1024, layernorm_f32_bf16_kernel<4>, dim3(grid), dim3(blk), 0, st, x, delta, gamma, beta, y, rows, cols, eps);
  else if (cols <= 128 * 10) launch_pdl(rows <= 1024, layernorm_f32_bf16_kernel<10>, dim3(grid), dim3(blk), 0, st, x, delta, gamma, beta, y, rows, cols, eps);
  else launch_pdl(rows <= 1024, layernorm_f32_bf16_kernel<16>, dim3(grid), dim3(blk), 0, st, x, delta, gamma, beta, y, rows, cols, eps);
  return ok();
}

int rmsnorm_bf16(const __nv_bfloat16* x, const float* partial, int n_partial, const __nv_bfloat16* residual,
                 __nv_bfloat16* h_out, const __nv_bfloat16* w, __nv_bfloat16* y, int rows, int cols, float eps,
                 cudaStream_t st) {
  if (cols % 8 || cols > 8192) return OPUS_ERR_ARG;
  if ((x == nullptr) == (partial == nullptr)) return OPUS_ERR_ARG;
  if (rows == 0) return OPUS_OK;
  const bool plain = partial == nullptr && residual == nullptr;
  if (rows <= 1024) {  // decode-sized: one 512-thread CTA per row
    if (cols <= 4096)
      launch_pdl(true, rmsnorm_bf16_kernel<512, 1, 1, false>, dim3(rows), dim3(512), 0, st, x, partial, n_partial, residual, h_out, w, y, rows, cols, eps);
    else
      launch_pdl(true, rmsnorm_bf16_kernel<512, 1, 2, false>, dim3(rows), dim3(512), 0, st, x, partial, n_partial, residual, h_out, w, y, rows, cols, eps);
  } else {             // prefill-sized: 128 threads per row, two rows per CTA
    const int grid = cdiv(rows, 2);
    if (cols <= 4096 && plain)
      launch_pdl(false, rmsnorm_bf16_kernel<128, 2, 4, true>, dim3(grid), dim3(256), 0, st, x, partial, n_partial, residual, h_out, w, y, rows, cols, eps);
    else if (cols <= 4096)
      launch_pdl(false, rmsnorm_bf16_kernel<128, 2, 4, false>, dim3(grid), dim3(256), 0, st, x, partial, n_partial, residual, h_out, w, y, rows, cols, eps);
    else
      launch_pdl(false, rmsnorm_bf16_kernel<128, 2, 8, false>, dim3(grid), dim3(256), 0, st, x, partial, n_partial, residual, h_out, w, y, rows, cols, eps);
  }
  return ok();
}

int splitk_reduce_bf16(const float* partial, int n_partial, const float* bias, __nv_bfloat16* out, int rows, int cols,
                       int ldo, int gelu, cudaStream_t st) {
  if (cols % 4 || ldo % 4) return OPUS_ERR_ARG;
  if (rows == 0) return OPUS_OK;
  const long long total = (long long)rows * cols / 4;
  int grid = cdiv(total, 256);
  if (grid > 148 * 8) grid = 148 * 8;
  launch_pdl(rows <= 1024, splitk_reduce_bf16_kernel, dim3(grid), dim3(256), 0, st, partial, n_partial, bias, out, (size_t)rows, cols, ldo, gelu);
  return ok();
}

int rope_esm(__nv_bfloat16* qkv, const int* pos, const float* cos_t, const float* sin_t, int n_tok, int n_heads,
             int head_dim, int ld, float q_scale, cudaStream_t st) {
  if (head_dim % 8 || ld % 4) return OPUS_ERR_ARG;
  if (n_tok == 0) return OPUS_OK;
  const long long total = (long long)n_tok * 2 * n_heads * (head_dim / 8);
  launch_pdl(n_tok <= 1024, rope_esm_kernel, dim3(cdiv(total, 256)), dim3(256), 0, st, qkv, pos, cos_t, sin_t, n_tok, n_heads, head_dim, ld, q_scale);
  return ok();
}

int rope_llama_kvappend(__nv_bfloat16* qkv, const float* partial, int n_partial, const int* pos, const int* slot,
                        const __nv_bfloat16* cos_t, const __nv_bfloat16* sin_t, __nv_bfloat16* kcache,
                        __nv_bfloat16* vcache, int n_tok, int n_q_heads, int n_kv_heads, int head_dim, int ld,
                        int block_size, cudaStream_t st, const float* bias) {
  if (head_dim % 16 || ld % 8) return OPUS_ERR_ARG;
  if (bias != nullptr && partial == nullptr) return OPUS_ERR_ARG;   // a finished bf16 row already contains the bias
  if (n_tok == 0) return OPUS_OK;
  const long long total = (long long)n_tok * (n_q_heads + 2 * n_kv_heads) * (head_dim / 16);
  launch_pdl(n_tok <= 1024, rope_llama_kvappend_kernel, dim3(cdiv(total, 256)), dim3(256), 0, st, qkv, partial, n_partial, pos, slot, cos_t, sin_t, kcache,
                                                              vcache, n_tok, n_q_heads, n_kv_heads, head_dim, ld,
                                                              block_size, bias);
  return ok();
}

int final_ln_meanpool(const float* x, const __nv_bfloat16* delta, const int* cu_seqlens, const float* gamma,
                      const float* beta, float* pooled,
                      __nv_bfloat16* pooled_l2, float* hidden_out, int n_seqs, int dim, float eps, cudaStream_t st) {
  if (dim % 4 || dim > 128 * 10) return OPUS_ERR_ARG;
  if (n_seqs == 0) return OPUS_OK;
  const int threads = 256;
  const size_t smem = (size_t)(threads / 32) * dim * sizeof(float);
  launch_pdl(false, final_ln_meanpool_kernel<10>, dim3(n_seqs), dim3(threads), smem, st, x, delta, cu_seqlens, gamma, beta, pooled, pooled_l2, hidden_out,
                                                             dim, eps);
  return ok();
}

int l2norm_f32_bf16(const float* x, __nv_bfloat16* y, int rows, int dim, cudaStream_t st) {
  if (rows == 0) return OPUS_OK;
  launch_pdl(rows <= 1024, l2norm_f32_bf16_kernel, dim3(cdiv(rows, WARPS_PER_BLOCK)), dim3(WARPS_PER_BLOCK * 32), 0, st, x, y, rows, dim);
  return ok();
}

int splice_gather(const int* src, const __nv_bfloat16* embed, const __nv_bfloat16* soft, __nv_bfloat16* out, int n_rows,
                  int dim, cudaStream_t st) {
  if (dim % 8) return OPUS_ERR_ARG;
  if (n_rows == 0) return OPUS_OK;
  launch_pdl(n_rows <= 1024, splice_gather_kernel, dim3(cdiv(n_rows, WARPS_PER_BLOCK)), dim3(WARPS_PER_BLOCK * 32), 0, st, src, embed, soft, out, n_rows, dim);
  return ok();
}

int argmax_eos(const __nv_bfloat16* logits, int ld, int vocab, int n_rows, int* finished, const int* eos_ids, int n_eos,
               int pad_id, int* next_tok, int* out_ids, int out_ld, int step, int* n_unfinished, cudaStream_t st,
               const int* step_ptr) {
  if (ld % 8) return OPUS_ERR_ARG;
  if (n_rows == 0) return OPUS_OK;
  launch_pdl(true, argmax_eos_kernel, dim3(n_rows), dim3(1024), 0, st, logits, ld, vocab, finished, eos_ids, n_eos, pad_id, next_tok, out_ids,
                                            out_ld, step, step_ptr, n_unfinished);
  return ok();
}

int sample_top_p(const __nv_bfloat16* logits, int ld, int vocab, int n_rows, float temperature, float top_p,
                 unsigned long long seed, int* finished, const int* eos_ids, int n_eos, int pad_id, int* next_tok,
                 int* out_ids, int out_ld, int step, int* n_unfinished, cudaStream_t st, const int* step_ptr,
                 int* kept_count, const unsigned long long* seed_ptr) {
  if (n_rows == 0) return OPUS_OK;
  if (!(temperature > 0.f) || !(top_p > 0.f) || vocab <= 0) return OPUS_ERR_ARG;
  launch_pdl(true, sample_top_p_kernel, dim3(n_rows), dim3(SAMPLE_THREADS), 0, st, logits, ld, vocab, 1.0f / temperature,
             top_p, seed, finished, eos_ids, n_eos, pad_id, next_tok, out_ids, out_ld, step, step_ptr, n_unfinished,
             kept_count, seed_ptr);
  note_launch();
  return cudaGetLastError() == cudaSuccess ? OPUS_OK : OPUS_ERR_CUDA;
}

int cross_entropy_rows(const __nv_bfloat16* logits, int ld, int vocab, const int* target, float* loss, int n_rows,
                       cudaStream_t st) {
  if (n_rows == 0) return OPUS_OK;
  if (vocab <= 0 || (ld % 8) || (reinterpret_cast<uintptr_t>(logits) & 15)) return OPUS_ERR_ARG;
  cross_entropy_rows_kernel<<<n_rows, 512, 0, st>>>(logits, ld, vocab, target, loss);
  note_launch();
  return cudaGetLastError() == cudaSuccess ? OPUS_OK : OPUS_ERR_CUDA;
}

int embed_gather(const int* tok, const __nv_bfloat16* table, __nv_bfloat16* x, int n_rows, int dim, cudaStream_t st) {
  if (dim % 8) return OPUS_ERR_ARG;
  if (n_rows == 0) return OPUS_OK;
  launch_pdl(n_rows <= 1024, embed_gather_kernel, dim3(cdiv(n_rows, WARPS_PER_BLOCK)), dim3(WARPS_PER_BLOCK * 32), 0, st, tok, table, x, n_rows, dim);
  return ok();
}

int embed_gather_sumsq(const int* tok, const __nv_bfloat16* table, __nv_bfloat16* x, float* sumsq, int sumsq_ld,
                       int n_rows, int dim, cudaStream_t st) {
  if (dim % 32 || sumsq_ld < n_rows) return OPUS_ERR_ARG;
  if (n_rows == 0) return OPUS_OK;
  launch_pdl(true, embed_gather_sumsq_kernel, dim3(n_rows), dim3(128), 0, st, tok, table, x, sumsq, sumsq_ld, dim);
  return ok();
}

int decode_advance(int* ctx_len, int* pos, int* slot, const int* block_table, int max_blocks, int block_size, int n,
                   cudaStream_t st, int* step) {
  if (n == 0) return OPUS_OK;
  launch_pdl(true, decode_advance_kernel, dim3(cdiv(n, 128)), dim3(128), 0, st, ctx_len, pos, slot, block_table, max_blocks, block_size, n, step);
  return ok();
}

int lora_merge(__nv_bfloat16* W, const __nv_bfloat16* A, const __nv_bfloat16* B, int out_f, int in_f, int r, float scale,
               cudaStream_t st) {
  const long long total = (long long)out_f * in_f;
  if (total == 0) return OPUS_OK;
  launch_pdl(false, lora_merge_kernel, dim3(cdiv(total, 256)), dim3(256), 0, st, W, A, B, out_f, in_f, r, scale);
  return ok();
}

}  // namespace opus
