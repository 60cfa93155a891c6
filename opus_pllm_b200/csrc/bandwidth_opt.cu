// HBM-bound kernels that only the OPT / Galactica sibling family needs (language_model/opus_opt.py -> HF OPTDecoder):
//   * LayerNorm over a bf16 residual stream (nn.LayerNorm on bf16: fp32 statistics, one rounding at the end), with the
//     same optional fusions as rmsnorm_bf16 (residual add written back, split-K partial reduction) plus the bias of the
//     linear whose partial sums are being reduced (out_proj / fc2 carry biases in OPT);
//   * the learned positional embedding add  h[i,:] = bf16(h[i,:] + P[pos[i] + 2, :])  (OPTLearnedPositionalEmbedding).
// Also here: the per-row stop-sequence check of the decode loop (mm_utils.py KeywordsStoppingCriteria on the device).
// Coalesced 128-bit accesses, the row kept in registers, warp-shuffle + shared-memory reductions.
#include "common.h"
#include "kernels.h"
#include "launch.cuh"
#include "ptx.cuh"

namespace opus {

namespace {

__device__ __forceinline__ float4 ldf4(const float* p) { return *reinterpret_cast<const float4*>(p); }

__device__ __forceinline__ void unpack8(const uint4& q, float* f) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  uint4 q;
  q.x = pack_bf16x2(f[0], f[1]);
  q.y = pack_bf16x2(f[2], f[3]);
  q.z = pack_bf16x2(f[4], f[5]);
  q.w = pack_bf16x2(f[6], f[7]);
  return q;
}

// TPR threads per row, RPC rows per CTA, MAXC 8-element chunks per thread (cols <= TPR * 8 * MAXC).
//   h = x                                   (x != nullptr)
//     | bf16(sum_s partial[s] + red_bias)   (partial != nullptr; the value the un-split linear would have stored)
//   h = bf16(h + residual)                  (residual != nullptr)          -> h_out (nullable, may alias residual)
//   y = bf16((h - mean) * rstd * gamma + beta)                             (y nullable: reduce + residual only)
template <int TPR, int RPC, int MAXC>
__global__ void __launch_bounds__(TPR* RPC)
layernorm_bf16_kernel(const __nv_bfloat16* x, const float* __restrict__ partial, int n_partial,
                      const float* __restrict__ red_bias, const __nv_bfloat16* residual, __nv_bfloat16* h_out,
                      const float* __restrict__ gamma, const float* __restrict__ beta, __nv_bfloat16* y, int rows,
                      int cols, float eps) {
  grid_dep_launch();
  grid_dep_wait();
  constexpr int WPR = TPR / 32;
  __shared__ float red[2][RPC][WPR];
  const int r_in = threadIdx.x / TPR, t = threadIdx.x % TPR;
  const int row = blockIdx.x * RPC + r_in;
  const bool row_ok = row < rows;
  uint4 raw[MAXC];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < MAXC; ++i) {
    const int c = (i * TPR + t) * 8;
    raw[i] = make_uint4(0, 0, 0, 0);
    if (row_ok && c < cols) {
      const size_t off = (size_t)row * cols + c;
      float v[8];
      if (partial != nullptr) {
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int s0 = 0; s0 < n_partial; s0 += 4) {   // four slices in flight per round; adds keep the slice order
          float4 a[4], b[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const bool ok = s0 + u < n_partial;
            const float* pp = partial + (size_t)(ok ? s0 + u : s0) * rows * cols + off;
            a[u] = ldf4(pp); b[u] = ldf4(pp + 4);
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            if (s0 + u < n_partial) {
              acc[0] += a[u].x; acc[1] += a[u].y; acc[2] += a[u].z; acc[3] += a[u].w;
              acc[4] += b[u].x; acc[5] += b[u].y; acc[6] += b[u].z; acc[7] += b[u].w;
            }
          }
        }
        if (red_bias != nullptr) {
          const float4 b0 = ldf4(red_bias + c), b1 = ldf4(red_bias + c + 4);
          acc[0] += b0.x; acc[1] += b0.y; acc[2] += b0.z; acc[3] += b0.w;
          acc[4] += b1.x; acc[5] += b1.y; acc[6] += b1.z; acc[7] += b1.w;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = bf16_round(acc[j]);
      } else {
        unpack8(*reinterpret_cast<const uint4*>(x + off), v);
      }
      if (residual != nullptr) {
        float r[8];
        unpack8(*reinterpret_cast<const uint4*>(residual + off), r);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = bf16_round(v[j] + r[j]);
      }
      raw[i] = pack8(v);
      if (h_out != nullptr) *reinterpret_cast<uint4*>(h_out + off) = raw[i];
#pragma unroll
      for (int j = 0; j < 8; ++j) sum += v[j];
    }
  }
  if (y == nullptr) return;
  sum = warp_sum(sum);
  if ((threadIdx.x & 31) == 0) red[0][r_in][t >> 5] = sum;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int i = 0; i < WPR; ++i) tot += red[0][r_in][i];
  const float mean = tot / cols;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < MAXC; ++i) {
    const int c = (i * TPR + t) * 8;
    if (row_ok && c < cols) {
      float v[8];
      unpack8(raw[i], v);
#pragma unroll
      for (int j = 0; j < 8; ++j) sq += (v[j] - mean) * (v[j] - mean);
    }
  }
  sq = warp_sum(sq);
  if ((threadIdx.x & 31) == 0) red[1][r_in][t >> 5] = sq;
  __syncthreads();
  float tot2 = 0.f;
#pragma unroll
  for (int i = 0; i < WPR; ++i) tot2 += red[1][r_in][i];
  const float rstd = rsqrtf(tot2 / cols + eps);
#pragma unroll
  for (int i = 0; i < MAXC; ++i) {
    const int c = (i * TPR + t) * 8;
    if (row_ok && c < cols) {
      float v[8], o[8];
      unpack8(raw[i], v);
      const float4 g0 = ldf4(gamma + c), g1 = ldf4(gamma + c + 4);
      float4 b0 = make_float4(0.f, 0.f, 0.f, 0.f), b1 = b0;
      if (beta != nullptr) { b0 = ldf4(beta + c); b1 = ldf4(beta + c + 4); }
      o[0] = (v[0] - mean) * rstd * g0.x + b0.x; o[1] = (v[1] - mean) * rstd * g0.y + b0.y;
      o[2] = (v[2] - mean) * rstd * g0.z + b0.z; o[3] = (v[3] - mean) * rstd * g0.w + b0.w;
      o[4] = (v[4] - mean) * rstd * g1.x + b1.x; o[5] = (v[5] - mean) * rstd * g1.y + b1.y;
      o[6] = (v[6] - mean) * rstd * g1.z + b1.z; o[7] = (v[7] - mean) * rstd * g1.w + b1.w;
      *reinterpret_cast<uint4*>(y + (size_t)row * cols + c) = pack8(o);
    }
  }
}

// h[i,:] = bf16(h[i,:] + table[pos[i] + offset, :]); one warp per row
__global__ void add_pos_embed_kernel(__nv_bfloat16* __restrict__ h, const __nv_bfloat16* __restrict__ table,
                                     const int* __restrict__ pos, int offset, int max_rows, int n_rows, int dim) {
  grid_dep_launch();
  grid_dep_wait();
  const int row = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (row >= n_rows) return;
  const int lane = threadIdx.x & 31;
  int p = pos[row] + offset;
  p = p < 0 ? 0 : (p >= max_rows ? max_rows - 1 : p);
  const uint4* tb = reinterpret_cast<const uint4*>(table + (size_t)p * dim);
  uint4* dst = reinterpret_cast<uint4*>(h + (size_t)row * dim);
  for (int c = lane; c < dim / 8; c += 32) {
    float a[8], b[8];
    unpack8(dst[c], a);
    unpack8(tb[c], b);
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] += b[j];
    dst[c] = pack8(a);
  }
}

// one thread per row: does the emitted tail out_ids[b, step-len+1 .. step] equal one of the stop sequences?
__global__ void stop_sequences_kernel(const int* __restrict__ out_ids, int out_ld, int n_rows, int step_imm,
                                      const int* __restrict__ step_ptr, const int* __restrict__ stop_seqs,
                                      const int* __restrict__ stop_lens, int n_stop, int stop_ld,
                                      int* __restrict__ finished, int* __restrict__ n_unfinished) {
  grid_dep_launch();
  grid_dep_wait();
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n_rows || finished[b]) return;
  const int step = step_ptr != nullptr ? *step_ptr : step_imm;
  const int* row = out_ids + (size_t)b * out_ld;
  for (int j = 0; j < n_stop; ++j) {
    const int len = stop_lens[j];
    if (len <= 0 || len > step + 1) continue;
    bool same = true;
    for (int i = 0; i < len && same; ++i) same = row[step - len + 1 + i] == stop_seqs[(size_t)j * stop_ld + i];
    if (same) {
      finished[b] = 1;
      if (n_unfinished != nullptr) atomicSub(n_unfinished, 1);
      return;
    }
  }
}

inline int cdiv_(long long a, long long b) { return (int)((a + b - 1) / b); }
inline int ok_() {
  note_launch();
  return cudaGetLastError() == cudaSuccess ? OPUS_OK : OPUS_ERR_CUDA;
}

}  // namespace

int layernorm_bf16(const __nv_bfloat16* x, const float* partial, int n_partial, const float* red_bias,
                   const __nv_bfloat16* residual, __nv_bfloat16* h_out, const float* gamma, const float* beta,
                   __nv_bfloat16* y, int rows, int cols, float eps, cudaStream_t st) {
  if (cols % 8 || cols > 8192) return OPUS_ERR_ARG;
  if ((x == nullptr) == (partial == nullptr)) return OPUS_ERR_ARG;
  if (y != nullptr && gamma == nullptr) return OPUS_ERR_ARG;
  if (rows == 0) return OPUS_OK;
  if (rows <= 1024) {  // decode-sized: one 512-thread CTA per row, every load of the row in flight at once
    if (cols <= 4096)
      launch_pdl(true, layernorm_bf16_kernel<512, 1, 1>, dim3(rows), dim3(512), 0, st, x, partial, n_partial, red_bias, residual, h_out, gamma, beta, y, rows, cols, eps);
    else
      launch_pdl(true, layernorm_bf16_kernel<512, 1, 2>, dim3(rows), dim3(512), 0, st, x, partial, n_partial, red_bias, residual, h_out, gamma, beta, y, rows, cols, eps);
  } else {             // prefill-sized: 128 threads per row, two rows per CTA
    const int grid = cdiv_(rows, 2);
    if (cols <= 4096)
      launch_pdl(false, layernorm_bf16_kernel<128, 2, 4>, dim3(grid), dim3(256), 0, st, x, partial, n_partial, red_bias, residual, h_out, gamma, beta, y, rows, cols, eps);
    else
      launch_pdl(false, layernorm_bf16_kernel<128, 2, 8>, dim3(grid), dim3(256), 0, st, x, partial, n_partial, red_bias, residual, h_out, gamma, beta, y, rows, cols, eps);
  }
  return ok_();
}

int add_pos_embed(__nv_bfloat16* h, const __nv_bfloat16* table, const int* pos, int offset, int table_rows, int n_rows,
                  int dim, cudaStream_t st) {
  if (dim % 8 || table_rows <= 0) return OPUS_ERR_ARG;
  if (n_rows == 0) return OPUS_OK;
  launch_pdl(n_rows <= 1024, add_pos_embed_kernel, dim3(cdiv_(n_rows, 4)), dim3(128), 0, st, h, table, pos, offset, table_rows, n_rows, dim);
  return ok_();
}

int stop_sequences(const int* out_ids, int out_ld, int n_rows, int step, const int* step_ptr, const int* stop_seqs,
                   const int* stop_lens, int n_stop, int stop_ld, int* finished, int* n_unfinished, cudaStream_t st) {
  if (n_rows == 0 || n_stop <= 0 || stop_seqs == nullptr) return OPUS_OK;
  if (stop_lens == nullptr || stop_ld <= 0 || finished == nullptr) return OPUS_ERR_ARG;
  launch_pdl(true, stop_sequences_kernel, dim3(cdiv_(n_rows, 128)), dim3(128), 0, st, out_ids, out_ld, n_rows, step,
             step_ptr, stop_seqs, stop_lens, n_stop, stop_ld, finished, n_unfinished);
  return ok_();
}

}  // namespace opus
