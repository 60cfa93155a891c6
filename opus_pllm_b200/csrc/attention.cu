// Attention kernels of the generation path (flash-style, online softmax, nothing T x T is ever materialised):
//   * attn_varlen: packed variable-length sequences (cu_seqlens), bidirectional (ESM-2 encoder, hd = 64) or causal with
//     grouped-query heads (Llama-3 prefill, hd = 128). Padded batches are the same kernel: the host passes the valid
//     spans, so pad rows cost no FLOPs and can never be attended to (key-padding mask of fair-esm / HF's 4-D mask).
//   * attn_decode_paged: one new token per sequence against the paged KV cache, the 4 query heads of a KV group share
//     each K/V panel, warps split the context and merge their (max, sum, acc) triples through shared memory.
// Round-1 implementation uses mma.sync m16n8k16 bf16 tiles fed by cp.async + ldmatrix; both kernels are HBM/L2 friendly
// (every K/V byte is read once per (sequence, kv head) q-tile). A tcgen05/TMEM variant is the planned successor for the
// long-protein encoder case where attention FLOPs stop being negligible.
//
// Replaces: fair-esm MultiheadAttention bmm/softmax/bmm (reached from cstp_v3/modelling.py:48) and HF Llama
// sdpa/eager attention + DynamicCache (reached from language_model/opus_llama.py:127-132).
#include "common.h"
#include "kernels.h"
#include "launch.cuh"
#include "ptx.cuh"

#include <cstdlib>
#include <mutex>

namespace opus {

namespace {

// ---------------------------------------------------------------------------------------------- primitives
__device__ __forceinline__ void cp_async16(uint32_t smem_dst, const void* gsrc, bool valid) {
  const int sz = valid ? 16 : 0;  // src-size 0 => zero-fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_dst), "l"(gsrc), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// Tile of ROWS x D bf16 in shared memory, stored as D/64 panels of [ROWS][64] with the 16-byte chunks of every
// 128-byte row XOR-swizzled by (row & 7): conflict-free for both cp.async stores and ldmatrix loads.
template <int ROWS>
__device__ __forceinline__ uint32_t tile_off(int row, int chunk) {
  return (uint32_t)((chunk >> 3) * (ROWS * 128) + row * 128 + (((chunk & 7) ^ (row & 7)) << 4));
}

// ---------------------------------------------------------------------------------------------- varlen / prefill
constexpr int BN = 64;   // keys per pipeline step
// query rows per CTA = BM (16 per warp): 128 for long sequences, 64 when that wastes fewer rows (e.g. T = 258)

struct AttnParams {
  const __nv_bfloat16* q;
  const __nv_bfloat16* k;
  const __nv_bfloat16* v;
  __nv_bfloat16* o;
  const int* cu_seqlens;
  int ldq, ldk, ldv, ldo;  // row strides (elements)
  int n_q_heads, group;    // group = q heads per kv head
  float scale_log2;        // softmax scale * log2(e)
  int tail_only;           // > 0 (grid.x = 1): only the LAST 256-row-aligned query block of a sequence, and only when it
                           // holds <= tail_only rows (the rows the tcgen05 kernel was told to skip)
};

template <int D, int ROWS, int THREADS>
__device__ __forceinline__ void load_tile_async(uint32_t smem_base, const __nv_bfloat16* g, int ld, int row0, int len,
                                                int tid) {
  constexpr int CH = D / 8;  // 16-byte chunks per row
  for (int i = tid; i < ROWS * CH; i += THREADS) {
    const int r = i / CH, c = i - r * CH;
    const int gr = row0 + r;
    const bool valid = gr < len;
    const __nv_bfloat16* src = g + (size_t)(valid ? gr : 0) * ld + c * 8;
    cp_async16(smem_base + tile_off<ROWS>(r, c), src, valid);
  }
}

template <int D, bool CAUSAL, int BM>
__global__ void __launch_bounds__(BM * 2, (D == 64) ? (BM == 64 ? 3 : 2) : 1) attn_varlen_kernel(const AttnParams p) {
  constexpr int ATT_THREADS = BM * 2;
  grid_dep_launch();
  grid_dep_wait();
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr int Q_BYTES = BM * D * 2, KV_BYTES = BN * D * 2;
  const uint32_t sQ = smem_u32(smem);
  const uint32_t sK = sQ + Q_BYTES;             // 2 stages
  const uint32_t sV = sK + 2 * KV_BYTES;        // 2 stages

  const int b = blockIdx.z, h = blockIdx.y;
  const int seq_start = p.cu_seqlens[b];
  const int len = p.cu_seqlens[b + 1] - seq_start;
  const int mblk = CAUSAL ? (gridDim.x - 1 - blockIdx.x) : blockIdx.x;  // heavy causal tiles first
  int q0 = mblk * BM;
  if (p.tail_only > 0) {
    if (len <= 0) return;
    q0 = ((len - 1) / 256) * 256;
    if (len - q0 > p.tail_only) return;
  }
  if (q0 >= len) return;
  const int kvh = h / p.group;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t4 = lane & 3;

  const __nv_bfloat16* gq = p.q + (size_t)seq_start * p.ldq + (size_t)h * D;
  const __nv_bfloat16* gk = p.k + (size_t)seq_start * p.ldk + (size_t)kvh * D;
  const __nv_bfloat16* gv = p.v + (size_t)seq_start * p.ldv + (size_t)kvh * D;

  const int kv_end = CAUSAL ? min(len, q0 + BM) : len;
  const int n_blocks = (kv_end + BN - 1) / BN;

  load_tile_async<D, BM, ATT_THREADS>(sQ, gq, p.ldq, q0, len, tid);
  load_tile_async<D, BN, ATT_THREADS>(sK, gk, p.ldk, 0, len, tid);
  load_tile_async<D, BN, ATT_THREADS>(sV, gv, p.ldv, 0, len, tid);
  cp_async_commit();

  uint32_t qf[D / 16][4];
  float o_acc[D / 8][4];
#pragma unroll
  for (int i = 0; i < D / 8; ++i) { o_acc[i][0] = o_acc[i][1] = o_acc[i][2] = o_acc[i][3] = 0.f; }
  float m_run[2] = {-INFINITY, -INFINITY};
  float l_run[2] = {0.f, 0.f};

  const int qrow_lo = q0 + warp * 16 + g;  // sequence-relative query index of this thread's first row (second: +8)

  for (int j = 0; j < n_blocks; ++j) {
    const int st = j & 1;
    if (j + 1 < n_blocks) {
      load_tile_async<D, BN, ATT_THREADS>(sK + (st ^ 1) * KV_BYTES, gk, p.ldk, (j + 1) * BN, len, tid);
      load_tile_async<D, BN, ATT_THREADS>(sV + (st ^ 1) * KV_BYTES, gv, p.ldv, (j + 1) * BN, len, tid);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();

    if (j == 0) {
#pragma unroll
      for (int kk = 0; kk < D / 16; ++kk) {
        const int row = warp * 16 + (lane & 15);
        const int chunk = kk * 2 + (lane >> 4);
        ldsm_x4(sQ + tile_off<BM>(row, chunk), qf[kk][0], qf[kk][1], qf[kk][2], qf[kk][3]);
      }
    }

    // warps whose 16 query rows are all above this key block (causal) or beyond the sequence have nothing to do
    const int k0 = j * BN;
    const bool warp_active = (q0 + warp * 16 < len) && (!CAUSAL || k0 <= q0 + warp * 16 + 15);
    if (warp_active) {
      float s[BN / 8][4];
#pragma unroll
      for (int i = 0; i < BN / 8; ++i) { s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f; }
      const uint32_t kb = sK + st * KV_BYTES;
#pragma unroll
      for (int kk = 0; kk < D / 16; ++kk) {
#pragma unroll
        for (int nb2 = 0; nb2 < BN / 16; ++nb2) {
          uint32_t b0, b1, b2, b3;
          const int n = nb2 * 16 + (lane & 7) + ((lane >> 4) << 3);
          const int chunk = kk * 2 + ((lane >> 3) & 1);
          ldsm_x4(kb + tile_off<BN>(n, chunk), b0, b1, b2, b3);
          mma_bf16_16816(s[2 * nb2], qf[kk], b0, b1);
          mma_bf16_16816(s[2 * nb2 + 1], qf[kk], b2, b3);
        }
      }
      // mask + running max
      float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
      for (int nb = 0; nb < BN / 8; ++nb) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int key = k0 + nb * 8 + t4 * 2 + (e & 1);
          const int qr = qrow_lo + ((e >> 1) << 3);
          const bool ok = key < len && (!CAUSAL || key <= qr);
          const float val = ok ? s[nb][e] * p.scale_log2 : -INFINITY;
          s[nb][e] = val;
          mx[e >> 1] = fmaxf(mx[e >> 1], val);
        }
      }
      float corr[2], m_use[2];
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
        mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
        const float m_new = fmaxf(m_run[r], mx[r]);
        m_use[r] = (m_new == -INFINITY) ? 0.f : m_new;  // fully masked row so far: avoid (-inf) - (-inf)
        corr[r] = exp2f(m_run[r] - m_use[r]);           // m_run = -inf -> 0
        m_run[r] = m_new;
        l_run[r] *= corr[r];
      }
      float ls[2] = {0.f, 0.f};
#pragma unroll
      for (int nb = 0; nb < BN / 8; ++nb) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float pv = exp2f(s[nb][e] - m_use[e >> 1]);
          s[nb][e] = pv;
          ls[e >> 1] += pv;
        }
      }
      l_run[0] += ls[0];
      l_run[1] += ls[1];
#pragma unroll
      for (int i = 0; i < D / 8; ++i) {
        o_acc[i][0] *= corr[0]; o_acc[i][1] *= corr[0];
        o_acc[i][2] *= corr[1]; o_acc[i][3] *= corr[1];
      }
      const uint32_t vb = sV + st * KV_BYTES;
#pragma unroll
      for (int kk = 0; kk < BN / 16; ++kk) {
        uint32_t a[4];
        a[0] = pack_bf16x2(s[2 * kk][0], s[2 * kk][1]);
        a[1] = pack_bf16x2(s[2 * kk][2], s[2 * kk][3]);
        a[2] = pack_bf16x2(s[2 * kk + 1][0], s[2 * kk + 1][1]);
        a[3] = pack_bf16x2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
        for (int db2 = 0; db2 < D / 16; ++db2) {
          uint32_t b0, b1, b2, b3;
          const int key = kk * 16 + (lane & 7) + (((lane >> 3) & 1) << 3);
          const int chunk = db2 * 2 + (lane >> 4);
          ldsm_x4_t(vb + tile_off<BN>(key, chunk), b0, b1, b2, b3);
          mma_bf16_16816(o_acc[2 * db2], a, b0, b1);
          mma_bf16_16816(o_acc[2 * db2 + 1], a, b2, b3);
        }
      }
    }
    __syncthreads();  // everyone done with stage `st` before it is refilled
  }

  // finalise: divide by the row sums (quad-reduced) and store bf16
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int qr = qrow_lo + r * 8;
    if (qr < len) {
      const float inv = l_run[r] > 0.f ? 1.0f / l_run[r] : 0.f;
      __nv_bfloat16* dst = p.o + (size_t)(seq_start + qr) * p.ldo + (size_t)h * D;
#pragma unroll
      for (int i = 0; i < D / 8; ++i) {
        const uint32_t w = pack_bf16x2(o_acc[i][2 * r] * inv, o_acc[i][2 * r + 1] * inv);
        *reinterpret_cast<uint32_t*>(dst + i * 8 + t4 * 2) = w;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------- paged decode
#ifndef OPUS_DEC_WARPS
#define OPUS_DEC_WARPS 2   // 2 warps x 36 KB: all 512 (kv head, sequence) CTAs of a batch-64 step are resident at once (4.41 -> 4.31 ms per step)
#endif
constexpr int DEC_WARPS = OPUS_DEC_WARPS;
constexpr int DEC_BS = 16;   // tokens per KV block (cache page)
constexpr int DEC_D = 128;   // head dim
constexpr int DEC_PANEL = DEC_BS * DEC_D * 2;  // 4 KB

__device__ __forceinline__ float4 ldf4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void unpack_bf16x8(const uint4& q, float* f) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 t = __bfloat1622float2(h[j]);
    f[2 * j] = t.x;
    f[2 * j + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack_bf16x8(const float* f) {
  uint4 q;
  q.x = pack_bf16x2(f[0], f[1]); q.y = pack_bf16x2(f[2], f[3]);
  q.z = pack_bf16x2(f[4], f[5]); q.w = pack_bf16x2(f[6], f[7]);
  return q;
}

struct DecodeParams {
  const __nv_bfloat16* q;  // [B, ldq], head h at column h*128 (already rotated)
  int ldq;
  const __nv_bfloat16* kcache;  // [blocks][n_kv_heads][16][128]
  const __nv_bfloat16* vcache;
  const int* block_table;  // [B, max_blocks]
  int max_blocks;
  const int* ctx_len;  // [B] number of valid cached tokens (including the one just appended)
  __nv_bfloat16* o;    // [B, ldo]
  int ldo;
  int n_kv_heads, group;
  float scale_log2;
  // Optional fused prologue (decode): the q|k|v row of this step still sits in fp32 split-K partials. The CTA of
  // (sequence, kv head) reduces ITS heads (GROUP query heads, one key head, one value head), applies RoPE with the
  // rounding points of rope_llama_kvappend_kernel, writes q/k/v to `qkv` and k/v to the paged cache, then attends.
  const float* partial;  // [n_partial][B][ldq] or nullptr (q / cache already final)
  int n_partial;
  __nv_bfloat16* qkv;    // == q base; written when partial != nullptr
  const int* pos;
  const int* slot;
  const __nv_bfloat16* cos_t;  // bf16 [max_pos, 128], cat(freqs, freqs)
  const __nv_bfloat16* sin_t;
  __nv_bfloat16* kcache_w;
  __nv_bfloat16* vcache_w;
  const float* bias;           // fp32 [ldq] q|k|v projection bias (Qwen2) or nullptr
  // Split-KV: the context of one (sequence, kv head) is cut into n_split page ranges handled by n_split CTAs (gridDim.z),
  // so the work units are fine enough to balance over the SMs (batch 64: 512 units on 148 SMs leave 68 SMs with four
  // units and 80 with three; 1024 half-units are handed out as CTA slots free up). Each CTA writes its un-normalised
  // (max, sum, acc) to split_ws; the CTA that arrives LAST at split_cnt[unit] merges all parts in part order.
  int n_split;
  float* split_ws;             // [units][n_split][GROUP][DEC_D + 2] fp32
  int* split_cnt;              // [units], zero before the launch; the merging CTA re-arms its entry
};

__device__ __forceinline__ void load_panel_async(uint32_t smem_base, const __nv_bfloat16* panel, int lane) {
  // 4 KB contiguous panel = 256 chunks of 16 B; row = chunk / 16 (256-byte rows), c = chunk % 16
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int id = i * 32 + lane;
    cp_async16(smem_base + tile_off<DEC_BS>(id >> 4, id & 15), reinterpret_cast<const uint8_t*>(panel) + id * 16, true);
  }
}

template <int GROUP>
__global__ void __launch_bounds__(DEC_WARPS * 32) attn_decode_paged_kernel(const DecodeParams p) {
  grid_dep_launch();
  extern __shared__ __align__(1024) uint8_t smem[];
  // per warp: 2 stages x (K panel + V panel) = 16 KB; then the merge area
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const int kvh = blockIdx.x, b = blockIdx.y;
  // ctx_len / block_table / pos / slot of this step were written by decode_advance, several kernels back: every kernel
  // in front of the q|k|v GEMM has completed by the time this grid can start (the GEMM triggers its dependents only
  // after its own griddepcontrol.wait), so they may be read before OUR wait; only the GEMM's output may not.
  const int ctx = p.ctx_len[b];
  const int n_blocks = (ctx + DEC_BS - 1) / DEC_BS;
  const int part = blockIdx.z, per_part = (n_blocks + p.n_split - 1) / p.n_split;
  const int blk_begin = min(n_blocks, part * per_part), blk_end = min(n_blocks, blk_begin + per_part);
  const bool owns_last = n_blocks == 0 ? part == 0 : (blk_begin <= n_blocks - 1 && n_blocks - 1 < blk_end);
  const uint32_t s_warp = smem_u32(smem) + warp * (4 * DEC_PANEL);
  float* s_merge = reinterpret_cast<float*>(smem + DEC_WARPS * 4 * DEC_PANEL);  // [warps][GROUP][128+2]
  const int* bt = p.block_table + (size_t)b * p.max_blocks;
  auto panel_ptr = [&](const __nv_bfloat16* cache, int blk_idx) {
    return cache + ((size_t)bt[blk_idx] * p.n_kv_heads + kvh) * (DEC_BS * DEC_D);
  };
  // The first K/V panels of this warp are cached tokens of EARLIER steps unless the block holds the token appended by
  // this step (the last block): request them before waiting for the preceding kernel, so the page fetch overlaps its
  // tail, the split-K reduce and the RoPE below.
  const bool early = p.partial != nullptr && blk_begin + warp < blk_end && blk_begin + warp < n_blocks - 1;
  if (early) {
    load_panel_async(s_warp, panel_ptr(p.kcache, blk_begin + warp), lane);
    load_panel_async(s_warp + DEC_PANEL, panel_ptr(p.vcache, blk_begin + warp), lane);
  }
  grid_dep_wait();

  if (p.partial != nullptr) {
    // ---- fused split-K reduce + RoPE + KV append for the heads of this (sequence, kv head) ----
    const int B = gridDim.y, Hq = p.n_kv_heads * GROUP;
    const int pos = p.pos[b], sl = p.slot[b];
    // every part needs the rotated query heads (identical values, written by all of them); the new key / value row is
    // reduced, rotated and appended by the part whose page range holds it
    for (int t = threadIdx.x; t < (owns_last ? GROUP + 2 : GROUP) * 8; t += DEC_WARPS * 32) {
      const int hh = t >> 3, j0 = (t & 7) * 8;          // head of this CTA, first of 8 rotation pairs
      const bool is_k = hh == GROUP, is_v = hh == GROUP + 1;
      const int hcol = (hh < GROUP ? kvh * GROUP + hh : (is_k ? Hq + kvh : Hq + p.n_kv_heads + kvh)) * DEC_D;
      float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, c[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      for (int s0 = 0; s0 < p.n_partial; s0 += 3) {     // three slices per round, loads first
        float4 t1[3], t2[3], t3[3], t4[3];
#pragma unroll
        for (int u = 0; u < 3; ++u) {
          const bool ok = s0 + u < p.n_partial;
          const float* pp = p.partial + ((size_t)(ok ? s0 + u : s0) * B + b) * p.ldq + hcol;
          t1[u] = ldf4(pp + j0); t2[u] = ldf4(pp + j0 + 4); t3[u] = ldf4(pp + 64 + j0); t4[u] = ldf4(pp + 64 + j0 + 4);
        }
#pragma unroll
        for (int u = 0; u < 3; ++u) {
          if (s0 + u < p.n_partial) {
            a[0] += t1[u].x; a[1] += t1[u].y; a[2] += t1[u].z; a[3] += t1[u].w;
            a[4] += t2[u].x; a[5] += t2[u].y; a[6] += t2[u].z; a[7] += t2[u].w;
            c[0] += t3[u].x; c[1] += t3[u].y; c[2] += t3[u].z; c[3] += t3[u].w;
            c[4] += t4[u].x; c[5] += t4[u].y; c[6] += t4[u].z; c[7] += t4[u].w;
          }
        }
      }
      if (p.bias != nullptr) {
        const float* bb = p.bias + hcol;
        const float4 b1 = ldf4(bb + j0), b2 = ldf4(bb + j0 + 4), b3 = ldf4(bb + 64 + j0), b4 = ldf4(bb + 64 + j0 + 4);
        a[0] += b1.x; a[1] += b1.y; a[2] += b1.z; a[3] += b1.w; a[4] += b2.x; a[5] += b2.y; a[6] += b2.z; a[7] += b2.w;
        c[0] += b3.x; c[1] += b3.y; c[2] += b3.z; c[3] += b3.w; c[4] += b4.x; c[5] += b4.y; c[6] += b4.z; c[7] += b4.w;
      }
      float x1[8], x2[8], o1[8], o2[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) { x1[i] = bf16_round(a[i]); x2[i] = bf16_round(c[i]); }
      if (!is_v) {
        float cs[8], sn[8];
        unpack_bf16x8(*reinterpret_cast<const uint4*>(p.cos_t + (size_t)pos * DEC_D + j0), cs);
        unpack_bf16x8(*reinterpret_cast<const uint4*>(p.sin_t + (size_t)pos * DEC_D + j0), sn);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          o1[i] = bf16_round(bf16_round(x1[i] * cs[i]) + bf16_round(-x2[i] * sn[i]));
          o2[i] = bf16_round(bf16_round(x2[i] * cs[i]) + bf16_round(x1[i] * sn[i]));
        }
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) { o1[i] = x1[i]; o2[i] = x2[i]; }
      }
      const uint4 w1 = pack_bf16x8(o1), w2 = pack_bf16x8(o2);
      __nv_bfloat16* dst = p.qkv + (size_t)b * p.ldq + hcol;
      *reinterpret_cast<uint4*>(dst + j0) = w1;
      *reinterpret_cast<uint4*>(dst + 64 + j0) = w2;
      if ((is_k || is_v) && sl >= 0) {
        const int blk = sl / DEC_BS, off = sl - blk * DEC_BS;
        __nv_bfloat16* cd = (is_v ? p.vcache_w : p.kcache_w) + (((size_t)blk * p.n_kv_heads + kvh) * DEC_BS + off) * DEC_D;
        *reinterpret_cast<uint4*>(cd + j0) = w1;
        *reinterpret_cast<uint4*>(cd + 64 + j0) = w2;
      }
    }
    __syncthreads();   // q for the fragments below and the new K/V row for the page loads are visible to this CTA
  }

  // Q fragments: rows 0..GROUP-1 of the m16 tile are the query heads of this KV group, the rest are zero
  uint32_t qf[DEC_D / 16][4];
  {
    const __nv_bfloat16* qrow = p.q + (size_t)b * p.ldq + (size_t)(kvh * GROUP + g) * DEC_D;
#pragma unroll
    for (int kk = 0; kk < DEC_D / 16; ++kk) {
      qf[kk][0] = (g < GROUP) ? *reinterpret_cast<const uint32_t*>(qrow + kk * 16 + t4 * 2) : 0u;
      qf[kk][1] = 0u;
      qf[kk][2] = (g < GROUP) ? *reinterpret_cast<const uint32_t*>(qrow + kk * 16 + 8 + t4 * 2) : 0u;
      qf[kk][3] = 0u;
    }
  }
  float o_acc[DEC_D / 8][4];
#pragma unroll
  for (int i = 0; i < DEC_D / 8; ++i) { o_acc[i][0] = o_acc[i][1] = o_acc[i][2] = o_acc[i][3] = 0.f; }
  float m_run = -INFINITY, l_run = 0.f;

  int it = 0;
  if (!early && blk_begin + warp < blk_end) {
    load_panel_async(s_warp, panel_ptr(p.kcache, blk_begin + warp), lane);
    load_panel_async(s_warp + DEC_PANEL, panel_ptr(p.vcache, blk_begin + warp), lane);
  }
  cp_async_commit();
  for (int blk = blk_begin + warp; blk < blk_end; blk += DEC_WARPS, ++it) {
    const int st = it & 1;
    const int nxt = blk + DEC_WARPS;
    if (nxt < blk_end) {
      load_panel_async(s_warp + (st ^ 1) * 2 * DEC_PANEL, panel_ptr(p.kcache, nxt), lane);
      load_panel_async(s_warp + (st ^ 1) * 2 * DEC_PANEL + DEC_PANEL, panel_ptr(p.vcache, nxt), lane);
    }
    cp_async_commit();
    cp_async_wait<1>();
    __syncwarp();
    const uint32_t kb = s_warp + st * 2 * DEC_PANEL, vb = kb + DEC_PANEL;

    float s[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
    for (int kk = 0; kk < DEC_D / 16; ++kk) {
      uint32_t b0, b1, b2, b3;
      const int n = (lane & 7) + ((lane >> 4) << 3);
      const int chunk = kk * 2 + ((lane >> 3) & 1);
      ldsm_x4(kb + tile_off<DEC_BS>(n, chunk), b0, b1, b2, b3);
      mma_bf16_16816(s[0], qf[kk], b0, b1);
      mma_bf16_16816(s[1], qf[kk], b2, b3);
    }
    // rows g (valid if g < GROUP); elements e=0,1 of each n-block
    float mx = -INFINITY;
#pragma unroll
    for (int nb = 0; nb < 2; ++nb) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int key = blk * DEC_BS + nb * 8 + t4 * 2 + e;
        const float val = key < ctx ? s[nb][e] * p.scale_log2 : -INFINITY;
        s[nb][e] = val;
        mx = fmaxf(mx, val);
      }
    }
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
    const float m_new = fmaxf(m_run, mx);
    const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
    const float corr = exp2f(m_run - m_use);
    m_run = m_new;
    float ls = 0.f;
#pragma unroll
    for (int nb = 0; nb < 2; ++nb) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const float pv = exp2f(s[nb][e] - m_use);
        s[nb][e] = pv;
        ls += pv;
      }
    }
    l_run = l_run * corr + ls;
#pragma unroll
    for (int i = 0; i < DEC_D / 8; ++i) { o_acc[i][0] *= corr; o_acc[i][1] *= corr; }
    uint32_t a[4];
    a[0] = pack_bf16x2(s[0][0], s[0][1]);
    a[1] = 0u;
    a[2] = pack_bf16x2(s[1][0], s[1][1]);
    a[3] = 0u;
#pragma unroll
    for (int db2 = 0; db2 < DEC_D / 16; ++db2) {
      uint32_t b0, b1, b2, b3;
      const int key = (lane & 7) + (((lane >> 3) & 1) << 3);
      const int chunk = db2 * 2 + (lane >> 4);
      ldsm_x4_t(vb + tile_off<DEC_BS>(key, chunk), b0, b1, b2, b3);
      mma_bf16_16816(o_acc[2 * db2], a, b0, b1);
      mma_bf16_16816(o_acc[2 * db2 + 1], a, b2, b3);
    }
    __syncwarp();
  }
  cp_async_wait<0>();

  // merge the per-warp partial softmax states
  l_run += __shfl_xor_sync(0xffffffffu, l_run, 1);
  l_run += __shfl_xor_sync(0xffffffffu, l_run, 2);
  constexpr int MS = DEC_D + 2;
  if (g < GROUP) {
    float* dst = s_merge + ((size_t)warp * GROUP + g) * MS;
#pragma unroll
    for (int i = 0; i < DEC_D / 8; ++i) {
      dst[i * 8 + t4 * 2] = o_acc[i][0];
      dst[i * 8 + t4 * 2 + 1] = o_acc[i][1];
    }
    if (t4 == 0) { dst[DEC_D] = m_run; dst[DEC_D + 1] = l_run; }
  }
  __syncthreads();
  if (p.n_split <= 1) {
    for (int idx = threadIdx.x; idx < GROUP * DEC_D; idx += DEC_WARPS * 32) {
      const int hh = idx / DEC_D, c = idx - hh * DEC_D;
      float mmax = -INFINITY;
#pragma unroll
      for (int w = 0; w < DEC_WARPS; ++w) mmax = fmaxf(mmax, s_merge[((size_t)w * GROUP + hh) * MS + DEC_D]);
      float num = 0.f, den = 0.f;
#pragma unroll
      for (int w = 0; w < DEC_WARPS; ++w) {
        const float* src = s_merge + ((size_t)w * GROUP + hh) * MS;
        const float mw = src[DEC_D];
        const float sc = (mw == -INFINITY) ? 0.f : exp2f(mw - mmax);
        num += src[c] * sc;
        den += src[DEC_D + 1] * sc;
      }
      p.o[(size_t)b * p.ldo + (size_t)(kvh * GROUP + hh) * DEC_D + c] = __float2bfloat16_rn(den > 0.f ? num / den : 0.f);
    }
    return;
  }
  // ---- split-KV: publish this part's (acc, max, sum) per head, the last part to arrive merges them in part order
  const int unit = b * p.n_kv_heads + kvh;
  float* mine = p.split_ws + ((size_t)unit * p.n_split + part) * GROUP * MS;
  for (int idx = threadIdx.x; idx < GROUP * MS; idx += DEC_WARPS * 32) {
    const int hh = idx / MS, c = idx - hh * MS;
    float mmax = -INFINITY;
#pragma unroll
    for (int w = 0; w < DEC_WARPS; ++w) mmax = fmaxf(mmax, s_merge[((size_t)w * GROUP + hh) * MS + DEC_D]);
    float v = 0.f;
    if (c == DEC_D) {
      v = mmax;
    } else {
#pragma unroll
      for (int w = 0; w < DEC_WARPS; ++w) {
        const float* src = s_merge + ((size_t)w * GROUP + hh) * MS;
        const float mw = src[DEC_D];
        v += src[c] * ((mw == -INFINITY) ? 0.f : exp2f(mw - mmax));   // c == DEC_D + 1: the sum, else the accumulator
      }
    }
    mine[idx] = v;
  }
  __shared__ int s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const int prev = atomicAdd(p.split_cnt + unit, 1);
    s_last = prev == p.n_split - 1;
    if (s_last) p.split_cnt[unit] = 0;       // re-armed for the next launch (stream order separates launches)
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  const float* parts = p.split_ws + (size_t)unit * p.n_split * GROUP * MS;
  for (int idx = threadIdx.x; idx < GROUP * DEC_D; idx += DEC_WARPS * 32) {
    const int hh = idx / DEC_D, c = idx - hh * DEC_D;
    float mmax = -INFINITY;
    for (int s2 = 0; s2 < p.n_split; ++s2) mmax = fmaxf(mmax, __ldcg(parts + ((size_t)s2 * GROUP + hh) * MS + DEC_D));
    float num = 0.f, den = 0.f;
    for (int s2 = 0; s2 < p.n_split; ++s2) {
      const float* src = parts + ((size_t)s2 * GROUP + hh) * MS;
      const float mw = __ldcg(src + DEC_D);
      const float sc = (mw == -INFINITY) ? 0.f : exp2f(mw - mmax);
      num += __ldcg(src + c) * sc;
      den += __ldcg(src + DEC_D + 1) * sc;
    }
    p.o[(size_t)b * p.ldo + (size_t)(kvh * GROUP + hh) * DEC_D + c] = __float2bfloat16_rn(den > 0.f ? num / den : 0.f);
  }
}

template <int D, bool CAUSAL, int BM>
int launch_varlen(const AttnParams& p, int n_seqs, int max_len, cudaStream_t st, bool tail = false) {
  constexpr int SMEM = BM * D * 2 + 4 * BN * D * 2;
  constexpr int ATT_THREADS = BM * 2;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(attn_varlen_kernel<D, CAUSAL, BM>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM) !=
        cudaSuccess)
      return OPUS_ERR_CUDA;
    configured = true;
  }
  dim3 grid(tail ? 1 : (max_len + BM - 1) / BM, p.n_q_heads, n_seqs);
  launch_pdl(false, attn_varlen_kernel<D, CAUSAL, BM>, dim3(grid), dim3(ATT_THREADS), SMEM, st, p);
  note_launch();
  return cudaGetLastError() == cudaSuccess ? OPUS_OK : OPUS_ERR_CUDA;
}

}  // namespace

int attn_varlen(const __nv_bfloat16* q, int ldq, const __nv_bfloat16* k, int ldk, const __nv_bfloat16* v, int ldv,
                __nv_bfloat16* o, int ldo, const int* cu_seqlens, int n_seqs, int n_tok, int max_len, int n_q_heads,
                int n_kv_heads, int head_dim, int causal, float scale, cudaStream_t st) {
  if (n_seqs == 0 || max_len == 0) return OPUS_OK;
  if (n_kv_heads <= 0 || n_q_heads % n_kv_heads) return OPUS_ERR_ARG;
  {
    // implementation choice: the tcgen05 kernel (attention_tc.cu; two 128-row query tiles per work item) unless every
    // sequence is shorter than 96 tokens, where the 64-row mma.sync tiles below waste fewer padded rows. Measured
    // (tools/sweep_attn.py): tcgen05 wins up to T = 2048 (363 vs 1143 us, causal hd 128).
    // OPUS_ATTN=mma|tc (read at context creation) forces one of them (A/B measurements, tests).
    const int mode = ctx().tun.attn_mode;
    const bool aligned = ((ldq | ldk | ldv | ldo) % 8) == 0 &&
                         ((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) |
                           reinterpret_cast<uintptr_t>(v) | reinterpret_cast<uintptr_t>(o)) & 15) == 0;
    // crossover measured with tools/sweep_attn_short.py (256 x T causal GQA hd 128 / 64 x T encoder hd 64): T = 64: 142 vs
    // 165 us, T = 96: 349 vs 268, T = 128: 365 vs 310, T = 160: 527 vs 356; encoder T = 66: 31 vs 44, T = 130: 65 vs 57
    const bool want_tc = mode == 2 || (mode == 0 && max_len >= 96);
    if (want_tc && aligned && (head_dim == 64 || head_dim == 128)) {
      // Short tails: a sequence whose last 256-row work item holds <= 64 query rows (T = 258: two rows) would push a
      // nearly empty 128-row tile through the whole tcgen05 pipeline (~6 us per (sequence, head), 40 % of the encoder's
      // attention time at C1). Those rows run on one 64-row mma.sync tile instead, launched first (it is tiny); the two
      // kernels write disjoint rows of `o`. Enabled when the longest sequence has such a tail (uniform batches) or for
      // ragged encoder batches; OPUS_ATTN_TAIL=0 disables it.
      const int tail_mode = ctx().tun.attn_tail;
      const int max_tail = max_len - ((max_len - 1) / 256) * 256;
      // tails of at most 16 rows (T = 258: the last residue and <eos>) take a one-warp, 16-row tile per (sequence, head)
      // instead of a 64-row tile whose other three warps only help to load K / V
      const int kTail = max_tail <= 16 ? 16 : 64;
      const bool split_tail = tail_mode && (max_tail <= kTail || (!causal && n_seqs >= 8));
      if (split_tail) {
        AttnParams tp;
        tp.q = q; tp.k = k; tp.v = v; tp.o = o;
        tp.cu_seqlens = cu_seqlens;
        tp.ldq = ldq; tp.ldk = ldk; tp.ldv = ldv; tp.ldo = ldo;
        tp.n_q_heads = n_q_heads;
        tp.group = n_q_heads / n_kv_heads;
        tp.scale_log2 = scale * 1.4426950408889634f;
        tp.tail_only = kTail;
        int rc;
        if (kTail == 16) {
          if (head_dim == 64) rc = causal ? launch_varlen<64, true, 16>(tp, n_seqs, max_len, st, true) : launch_varlen<64, false, 16>(tp, n_seqs, max_len, st, true);
          else rc = causal ? launch_varlen<128, true, 16>(tp, n_seqs, max_len, st, true) : launch_varlen<128, false, 16>(tp, n_seqs, max_len, st, true);
        } else if (head_dim == 64) rc = causal ? launch_varlen<64, true, 64>(tp, n_seqs, max_len, st, true) : launch_varlen<64, false, 64>(tp, n_seqs, max_len, st, true);
        else rc = causal ? launch_varlen<128, true, 64>(tp, n_seqs, max_len, st, true) : launch_varlen<128, false, 64>(tp, n_seqs, max_len, st, true);
        if (rc != OPUS_OK) return rc;
      }
      return attn_varlen_tc(q, ldq, k, ldk, v, ldv, o, ldo, cu_seqlens, n_seqs, n_tok, max_len, n_q_heads, n_kv_heads,
                            head_dim, causal, scale, st, split_tail ? kTail : 0);
    }
  }
  if ((ldq | ldk | ldv) % 8 || ldo % 2) return OPUS_ERR_ARG;
  AttnParams p;
  p.q = q; p.k = k; p.v = v; p.o = o;
  p.cu_seqlens = cu_seqlens;
  p.ldq = ldq; p.ldk = ldk; p.ldv = ldv; p.ldo = ldo;
  p.n_q_heads = n_q_heads;
  p.group = n_q_heads / n_kv_heads;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.tail_only = 0;
  // 64-row query tiles when they waste fewer padded rows than 128-row tiles (or the sequences are short)
  const bool bm64 = ((max_len + 63) / 64) * 64 < ((max_len + 127) / 128) * 128;
  if (head_dim == 64 && !causal) return bm64 ? launch_varlen<64, false, 64>(p, n_seqs, max_len, st) : launch_varlen<64, false, 128>(p, n_seqs, max_len, st);
  if (head_dim == 64 && causal) return bm64 ? launch_varlen<64, true, 64>(p, n_seqs, max_len, st) : launch_varlen<64, true, 128>(p, n_seqs, max_len, st);
  if (head_dim == 128 && !causal) return bm64 ? launch_varlen<128, false, 64>(p, n_seqs, max_len, st) : launch_varlen<128, false, 128>(p, n_seqs, max_len, st);
  if (head_dim == 128 && causal) return bm64 ? launch_varlen<128, true, 64>(p, n_seqs, max_len, st) : launch_varlen<128, true, 128>(p, n_seqs, max_len, st);
  return OPUS_ERR_ARG;
}

namespace {
int decode_num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}
// Context-owned scratch of the split-KV merge: partial results + one arrival counter per (sequence, kv head) unit.
// Allocated on first use outside a stream capture (the decode loop touches it before it captures), grown never.
bool decode_split_scratch(size_t need_bytes, int units, float** ws, int** cnt) {
  Context& c = ctx();
  std::lock_guard<std::mutex> lk(c.sk_mu);
  constexpr size_t kBytes = 16u << 20;
  constexpr int kUnits = 16384;
  if (c.attn_ws == nullptr) {
    if (cudaMalloc(&c.attn_ws, kBytes) != cudaSuccess || cudaMalloc(&c.attn_cnt, kUnits * sizeof(int)) != cudaSuccess ||
        cudaMemset(c.attn_cnt, 0, kUnits * sizeof(int)) != cudaSuccess) {
      cudaGetLastError();      // e.g. first call inside a capture: this launch runs unsplit
      if (c.attn_ws) { cudaFree(c.attn_ws); c.attn_ws = nullptr; }
      if (c.attn_cnt) { cudaFree(c.attn_cnt); c.attn_cnt = nullptr; }
      return false;
    }
  }
  if (need_bytes > kBytes || units > kUnits) return false;
  *ws = c.attn_ws; *cnt = c.attn_cnt;
  return true;
}
}  // namespace

int attn_decode_warmup() {
  float* ws; int* cnt;
  return decode_split_scratch(0, 0, &ws, &cnt) ? OPUS_OK : OPUS_ERR_CUDA;
}

int attn_decode_paged(const __nv_bfloat16* q, int ldq, const __nv_bfloat16* kcache, const __nv_bfloat16* vcache,
                      const int* block_table, int max_blocks, const int* ctx_len, __nv_bfloat16* o, int ldo,
                      int n_seqs, int n_q_heads, int n_kv_heads, int head_dim, int block_size, float scale,
                      cudaStream_t st) {
  return attn_decode_paged_fused(const_cast<__nv_bfloat16*>(q), ldq, nullptr, 0, nullptr, nullptr, nullptr, nullptr,
                                 const_cast<__nv_bfloat16*>(kcache), const_cast<__nv_bfloat16*>(vcache), block_table,
                                 max_blocks, ctx_len, o, ldo, n_seqs, n_q_heads, n_kv_heads, head_dim, block_size,
                                 scale, st);
}

int attn_decode_paged_fused(__nv_bfloat16* qkv, int ldq, const float* partial, int n_partial, const int* pos,
                            const int* slot, const __nv_bfloat16* cos_t, const __nv_bfloat16* sin_t,
                            __nv_bfloat16* kcache, __nv_bfloat16* vcache, const int* block_table, int max_blocks,
                            const int* ctx_len, __nv_bfloat16* o, int ldo, int n_seqs, int n_q_heads, int n_kv_heads,
                            int head_dim, int block_size, float scale, cudaStream_t st, const float* bias) {
  const __nv_bfloat16* q = qkv;
  if (n_seqs == 0) return OPUS_OK;
  if (partial != nullptr && (pos == nullptr || slot == nullptr || cos_t == nullptr || sin_t == nullptr || (ldq % 8)))
    return OPUS_ERR_ARG;
  if (head_dim != DEC_D || block_size != DEC_BS || n_kv_heads <= 0 || n_q_heads % n_kv_heads) return OPUS_ERR_ARG;
  const int group = n_q_heads / n_kv_heads;
  DecodeParams p;
  p.q = q; p.ldq = ldq;
  p.kcache = kcache; p.vcache = vcache;
  p.block_table = block_table; p.max_blocks = max_blocks;
  p.ctx_len = ctx_len;
  p.o = o; p.ldo = ldo;
  p.n_kv_heads = n_kv_heads; p.group = group;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.partial = partial; p.n_partial = n_partial;
  p.qkv = qkv; p.pos = pos; p.slot = slot; p.cos_t = cos_t; p.sin_t = sin_t;
  p.kcache_w = kcache; p.vcache_w = vcache;
  p.bias = partial != nullptr ? bias : nullptr;
  // split-KV when the (sequence, kv head) units are too coarse to balance over the SMs (see DecodeParams)
  p.n_split = 1; p.split_ws = nullptr; p.split_cnt = nullptr;
  {
    const int units = n_kv_heads * n_seqs, sms = decode_num_sms();
    const int slots = sms * 6;                           // resident CTAs (36 KB each)
    int want = 1;
    if (ctx().tun.attn_split != 0 && units < 2 * slots) {
      want = ctx().tun.attn_split > 0 ? ctx().tun.attn_split : (units * 2 <= slots ? 4 : 2);
      while (want > 1 && max_blocks < 2 * want) want >>= 1;       // at least two pages per part
    }
    if (want > 1) {
      float* ws = nullptr; int* cnt = nullptr;
      const size_t need = (size_t)units * want * group * (DEC_D + 2) * sizeof(float);
      if (decode_split_scratch(need, units, &ws, &cnt)) { p.n_split = want; p.split_ws = ws; p.split_cnt = cnt; }
    }
  }
  dim3 grid(n_kv_heads, n_seqs, p.n_split);
  const int smem = DEC_WARPS * 4 * DEC_PANEL + DEC_WARPS * group * (DEC_D + 2) * 4;
  static bool configured = false;
  if (!configured) {
    auto bytes_for = [](int g) { return DEC_WARPS * 4 * DEC_PANEL + DEC_WARPS * g * (DEC_D + 2) * 4; };
#define OPUS_DEC_CFG(G) cudaFuncSetAttribute(attn_decode_paged_kernel<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes_for(G));
    OPUS_DEC_CFG(1) OPUS_DEC_CFG(2) OPUS_DEC_CFG(3) OPUS_DEC_CFG(4) OPUS_DEC_CFG(5) OPUS_DEC_CFG(6) OPUS_DEC_CFG(7) OPUS_DEC_CFG(8)
#undef OPUS_DEC_CFG
    configured = true;
  }
  switch (group) {   // the GROUP query heads of a kv head share the rows of one m16 MMA tile: any group size up to 8
#define OPUS_DEC_CASE(G) case G: launch_pdl(true, attn_decode_paged_kernel<G>, dim3(grid), dim3(DEC_WARPS * 32), smem, st, p); break;
    OPUS_DEC_CASE(1) OPUS_DEC_CASE(2) OPUS_DEC_CASE(3) OPUS_DEC_CASE(4) OPUS_DEC_CASE(5) OPUS_DEC_CASE(6) OPUS_DEC_CASE(7) OPUS_DEC_CASE(8)
#undef OPUS_DEC_CASE
    default: return OPUS_ERR_ARG;
  }
  note_launch();
  return cudaGetLastError() == cudaSuccess ? OPUS_OK : OPUS_ERR_CUDA;
}

}  // namespace opus
