// bf16 GEMM on the 5th-gen tensor cores:  D[M,N] = epilogue( A[M,K] * B[N,K]^T ),  fp32 accumulate in TMEM.
//
//   * operands are K-major (row-major [rows, K]) bf16 in global memory, moved by TMA (SWIZZLE_128B, 64-wide K boxes)
//     into a multi-stage shared-memory ring;
//   * one elected thread issues tcgen05.mma (UMMA 128 x BN x 16) into one of two TMEM accumulator buffers;
//   * four epilogue warps drain the other accumulator with tcgen05.ld and apply the fused epilogue
//     (bias / erf-GELU / residual / SwiGLU / split-K partial), so the epilogue of tile i overlaps the MMAs of tile i+1;
//   * persistent CTAs (one per SM) walk a grouped-M raster of (m, n, k-split) work items.
//
// "Transposed" mode is the swap-AB form used by decode and the projectors: the WEIGHT matrix is the A operand
// (128 output features per tile) and the small batch is the B operand (BN = 32..256 rows), and the epilogue stores
// D^T so the result is again [batch, features] row-major. That keeps M = 128 UMMA tiles full at batch <= 256 while
// the kernel is purely weight-streaming (HBM-bound).
//
// Replaces, on the reference's path, every nn.Linear executed by cuBLAS: fair-esm q/k/v/out/fc1/fc2
// (cstp_v3/modelling.py:48), CSTP projection (modelling.py:396-400), switch projector (protein_mlp/builder.py:21-24),
// and HF Llama q/k/v/o/gate/up/down/lm_head (reached through language_model/opus_llama.py:82-93).
#include "common.h"
#include "context.h"
#include "gemm.h"
#include "launch.cuh"
#include "ptx.cuh"

#include <cuda.h>
#include <cstdlib>
#include <mutex>

namespace opus {

namespace {

constexpr int BM = 128;       // UMMA M (rows of A per tile) == TMEM lanes
constexpr int BK = 64;        // K elements per stage = one 128-byte swizzle atom of bf16
constexpr int UMMA_K = 16;    // K per tcgen05.mma for 16-bit inputs
constexpr int NUM_THREADS = 320;  // warp 0: TMA producer, warp 1: MMA issuer + TMEM owner, warps 2-9: epilogue
constexpr int NUM_EPI_THREADS = 256;  // two warps per TMEM lane quadrant, each draining every other 32-column chunk

// TR (swap-AB) launches have no TMA-store staging, so the ring takes that space too: at decode sizes the bytes in flight
// per SM set the streaming rate (measured on lm_head: 64 KB of weight tiles in flight 4.99 TB/s, 96 KB 5.99, 128 KB 6.45).
template <int BN, bool TR = false, bool NORM = false>
struct Cfg {
  static constexpr int STAGE_A = BM * BK * 2;
  static constexpr int STAGE_B = BN * BK * 2;
  static constexpr int STAGE = STAGE_A + STAGE_B;
  static constexpr int STORE_BYTES = TR ? 0 : 8 * 4096;  // epilogue staging: one [32 rows x 64 bf16] swizzled box per warp
  static constexpr int TAIL = 1024 /*align slack*/ + 256 /*barriers*/ + 1024 /*epilogue bias slice / RMSNorm row scales*/;
  static constexpr int STAGES_PLAIN = (BN >= 256) ? 4 : (BN >= 128 ? 6 : 8);
  static constexpr int STAGES_FIT = (232448 - TAIL - STORE_BYTES) / STAGE;   // 227 KB of dynamic shared memory per CTA
  static constexpr int STAGES_MAX = NORM ? 8 : 11;   // barrier block: (2 or 3) * STAGES + 4 mbarriers in 256 bytes
  static constexpr int STAGES = TR ? (STAGES_FIT > STAGES_MAX ? STAGES_MAX : STAGES_FIT) : STAGES_PLAIN;
  static constexpr int N_BARS = (NORM ? 3 : 2) * STAGES + 4;
  static constexpr int TMEM_COLS = (2 * BN < 32) ? 32 : 2 * BN;  // two accumulator buffers
  static constexpr int BARS_OFF = STAGES * STAGE + STORE_BYTES;
  static constexpr int SMEM = BARS_OFF + TAIL;
  // barriers + TMEM slot (+ the chain kernel's 32-byte reduction scratch, never used together with NORM) in 256 bytes
  static_assert(N_BARS * 8 + 8 + (NORM ? 0 : 32) <= 256, "barrier block overflows its 256 bytes");
};
constexpr int NUM_THREADS_NORM = NUM_THREADS + 128;   // + warps 10-13: RMSNorm of the activation k-slices in shared memory

enum WorkKind : int { WORK_TILE = 0, WORK_SK_PARTIAL = 1, WORK_SK_OWNER = 2 };

struct TileCoord {
  int m, n, kb_begin, kb_end, split;
  int kind;       // WorkKind
  int sk_tile;    // index of the stream-K tile
  int first_cta;  // owner only: partials of CTAs first_cta, first_cta + cta_stride, ... < blockIdx.x belong to this tile
  int cta_stride; // 1; 2 in the CTA-pair kernel (the contributors are the same-rank CTAs of the preceding pairs)
  int cnt_slot;   // arrival counter of the tile (pair kernel: one per rank)
};

// (m, n) of output tile `mn` in the grouped-M raster
__device__ __forceinline__ void raster_mn(const GemmParams& p, int mn, int& m, int& n) {
  const int per_group = p.group_m * p.num_n_tiles;
  const int g = mn / per_group;
  const int r = mn - g * per_group;
  const int first_m = g * p.group_m;
  const int gsz = min(p.num_m_tiles - first_m, p.group_m);
  m = first_m + r % gsz;
  n = r / gsz;
}

// Work item number `it` (0, 1, 2, ...) of this CTA; false when the CTA is done. The producer, the MMA issuer and the
// epilogue warps all walk the same sequence:
//   * the CTA's share of the stream-K tail: unit range [b*U/ge, (b+1)*U/ge) of U = sk_tiles * k_blocks k-block units
//     (ge = min(g, U) CTAs take part). The range is shorter than one tile's K extent (sk_tiles < g), so it touches at
//     most two tiles: the head of a tile it does not finish (partial dump) and the tail of a tile it finishes (owner:
//     fix-up + epilogue);
//   * round-robin items t = blockIdx.x + i * gridDim.x < dp_items: whole (m, n, k-split) tiles.
// Order: round-robin tiles first, then the partial piece, then the piece(s) this CTA finishes (no CTA ever waits on a CTA
// that is itself waiting). Walking the partial piece before the tiles was measured slower (gate/up 42.4 -> 43.7 us).
// b = index of the unit walking the sequence (CTA, or CTA pair), g = number of such units
__device__ __forceinline__ bool get_work_bg(const GemmParams& p, int it, TileCoord& c, const int b, const int g) {
  c.kind = WORK_TILE; c.sk_tile = 0; c.first_cta = 0; c.cta_stride = 1; c.cnt_slot = 0;
  const int n_dp = b < p.dp_items ? (p.dp_items - b + g - 1) / g : 0;
  // ---- tail pieces of this CTA
  int n_pre = 0, n_post = 0;
  int tile[1], k0[1], k1[1];        // the piece this CTA does not finish (partial dump), if any
  int ptile[2], pk0[2], pk1[2];     // the piece(s) it finishes, in walking order
  const int kb = p.k_blocks;
  if (p.sk_tiles > 0) {
    const long long U = (long long)p.sk_tiles * kb;
    int ge = U < g ? (int)U : g;  // CTAs sharing the tail: every one of them gets at least one unit
    if (p.sk_share > 0 && p.sk_share < ge) ge = p.sk_share;   // a small tail: few pieces per tile, short fix-ups
    if (b < ge) {
      const int u0 = (int)((long long)b * U / ge), u1 = (int)((long long)(b + 1) * U / ge);
      const int tA = u0 / kb, tB = (u1 - 1) / kb;
      auto add = [&](int t, int a0, int a1) {
        if (a1 < kb) { tile[0] = t; k0[0] = a0; k1[0] = a1; n_pre = 1; }           // not finished here: partial
        else { ptile[n_post] = t; pk0[n_post] = a0; pk1[n_post] = a1; ++n_post; }  // finished here
      };
      if (tA == tB) {
        add(tA, u0 - tA * kb, u1 - tA * kb);
      } else {
        add(tB, 0, u1 - tB * kb);        // head of the next tile
        add(tA, u0 - tA * kb, kb);       // tail of the previous tile
      }
      if (it >= n_dp + n_pre) {
        const int j = it - n_dp - n_pre;
        if (j >= n_post) return false;
        c.split = 0;
        raster_mn(p, p.dp_items + ptile[j], c.m, c.n);       // stream-K is only used with split_k == 1
        c.kb_begin = pk0[j]; c.kb_end = pk1[j];
        c.sk_tile = c.cnt_slot = ptile[j];
        // CTA holding unit x is ((x + 1) * ge - 1) / U
        c.first_cta = (int)((((long long)ptile[j] * kb + 1) * ge - 1) / U);
        c.kind = (c.first_cta == b) ? WORK_TILE : WORK_SK_OWNER;  // whole tile in this CTA's range: nothing to fix up
        return true;
      }
      if (it >= n_dp) {
        c.split = 0;
        raster_mn(p, p.dp_items + tile[0], c.m, c.n);
        c.kb_begin = k0[0]; c.kb_end = k1[0];
        c.sk_tile = c.cnt_slot = tile[0];
        c.kind = WORK_SK_PARTIAL;
        return true;
      }
    }
  }
  if (it >= n_dp) return false;
  const int t = b + it * g;
  const int mn = t / p.split_k;
  c.split = t - mn * p.split_k;
  raster_mn(p, mn, c.m, c.n);
  c.kb_begin = (int)(((long long)c.split * p.k_blocks) / p.split_k);
  c.kb_end = (int)(((long long)(c.split + 1) * p.k_blocks) / p.split_k);
  if (p.sk_fix && p.split_k > 1) {
    // one wave (t == blockIdx.x): the CTAs of a tile are neighbours, the one with the last split finishes the tile
    c.sk_tile = c.cnt_slot = mn;
    c.first_cta = t - c.split;
    c.kind = (c.split == p.split_k - 1) ? WORK_SK_OWNER : WORK_SK_PARTIAL;
  }
  return true;
}
__device__ __forceinline__ bool get_work(const GemmParams& p, int it, TileCoord& c) {
  return get_work_bg(p, it, c, (int)blockIdx.x, (int)gridDim.x);
}

__device__ __forceinline__ void bf16x8_unpack(const uint4 q, float* f) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 t = __bfloat1622float2(h[j]);
    f[2 * j] = t.x;
    f[2 * j + 1] = t.y;
  }
}

// ---- epilogue for one 32-column chunk held by one thread (= one accumulator row) ----
__device__ __forceinline__ void store_bf16x8(__nv_bfloat16* dst, const float* v) {
  uint4 q;
  q.x = pack_bf16x2(v[0], v[1]);
  q.y = pack_bf16x2(v[2], v[3]);
  q.z = pack_bf16x2(v[4], v[5]);
  q.w = pack_bf16x2(v[6], v[7]);
  *reinterpret_cast<uint4*>(dst) = q;
}

// Transposed-form bf16 stores. Thread = accumulator row (consecutive lanes = consecutive output addresses), so a plain
// store is 2 bytes per lane — measured ~10 us slower per tile than 4-byte stores. Lane pairs therefore exchange one value
// per column pair: even lanes write column i for rows (l, l+1), odd lanes write column i+1 for rows (l-1, l), 4 bytes each.
// x[i] = value of (this thread's row, column i); `base` points at (column 0, this thread's row); requires even `row`
// alignment of the pair (row of an even lane is even) and an even number of valid rows.
template <int NC>
__device__ __forceinline__ void store_bf16_transposed_paired(__nv_bfloat16* base, size_t ld, const float (&x)[NC],
                                                            int n_valid, int lane) {
  const bool odd = lane & 1;
#pragma unroll
  for (int i = 0; i < NC; i += 2) {
    const float send = odd ? x[i] : x[i + 1];
    const float recv = __shfl_xor_sync(0xffffffffu, send, 1);
    const int col = odd ? i + 1 : i;
    const uint32_t packed = odd ? pack_bf16x2(recv, x[i + 1]) : pack_bf16x2(x[i], recv);
    if (col < n_valid) *reinterpret_cast<uint32_t*>(base + (size_t)col * ld - (odd ? 1 : 0)) = packed;
  }
}

__device__ __forceinline__ void epilogue_chunk(const GemmParams& p, const uint32_t (&r)[32], int row, int col0,
                                               int split) {
  float v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);

  if (!p.transposed) {
    // element (row, col0+i) -> out[row*ldo + col0+i]; bias indexed by column
    const bool row_ok = row < p.M;
    if (p.epi == EPI_PARTIAL_F32) {
      if (row_ok) {
        float* dst = reinterpret_cast<float*>(p.out) + ((size_t)split * p.M + row) * p.ldo + col0;
#pragma unroll
        for (int g = 0; g < 8; ++g)
          if (col0 + g * 4 + 4 <= p.N) *reinterpret_cast<float4*>(dst + g * 4) = make_float4(v[g * 4], v[g * 4 + 1], v[g * 4 + 2], v[g * 4 + 3]);
      }
      return;
    }
    if (p.bias != nullptr) {
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        if (col0 + g * 4 + 4 <= p.N) {
          const float4 b = *reinterpret_cast<const float4*>(p.bias + col0 + g * 4);
          v[g * 4] += b.x; v[g * 4 + 1] += b.y; v[g * 4 + 2] += b.z; v[g * 4 + 3] += b.w;
        }
      }
    }
    if (p.epi == EPI_BF16 || p.epi == EPI_BF16_GELU || p.epi == EPI_BF16_RELU) {
      if (p.epi == EPI_BF16_GELU) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = gelu_erf(v[i]);
      } else if (p.epi == EPI_BF16_RELU) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.0f);
      }
      if (row_ok) {
        __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.out) + (size_t)row * p.ldo + col0;
#pragma unroll
        for (int g = 0; g < 4; ++g)
          if (col0 + g * 8 + 8 <= p.N) store_bf16x8(dst + g * 8, v + g * 8);
      }
    } else if (p.epi == EPI_RES_F32) {
      // out_f32 = residual_f32 + (acc + bias)   (fp32 residual stream of the encoder; in-place allowed)
      if (row_ok) {
        const float* res = reinterpret_cast<const float*>(p.residual) + (size_t)row * p.ldr + col0;
        float* dst = reinterpret_cast<float*>(p.out) + (size_t)row * p.ldo + col0;
        // all residual loads first: `out` may alias `residual` (in-place), so interleaving load/store pairs would
        // serialise eight L2 round trips per chunk
        float4 q[8];
#pragma unroll
        for (int g = 0; g < 8; ++g)
          q[g] = (col0 + g * 4 + 4 <= p.N) ? *reinterpret_cast<const float4*>(res + g * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          if (col0 + g * 4 + 4 <= p.N) {
            *reinterpret_cast<float4*>(dst + g * 4) =
                make_float4(q[g].x + v[g * 4], q[g].y + v[g * 4 + 1], q[g].z + v[g * 4 + 2], q[g].w + v[g * 4 + 3]);
          }
        }
      }
    } else if (p.epi == EPI_RES_BF16) {
      // out = bf16( residual + bf16(acc) )  -- mirrors `residual + self.o_proj(x)` in bf16 (HF modeling_llama.py)
      if (row_ok) {
        const __nv_bfloat16* res = reinterpret_cast<const __nv_bfloat16*>(p.residual) + (size_t)row * p.ldr + col0;
        __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.out) + (size_t)row * p.ldo + col0;
        uint4 qq[4];
#pragma unroll
        for (int g = 0; g < 4; ++g)
          qq[g] = (col0 + g * 8 + 8 <= p.N) ? *reinterpret_cast<const uint4*>(res + g * 8) : make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          if (col0 + g * 8 + 8 <= p.N) {
            const uint4 q = qq[g];
            const __nv_bfloat162* rb = reinterpret_cast<const __nv_bfloat162*>(&q);
            float o[8];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float2 rf = __bfloat1622float2(rb[j]);
              o[2 * j] = rf.x + bf16_round(v[g * 8 + 2 * j]);
              o[2 * j + 1] = rf.y + bf16_round(v[g * 8 + 2 * j + 1]);
            }
            store_bf16x8(dst + g * 8, o);
          }
        }
      }
    } else if (p.epi == EPI_SWIGLU) {
      // columns (2j, 2j+1) = (gate_j, up_j) -> out[row, j] = bf16( bf16(silu(bf16 gate)) * bf16 up )
      if (row_ok) {
        float o[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float g = bf16_round(v[2 * j]);
          const float u = bf16_round(v[2 * j + 1]);
          o[j] = bf16_round(silu_f(g)) * u;
        }
        __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.out) + (size_t)row * p.ldo + (col0 >> 1);
        if (col0 + 16 <= p.N) store_bf16x8(dst, o);
        if (col0 + 32 <= p.N) store_bf16x8(dst + 8, o + 8);
      }
    }
    return;
  }

}


// ---- transposed (swap-AB) epilogue: accumulator row = output feature `row`, column = batch row n -> out[n*ldo + row].
// NC columns per call. It is instantiated with NC = 8 and driven by a ROLLED loop: a decode-sized launch executes this
// code exactly once per CTA with a cold instruction cache, and a 32-way unrolled body cost ~6-18 us per launch.
// dry = instruction-cache warm-up pass (see the epilogue warps of gemm_bf16_tcgen05_kernel): same instructions, every
// memory access masked off.
template <int NC>
__device__ __forceinline__ void epilogue_transposed(const GemmParams& p, const uint32_t (&r)[NC], int row, int col0,
                                                    int split, bool dry = false) {
  float v[NC];
#pragma unroll
  for (int i = 0; i < NC; ++i) v[i] = __uint_as_float(r[i]);
  if (dry) row = p.M + 2;   // beyond the matrix: every load / store predicate below turns false
  const bool row_ok = row < p.M;
  if (p.epi == EPI_PARTIAL_F32) {
    float* base = reinterpret_cast<float*>(p.out) + ((size_t)split * p.N + col0) * p.ldo + row;
    const int nv = row_ok ? min(NC, p.N - col0) : 0;
#pragma unroll
    for (int i = 0; i < NC; ++i)
      if (i < nv) base[(size_t)i * p.ldo] = v[i];
    return;
  }
  if (p.epi == EPI_SWIGLU) {
    // rows (2j, 2j+1) = (gate_j, up_j) live in adjacent lanes
    __nv_bfloat16* base = reinterpret_cast<__nv_bfloat16*>(p.out) + (size_t)col0 * p.ldo + (row >> 1);
    float o[NC];
#pragma unroll
    for (int i = 0; i < NC; ++i) {
      const float other = __shfl_down_sync(0xffffffffu, v[i], 1);
      o[i] = bf16_round(silu_f(bf16_round(v[i]))) * bf16_round(other);  // meaningful on even lanes only
    }
    // even lanes hold out[row/2]; lanes l and l+2 pair up so that every store is 4 bytes (see the note above)
    const int lane = row & 31;
    const bool hi = lane & 2;
    const int nv = (((lane & 1) == 0) && (row | 3) < p.M) ? min(NC, p.N - col0) : 0;   // dry: row >= M
#pragma unroll
    for (int i = 0; i < NC; i += 2) {
      const float send = hi ? o[i] : o[i + 1];
      const float recv = __shfl_xor_sync(0xffffffffu, send, 2);
      const int col = hi ? i + 1 : i;
      const uint32_t packed = hi ? pack_bf16x2(recv, o[i + 1]) : pack_bf16x2(o[i], recv);
      if (col < nv) *reinterpret_cast<uint32_t*>(base + (size_t)col * p.ldo - (hi ? 1 : 0)) = packed;
    }
    return;
  }
  // Remaining modes: one specialised, branch-free store loop per epilogue (mode checks hoisted out of the element loop;
  // residual values are all loaded before the first store because `out` may alias `residual`).
  const float b = (p.bias != nullptr && row_ok) ? p.bias[row] : 0.0f;
  const int n_valid = row_ok ? min(NC, p.N - col0) : 0;  // columns (batch rows) of this chunk that exist
  if (p.epi == EPI_BF16 || p.epi == EPI_BF16_GELU || p.epi == EPI_BF16_RELU) {
    __nv_bfloat16* base = reinterpret_cast<__nv_bfloat16*>(p.out) + (size_t)col0 * p.ldo + row;
    const bool gelu = p.epi == EPI_BF16_GELU;
    const float floor_v = p.epi == EPI_BF16_RELU ? 0.0f : -INFINITY;   // ReLU (OPT fc1) as a clamp: no extra branch
    float x[NC];
#pragma unroll
    for (int i = 0; i < NC; ++i) {
      x[i] = v[i] + b;
      if (gelu) x[i] = gelu_erf(x[i]);
      x[i] = fmaxf(x[i], floor_v);
    }
    // M (features) is even and tiles start at multiples of 128, so a lane pair is either fully valid or fully invalid
    const int nv = (row | 1) < p.M ? min(NC, p.N - col0) : 0;
    store_bf16_transposed_paired(base, (size_t)p.ldo, x, nv, row & 31);
  } else if (p.epi == EPI_F32) {
    float* base = reinterpret_cast<float*>(p.out) + (size_t)col0 * p.ldo + row;
#pragma unroll
    for (int i = 0; i < NC; ++i)
      if (i < n_valid) base[(size_t)i * p.ldo] = v[i] + b;
  } else if (p.epi == EPI_RES_BF16) {
    const __nv_bfloat16* rbase = reinterpret_cast<const __nv_bfloat16*>(p.residual) + (size_t)col0 * p.ldr + row;
    __nv_bfloat16* base = reinterpret_cast<__nv_bfloat16*>(p.out) + (size_t)col0 * p.ldo + row;
    float rr[NC];
#pragma unroll
    for (int i = 0; i < NC; ++i) rr[i] = (i < n_valid) ? __bfloat162float(rbase[(size_t)i * p.ldr]) : 0.f;
    float x[NC];
#pragma unroll
    for (int i = 0; i < NC; ++i) x[i] = rr[i] + bf16_round(v[i] + b);
    const int nv = (row | 1) < p.M ? min(NC, p.N - col0) : 0;
    store_bf16_transposed_paired(base, (size_t)p.ldo, x, nv, row & 31);
    if constexpr (NC == 8) {
      if (p.sumsq_out != nullptr) {
        // sum of squares of the stored bf16 values over this warp's 32 features, per batch column: transpose-reduce (at
        // every step a lane hands half of its columns to its partner), then two plain butterfly steps. Fixed order.
        const int lane = row & 31;
        float a[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float hv = (i < n_valid) ? bf16_round(x[i]) : 0.f;
          a[i] = hv * hv;
        }
        const bool u16 = lane & 16, u8 = lane & 8, u4 = lane & 4;
        float bq[4], cq[2];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float recv = __shfl_xor_sync(0xffffffffu, u16 ? a[j] : a[j + 4], 16);
          bq[j] = (u16 ? a[j + 4] : a[j]) + recv;
        }
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const float recv = __shfl_xor_sync(0xffffffffu, u8 ? bq[j] : bq[j + 2], 8);
          cq[j] = (u8 ? bq[j + 2] : bq[j]) + recv;
        }
        float d = (u4 ? cq[1] : cq[0]) + __shfl_xor_sync(0xffffffffu, u4 ? cq[0] : cq[1], 4);
        d += __shfl_xor_sync(0xffffffffu, d, 2);
        d += __shfl_xor_sync(0xffffffffu, d, 1);
        const int c = (u16 ? 4 : 0) + (u8 ? 2 : 0) + (u4 ? 1 : 0);
        if (!dry && (lane & 3) == 0 && col0 + c < p.N) p.sumsq_out[(size_t)(row >> 5) * p.sumsq_ld + col0 + c] = d;
      }
    }
  } else if (p.epi == EPI_RES_F32) {
    const float* rbase = reinterpret_cast<const float*>(p.residual) + (size_t)col0 * p.ldr + row;
    float* base = reinterpret_cast<float*>(p.out) + (size_t)col0 * p.ldo + row;
    float rr[NC];
#pragma unroll
    for (int i = 0; i < NC; ++i) rr[i] = (i < n_valid) ? rbase[(size_t)i * p.ldr] : 0.f;
#pragma unroll
    for (int i = 0; i < NC; ++i)
      if (i < n_valid) base[(size_t)i * p.ldo] = rr[i] + (v[i] + b);
  }
}

// Epilogue of one work item by the 8 epilogue warps: drain the accumulator buffer `acc` (TMEM -> registers) and either
// run the fused epilogue, dump a stream-K partial, or (owner) add the other CTAs' partials first.
template <int BN, bool TR>
__device__ __forceinline__ void epilogue_item(const GemmParams& p, const TileCoord& tc, uint32_t tmem_base, int acc,
                                              int quad, int lane, int chunk0, int epi_tid,
                                              const CUtensorMap* tm_out = nullptr, uint8_t* stage = nullptr,
                                              bool dry = false, uint64_t* acc_bar = nullptr, uint32_t acc_parity = 0) {
  // acc_bar != nullptr: the accumulator-ready barrier has NOT been waited for yet; this function waits for it after the
  // owner's fix-up operands (the other CTAs' partial sums) have been requested, so that L2 round trip overlaps the tail
  // of the item's own main loop instead of following it.
  const int lrow = quad * 32 + lane;  // accumulator row inside the tile
  const int row = tc.m * BM + lrow;
  const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * BN;
  // stream-K partials of this tile: sk_ws[cta][column][row] fp32 (lanes = consecutive rows -> coalesced)
  float* my_part = p.sk_ws + (size_t)blockIdx.x * (BM * BN) + lrow;
  if (tc.kind == WORK_SK_OWNER) {
    if (epi_tid == 0 && !dry) {
      const int need = ((int)blockIdx.x - tc.first_cta) / tc.cta_stride;
      const long long t0 = clock64();
      while (ld_acquire_gpu(p.sk_cnt + tc.cnt_slot) < need) {
        if (clock64() - t0 > 20000000000LL) {
          printf("opus_b200: stream-K fix-up wait timed out (block %d, tile %d)\n", (int)blockIdx.x, tc.sk_tile);
          __trap();
        }
      }
      p.sk_cnt[tc.cnt_slot] = 0;  // re-armed for the next launch (stream order separates launches)
    }
    named_bar_sync(1, NUM_EPI_THREADS);
  }
  if constexpr (TR) {
    // Owner fix-up, decode-sized tiles: the partial sums of ALL column groups of this thread are requested at once per
    // contributing CTA (one L2 round trip per contributor instead of one per group), summed in CTA order.
    constexpr int NG = BN / 16;                 // 8-column groups per warp (the two warps of a quadrant interleave)
    constexpr bool kPrefetchFix = NG <= 4;      // BN <= 64: 32 registers
    float fix[kPrefetchFix ? NG : 1][8];
    if (kPrefetchFix && tc.kind == WORK_SK_OWNER) {
#pragma unroll
      for (int q = 0; q < NG; ++q)
#pragma unroll
        for (int i = 0; i < 8; ++i) fix[q][i] = 0.f;
      for (int c = dry ? (int)blockIdx.x : tc.first_cta; c < (int)blockIdx.x; c += tc.cta_stride) {
        const float* src = p.sk_ws + (size_t)c * (BM * BN) + lrow;
        float t[kPrefetchFix ? NG : 1][8];
#pragma unroll
        for (int q = 0; q < NG; ++q)
#pragma unroll
          for (int i = 0; i < 8; ++i) t[q][i] = __ldcg(src + (size_t)((chunk0 + 2 * q) * 8 + i) * BM);
#pragma unroll
        for (int q = 0; q < NG; ++q)
#pragma unroll
          for (int i = 0; i < 8; ++i) fix[q][i] += t[q][i];
      }
    }
    if (acc_bar != nullptr && !dry) {
      mbar_wait(acc_bar, acc_parity);
      tc_fence_after();
    }
    if constexpr (kPrefetchFix) {
      // Decode-sized tiles: every TMEM load of this warp is issued before the first global store. (A tcgen05.ld issued
      // after stores waits for them: with one load per 8-column group each group cost a store round trip, ~1.9 us.)
      uint32_t av[NG][8];
#pragma unroll
      for (int q = 0; q < NG; ++q) tmem_ld_32x8(taddr + (chunk0 + 2 * q) * 8, av[q]);
      tmem_ld_wait();
#pragma unroll
      for (int q = 0; q < NG; ++q) {
        const int g = chunk0 + 2 * q;
        if (tc.kind == WORK_SK_PARTIAL) {
          if (!dry) {
#pragma unroll
            for (int i = 0; i < 8; ++i) my_part[(size_t)(g * 8 + i) * BM] = __uint_as_float(av[q][i]);
          }
          continue;
        }
        if (tc.kind == WORK_SK_OWNER) {
#pragma unroll
          for (int i = 0; i < 8; ++i) av[q][i] = __float_as_uint(__uint_as_float(av[q][i]) + fix[q][i]);
        }
        const int col0 = tc.n * BN + g * 8;
        if (col0 < p.N) epilogue_transposed<8>(p, av[q], row, col0, tc.split, dry);
      }
    } else {
      // rolled loop over 8-column groups (batch > 64)
#pragma unroll 1
      for (int g = chunk0; g < BN / 8; g += 2) {
        uint32_t r8[8];
        tmem_ld_32x8(taddr + g * 8, r8);
        tmem_ld_wait();
        if (tc.kind == WORK_SK_PARTIAL) {
          if (!dry) {
#pragma unroll
            for (int i = 0; i < 8; ++i) my_part[(size_t)(g * 8 + i) * BM] = __uint_as_float(r8[i]);
          }
          continue;
        }
        if (tc.kind == WORK_SK_OWNER) {
          for (int c = dry ? (int)blockIdx.x : tc.first_cta; c < (int)blockIdx.x; c += tc.cta_stride) {
            const float* src = p.sk_ws + (size_t)c * (BM * BN) + (size_t)(g * 8) * BM + lrow;
#pragma unroll
            for (int i = 0; i < 8; ++i) r8[i] = __float_as_uint(__uint_as_float(r8[i]) + __ldcg(src + (size_t)i * BM));
          }
        }
        const int col0 = tc.n * BN + g * 8;
        if (col0 < p.N) epilogue_transposed<8>(p, r8, row, col0, tc.split, dry);
      }
    }
  } else {
  if (acc_bar != nullptr && !dry) {
    mbar_wait(acc_bar, acc_parity);
    tc_fence_after();
  }
  if (p.tma_store && tm_out != nullptr && tc.kind == WORK_TILE) {
    // bf16 (+bias, +GELU) tiles leave through shared memory: a thread owns one accumulator ROW, so direct stores put 32
    // different 128-byte lines behind every store instruction (measured: the K = 1280 encoder GEMMs ran at the pace of
    // their output bytes, 0.7 TB/s). Each warp packs a [32 rows x 64 columns] box into its own swizzled staging buffer
    // and lane 0 hands it to the TMA unit, which writes whole lines and clips at the M / N edges.
    uint8_t* my_stage = stage + (quad + 4 * chunk0) * 4096;
    const uint32_t st_row = smem_u32(my_stage) + lane * 128;
    if constexpr (BN == 256) {
      if (p.epi == EPI_SWIGLU) {
        // interleaved (gate, up) columns: this warp pair owns input columns [chunk0*128, +128) of the tile = 64 output
        // columns = one store box. HF rounding points: bf16(silu(bf16 gate)) * bf16 up, rounded once more on the pack.
        const int ocol0 = (tc.n * BN + chunk0 * 128) >> 1;          // first output column of the box
        if (tc.n * BN + chunk0 * 128 < p.N) {
          uint32_t pk[32];
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            uint32_t r[64];
            tmem_ld_32x32(taddr + chunk0 * 128 + hf * 64, *reinterpret_cast<uint32_t(*)[32]>(&r[0]));
            tmem_ld_32x32(taddr + chunk0 * 128 + hf * 64 + 32, *reinterpret_cast<uint32_t(*)[32]>(&r[32]));
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              float o[2];
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                const float g = bf16_round(__uint_as_float(r[2 * (j + e)]));
                const float u = bf16_round(__uint_as_float(r[2 * (j + e) + 1]));
                o[e] = bf16_round(silu_f(g)) * u;
              }
              pk[hf * 16 + (j >> 1)] = pack_bf16x2(o[0], o[1]);
            }
          }
          if (lane == 0) tma_store_wait_read();
          __syncwarp();
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const int ch = q ^ (lane & 7);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(st_row + ch * 16), "r"(pk[q * 4]),
                         "r"(pk[q * 4 + 1]), "r"(pk[q * 4 + 2]), "r"(pk[q * 4 + 3])
                         : "memory");
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(tm_out, my_stage, ocol0, tc.m * BM + quad * 32);
            tma_store_commit();
          }
        }
        return;
      }
      if (p.rl_cos != nullptr) {
        // ---- Llama q|k|v projection: rotary embedding + paged KV append in the epilogue. This warp pair (chunk0)
        // owns head 2n + chunk0 of the tile; a thread holds its token's 128 accumulator columns of that head.
        const int head = tc.n * 2 + chunk0;
        const int colh = head * 128;
        if (colh < p.N) {
          const bool ok_row = row < p.M;
          const bool is_v = head >= p.rl_hq + p.rl_hkv;
          const bool to_cache = head >= p.rl_hq;
          const int pos = ok_row ? p.rl_pos[row] : 0;
          const int sl = (ok_row && to_cache && p.rl_slot != nullptr) ? p.rl_slot[row] : -1;
          const __nv_bfloat16* cos_t = static_cast<const __nv_bfloat16*>(p.rl_cos) + (size_t)pos * 128;
          const __nv_bfloat16* sin_t = static_cast<const __nv_bfloat16*>(p.rl_sin) + (size_t)pos * 128;
          __nv_bfloat16* cdst = nullptr;
          if (sl >= 0) {
            const int kvh = is_v ? head - p.rl_hq - p.rl_hkv : head - p.rl_hq;
            const int blk = sl / p.rl_bs, off = sl - blk * p.rl_bs;
            cdst = static_cast<__nv_bfloat16*>(is_v ? p.rl_vcache : p.rl_kcache) +
                   (((size_t)blk * p.rl_hkv + kvh) * p.rl_bs + off) * 128;
          }
          const uint32_t th = taddr + chunk0 * 128;   // TMEM columns of this head
          const float* bq = p.bias != nullptr ? p.bias + colh : nullptr;
#pragma unroll 1
          for (int half = 0; half < 2; ++half) {
            uint32_t pk[32];
#pragma unroll
            for (int sub = 0; sub < 4; ++sub) {        // 16 rotation pairs per step keeps the live set small
              uint32_t a1[16], a2[16];
              tmem_ld_32x16(th + sub * 16, a1);        // columns j      (first half of the head)
              tmem_ld_32x16(th + 64 + sub * 16, a2);   // columns j + 64 (second half)
              tmem_ld_wait();
              float cs[16], sn[16];
              if (!is_v) {
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                  bf16x8_unpack(__ldg(reinterpret_cast<const uint4*>(cos_t + sub * 16 + q * 8)), &cs[q * 8]);
                  bf16x8_unpack(__ldg(reinterpret_cast<const uint4*>(sin_t + sub * 16 + q * 8)), &sn[q * 8]);
                }
              }
#pragma unroll
              for (int i = 0; i < 16; i += 2) {
                float o[2];
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                  // the bf16 value the plain epilogue stores (+ projection bias, Qwen2)
                  const float x1 = bf16_round(__uint_as_float(a1[i + e]) + (bq != nullptr ? __ldg(bq + sub * 16 + i + e) : 0.f));
                  const float x2 = bf16_round(__uint_as_float(a2[i + e]) + (bq != nullptr ? __ldg(bq + 64 + sub * 16 + i + e) : 0.f));
                  if (is_v) o[e] = half == 0 ? x1 : x2;
                  else if (half == 0) o[e] = bf16_round(bf16_round(x1 * cs[i + e]) + bf16_round(-x2 * sn[i + e]));
                  else o[e] = bf16_round(bf16_round(x2 * cs[i + e]) + bf16_round(x1 * sn[i + e]));
                }
                pk[sub * 8 + (i >> 1)] = pack_bf16x2(o[0], o[1]);
              }
            }
            if (cdst != nullptr) {
#pragma unroll
              for (int q = 0; q < 8; ++q)
                *reinterpret_cast<uint4*>(cdst + half * 64 + q * 8) = make_uint4(pk[q * 4], pk[q * 4 + 1], pk[q * 4 + 2], pk[q * 4 + 3]);
            }
            if (lane == 0) tma_store_wait_read();
            __syncwarp();
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const int ch = q ^ (lane & 7);
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(st_row + ch * 16), "r"(pk[q * 4]),
                           "r"(pk[q * 4 + 1]), "r"(pk[q * 4 + 2]), "r"(pk[q * 4 + 3])
                           : "memory");
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(tm_out, my_stage, colh + half * 64, tc.m * BM + quad * 32);
              tma_store_commit();
            }
          }
        }
        return;
      }
    }
    // the tile's bias slice goes through shared memory once (one float per epilogue thread, BN <= 256): per-group global
    // loads showed up as long-scoreboard stalls in front of every add (ncu source view)
    float* s_bias = reinterpret_cast<float*>(stage + Cfg<BN, false>::STORE_BYTES) + 64;   // after the barriers (256 B)
    if (p.bias != nullptr) {
      named_bar_sync(2, NUM_EPI_THREADS);                       // the previous tile's readers are done
      const int bc = tc.n * BN + epi_tid;
      if (epi_tid < BN) s_bias[epi_tid] = bc < p.N ? __ldg(p.bias + bc) : 0.f;
      named_bar_sync(2, NUM_EPI_THREADS);
    }
#pragma unroll 1
    for (int gi = chunk0; gi < BN / 64; gi += 2) {
      const int col0 = tc.n * BN + gi * 64;
      if (col0 >= p.N) break;
      uint32_t r[64];
      tmem_ld_32x32(taddr + gi * 64, *reinterpret_cast<uint32_t(*)[32]>(&r[0]));
      tmem_ld_32x32(taddr + gi * 64 + 32, *reinterpret_cast<uint32_t(*)[32]>(&r[32]));
      tmem_ld_wait();
      uint32_t pk[32];
#pragma unroll
      for (int g4 = 0; g4 < 16; ++g4) {
        float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.bias != nullptr) b = *reinterpret_cast<const float4*>(s_bias + gi * 64 + g4 * 4);
        float v0 = __uint_as_float(r[g4 * 4]) + b.x, v1 = __uint_as_float(r[g4 * 4 + 1]) + b.y;
        float v2 = __uint_as_float(r[g4 * 4 + 2]) + b.z, v3 = __uint_as_float(r[g4 * 4 + 3]) + b.w;
        if (p.epi == EPI_BF16_GELU) { v0 = gelu_erf(v0); v1 = gelu_erf(v1); v2 = gelu_erf(v2); v3 = gelu_erf(v3); }
        if (p.epi == EPI_BF16_RELU) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); v2 = fmaxf(v2, 0.f); v3 = fmaxf(v3, 0.f); }
        pk[g4 * 2] = pack_bf16x2(v0, v1);
        pk[g4 * 2 + 1] = pack_bf16x2(v2, v3);
      }
      if (p.rope_cos != nullptr && col0 < p.rope_cols) {
        // fused ESM rotary: this thread holds the whole 64-wide head of its token. Same arithmetic as rope_esm_kernel:
        // x = bf16(acc + bias) (the value the unfused path stores), scaled, rotated in fp32, rounded once.
        const int pos = row < p.M ? p.rope_pos[row] : 0;
        const float sc = col0 < p.rope_q_cols ? p.rope_q_scale : 1.0f;
        const float4* ct = reinterpret_cast<const float4*>(p.rope_cos + (size_t)pos * 32);
        const float4* st4 = reinterpret_cast<const float4*>(p.rope_sin + (size_t)pos * 32);
#pragma unroll
        for (int g4 = 0; g4 < 8; ++g4) {
          const float4 c4 = __ldg(ct + g4), s4 = __ldg(st4 + g4);
          const float cc[4] = {c4.x, c4.y, c4.z, c4.w}, ss[4] = {s4.x, s4.y, s4.z, s4.w};
          const __nv_bfloat162* lo = reinterpret_cast<const __nv_bfloat162*>(&pk[g4 * 2]);
          const __nv_bfloat162* hi = reinterpret_cast<const __nv_bfloat162*>(&pk[16 + g4 * 2]);
          const float2 l0 = __bfloat1622float2(lo[0]), l1 = __bfloat1622float2(lo[1]);
          const float2 h0 = __bfloat1622float2(hi[0]), h1 = __bfloat1622float2(hi[1]);
          const float x1[4] = {l0.x * sc, l0.y * sc, l1.x * sc, l1.y * sc};
          const float x2[4] = {h0.x * sc, h0.y * sc, h1.x * sc, h1.y * sc};
          float o1[4], o2[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            o1[i] = x1[i] * cc[i] - x2[i] * ss[i];
            o2[i] = x2[i] * cc[i] + x1[i] * ss[i];
          }
          pk[g4 * 2] = pack_bf16x2(o1[0], o1[1]);
          pk[g4 * 2 + 1] = pack_bf16x2(o1[2], o1[3]);
          pk[16 + g4 * 2] = pack_bf16x2(o2[0], o2[1]);
          pk[16 + g4 * 2 + 1] = pack_bf16x2(o2[2], o2[3]);
        }
      }
      if (lane == 0) tma_store_wait_read();   // the previous box of this warp has left the staging buffer
      __syncwarp();
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int ch = q ^ (lane & 7);          // SWIZZLE_128B: 16-byte chunk index XOR (row % 8)
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(st_row + ch * 16), "r"(pk[q * 4]), "r"(pk[q * 4 + 1]),
                     "r"(pk[q * 4 + 2]), "r"(pk[q * 4 + 3])
                     : "memory");
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        tma_store_2d(tm_out, my_stage, col0, tc.m * BM + quad * 32);
        tma_store_commit();
      }
    }
  } else {
#pragma unroll 1
    for (int c = chunk0; c < BN / 32; c += 2) {
      uint32_t r[32];
      tmem_ld_32x32(taddr + c * 32, r);
      tmem_ld_wait();
      if (tc.kind == WORK_SK_PARTIAL) {
#pragma unroll
        for (int i = 0; i < 32; ++i) my_part[(size_t)(c * 32 + i) * BM] = __uint_as_float(r[i]);
        continue;
      }
      if (tc.kind == WORK_SK_OWNER) {
        for (int cc = tc.first_cta; cc < (int)blockIdx.x; cc += tc.cta_stride) {
          const float* src = p.sk_ws + (size_t)cc * (BM * BN) + (size_t)(c * 32) * BM + lrow;
#pragma unroll
          for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) + __ldcg(src + (size_t)i * BM));
        }
      }
      const int col0 = tc.n * BN + c * 32;
      if (col0 < p.N) epilogue_chunk(p, r, row, col0, tc.split);
    }
  }
  }  // plain form
  if (tc.kind == WORK_SK_PARTIAL) {
    named_bar_sync(1, NUM_EPI_THREADS);   // every thread's partial stores happen-before thread 0's fence + release
    if (epi_tid == 0 && !dry) {
      __threadfence();
      red_release_gpu_add(p.sk_cnt + tc.cnt_slot, 1);
    }
  }
}

// TR = swap-AB ("transposed") form. The two forms are separate instantiations so that a decode-sized launch does not carry
// the plain form's epilogues in its instruction stream (and vice versa).
// NORM (swap-AB, batch <= 64): four more warps apply RMSNorm to every landed activation k-slice in shared memory before
// the MMA warp may read it (GemmParams::nl_*), so the GEMM consumes the raw residual stream and no norm kernel runs.
template <int BN, bool TR, bool NORM = false>
__global__ void __launch_bounds__(NORM ? NUM_THREADS_NORM : NUM_THREADS, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                         const __grid_constant__ CUtensorMap tmap_pf, const __grid_constant__ CUtensorMap tmap_out,
                         const GemmParams p) {
  static_assert(!NORM || (TR && BN <= 64), "norm-on-load: swap-AB form, batch tile <= 64");
  using C = Cfg<BN, TR, NORM>;
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B tiles need 1024-byte aligned bases
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::BARS_OFF);
  uint64_t* full_bar = bars;                     // [STAGES]  TMA -> MMA (NORM: TMA -> norm warps)
  uint64_t* empty_bar = bars + C::STAGES;        // [STAGES]  MMA -> TMA
  uint64_t* acc_full = bars + 2 * C::STAGES;     // [2]       MMA -> epilogue
  uint64_t* acc_empty = bars + 2 * C::STAGES + 2;  // [2]     epilogue -> MMA
  uint64_t* xf_bar = bars + 2 * C::STAGES + 4;   // [STAGES]  norm warps -> MMA (NORM only)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + C::N_BARS);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&acc_full[a], 1);
      mbar_init(&acc_empty[a], NUM_EPI_THREADS);
    }
    if constexpr (NORM)
      for (int s = 0; s < C::STAGES; ++s) mbar_init(&xf_bar[s], 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, C::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (NORM && warp >= 10) {
    // ===================== RMSNorm warps (10..13): activation k-slices, in place =====================
    const int tt = threadIdx.x - NUM_THREADS;                                 // 0..127
    const int nw = warp - 10;                                                 // this warp takes k-blocks nw, nw + 4, ...
    float* s_rstd = reinterpret_cast<float*>(smem + C::BARS_OFF + 256);       // [64] row scales, then [64] scratch
    grid_dep_wait();                                                          // nl_sumsq comes from the preceding kernels
    {
      // rstd[row] from the per-slab sums of squares: two threads per row, each sums one half of the slabs in slab
      // order (loads issued sixteen at a time), halves combined lower + upper: fixed order -> bit-reproducible
      const int row = tt & 63, half = tt >> 6;
      float acc = 0.f;
      if (row < p.N) {
        const int mid = p.nl_slabs >> 1;
        const int s0 = half ? mid : 0, s1 = half ? p.nl_slabs : mid;
        const float* src = p.nl_sumsq + row;
        for (int sl = s0; sl < s1; sl += 16) {
          float t[16];
#pragma unroll
          for (int u = 0; u < 16; ++u) t[u] = (sl + u < s1) ? __ldcg(src + (size_t)(sl + u) * p.nl_ld) : 0.f;
#pragma unroll
          for (int u = 0; u < 16; ++u) acc += t[u];
        }
      }
      if (half) s_rstd[64 + row] = acc;
      named_bar_sync(3, 128);
      if (!half) s_rstd[row] = row < p.N ? rsqrtf((acc + s_rstd[64 + row]) / (float)p.K + p.nl_eps) : 0.f;
      named_bar_sync(3, 128);
    }
    // Each warp owns whole k-slices (every fourth one), so four slices are in flight and the wait / fence latency of one
    // overlaps the arithmetic of the others. Lane -> 16-byte chunk position `cpos` of rows r0, r0 + 4, r0 + 8, ...:
    // (row & 7) alternates between r0 and r0 + 4, so the logical k-chunk behind the 128-byte swizzle (cpos ^ (row & 7))
    // and with it the gamma slice alternates between two values per lane.
    const int cpos = lane & 7, r0 = lane >> 3;
    const __nv_bfloat16* gam0 = static_cast<const __nv_bfloat16*>(p.nl_gamma) + (cpos ^ r0) * 8;
    const __nv_bfloat16* gam1 = static_cast<const __nv_bfloat16*>(p.nl_gamma) + (cpos ^ (r0 + 4)) * 8;
    float rs[BN / 4];
#pragma unroll
    for (int i = 0; i < BN / 4; ++i) rs[i] = s_rstd[r0 + 4 * i];
    int gidx = 0;                                                             // k-slices of this CTA, over all its items
    TileCoord tc;
    for (int it = 0; get_work(p, it, tc); ++it) {
      for (int kb = tc.kb_begin; kb < tc.kb_end; ++kb, ++gidx) {
        if ((gidx & 3) != nw) continue;
        const int stage = gidx % C::STAGES;
        const uint32_t phase = (uint32_t)(gidx / C::STAGES) & 1u;
        const uint4 gq0 = __ldg(reinterpret_cast<const uint4*>(gam0 + (size_t)kb * BK));   // in flight across the wait
        const uint4 gq1 = __ldg(reinterpret_cast<const uint4*>(gam1 + (size_t)kb * BK));
        mbar_wait(&full_bar[stage], phase);
        uint8_t* sb = smem + stage * C::STAGE + C::STAGE_A + r0 * 128 + cpos * 16;
        float g0[8], g1[8];
        bf16x8_unpack(gq0, g0);
        bf16x8_unpack(gq1, g1);
#pragma unroll
        for (int i = 0; i < BN / 4; ++i) {
          uint4* ptr = reinterpret_cast<uint4*>(sb + i * 4 * 128);
          float x[8], o[8];
          bf16x8_unpack(*ptr, x);
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = ((i & 1) ? g1[j] : g0[j]) * bf16_round(x[j] * rs[i]);
          uint4 q;
          q.x = pack_bf16x2(o[0], o[1]); q.y = pack_bf16x2(o[2], o[3]);
          q.z = pack_bf16x2(o[4], o[5]); q.w = pack_bf16x2(o[6], o[7]);
          *ptr = q;
        }
        fence_proxy_async();               // generic-proxy writes -> visible to the tensor core's (async proxy) reads
        __syncwarp();
        if (lane == 0) mbar_arrive(&xf_bar[stage]);
      }
    }
  } else if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      // PDL prologue: the WEIGHT operand never depends on the preceding kernel, so the first ring pass of weight
      // tiles is requested before waiting for it; only the activation operand has to wait. At decode batch sizes this
      // keeps HBM busy across kernel boundaries.
      int pre = 0;
      TileCoord tc;
      if (get_work(p, 0, tc)) {
        pre = min(C::STAGES, tc.kb_end - tc.kb_begin);
        for (int i = 0; i < pre; ++i) {
          uint8_t* sa = smem + i * C::STAGE;
          mbar_arrive_expect_tx(&full_bar[i], C::STAGE);
          if (TR) tma_load_2d_hint(sa, &tmap_a, &full_bar[i], (tc.kb_begin + i) * BK, tc.m * BM, p.hint_a);
          else tma_load_2d_hint(sa + C::STAGE_A, &tmap_b, &full_bar[i], (tc.kb_begin + i) * BK, tc.n * BN, p.hint_b);
        }
        if constexpr (TR) {
          // Weight streaming is latency x concurrency bound: the ring holds ~144 KB per SM, and at the power-capped clock
          // the decode loop inherits from the prefill the L2 / crossbar part of the latency stretches (the same launch
          // streams 18-24 % slower). Requesting the NEXT l2_ahead k-blocks into L2 keeps more HBM requests in flight than
          // shared memory can hold; the ring then refills from L2.
          for (int i = pre; i < min(pre + p.l2_ahead, tc.kb_end - tc.kb_begin); ++i)
            tma_prefetch_l2_2d(&tmap_a, (tc.kb_begin + i) * BK, tc.m * BM);
        }
        grid_dep_wait();
        for (int i = 0; i < pre; ++i) {
          uint8_t* sa = smem + i * C::STAGE;
          if (TR) tma_load_2d_hint(sa + C::STAGE_A, &tmap_b, &full_bar[i], (tc.kb_begin + i) * BK, tc.n * BN, p.hint_b);
          else tma_load_2d_hint(sa, &tmap_a, &full_bar[i], (tc.kb_begin + i) * BK, tc.m * BM, p.hint_a);
        }
        if (pre == C::STAGES) { stage = 0; phase = 1; } else { stage = pre; }
      } else {
        grid_dep_wait();
      }
      for (int it = 0; get_work(p, it, tc); ++it) {
        for (int kb = tc.kb_begin + (it == 0 ? pre : 0); kb < tc.kb_end; ++kb) {
          if constexpr (TR) {
            if (p.l2_ahead > 0 && kb + p.l2_ahead < tc.kb_end)       // keep the L2 lookahead window full
              tma_prefetch_l2_2d(&tmap_a, (kb + p.l2_ahead) * BK, tc.m * BM);
          }
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * C::STAGE;
          uint8_t* sb = sa + C::STAGE_A;
          mbar_arrive_expect_tx(&full_bar[stage], C::STAGE);
          tma_load_2d_hint(sa, &tmap_a, &full_bar[stage], kb * BK, tc.m * BM, p.hint_a);
          tma_load_2d_hint(sb, &tmap_b, &full_bar[stage], kb * BK, tc.n * BN, p.hint_b);
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
      }
      // Next-weight prefetch: this CTA's own stream is fully requested; keep HBM busy through the drain / launch gap.
      for (int t = blockIdx.x; t < p.pf_items; t += gridDim.x) {
        const int m = t / p.pf_split_k;
        const int s = t - m * p.pf_split_k;
        const int kb0 = (int)(((long long)s * p.pf_k_blocks) / p.pf_split_k);
        const int kb1 = (int)(((long long)(s + 1) * p.pf_k_blocks) / p.pf_split_k);
        const int n = min(p.pf_depth, kb1 - kb0);
        for (int i = 0; i < n; ++i) tma_prefetch_l2_2d(&tmap_pf, (kb0 + i) * BK, m * BM);
      }
      // PDL trigger, issued LATE (all loads of this CTA are in flight): dependents that became resident earlier would
      // only sit in griddepcontrol.wait next to a long-running GEMM.
      grid_dep_launch();
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      TileCoord tc;
      for (int it = 0; get_work(p, it, tc); ++it) {
        mbar_wait(&acc_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = tc.kb_begin; kb < tc.kb_end; ++kb) {
          mbar_wait(NORM ? &xf_bar[stage] : &full_bar[stage], phase);   // NORM: the slice has been normalised in place
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * C::STAGE);
          const uint64_t da = umma_smem_desc_sw128(sa);
          const uint64_t db = umma_smem_desc_sw128(sa + C::STAGE_A);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            // advance 32 bytes (= 16 bf16) along K inside the swizzle atom: +2 in 16-byte address units
            umma_bf16_ss(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb > tc.kb_begin || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);  // frees the smem slot when these MMAs retire
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&acc_full[acc]);  // accumulator complete -> epilogue
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue warps (2..9) =====================
    const int quad = warp & 3;          // TMEM lane quadrant this warp may access (hardware rule: warp id % 4)
    const int chunk0 = (warp - 2) >> 2;  // 0 or 1: the pair of warps of a quadrant interleave the column chunks
    int acc = 0;
    uint32_t acc_phase = 0;
    const int epi_tid = threadIdx.x - 64;  // 0..255
    TileCoord tc;
    // Swap-AB (decode-sized) launches execute their epilogue once per CTA, i.e. with a cold instruction cache, on the
    // critical path between two dependent kernels. Pass -1 runs the first item's epilogue "dry" (same instructions,
    // every memory access masked, TMEM contents ignored) while the main loop is still streaming weights, so the real
    // pass finds its code cached. The dry pass touches no global memory and therefore runs before griddepcontrol.wait.
    for (int it = (TR && p.epi_warm) ? -1 : 0;; ++it) {
      const bool dry = it < 0;
      if (dry) {
        // warm the path of the CTA's LAST item: that epilogue (plain tile or stream-K owner) is the one the next kernel
        // waits for; earlier items overlap the main loop anyway
        int n = 0;
        while (get_work(p, n, tc)) ++n;
        if (n == 0) continue;                  // no work at all for this CTA: it = 0 breaks
        get_work(p, n - 1, tc);
      } else {
        if (it == 0) grid_dep_wait();          // residual / output buffers may still be in use by the preceding kernel
        if (!get_work(p, it, tc)) break;
      }
      // the accumulator-ready wait happens inside (after a stream-K owner has requested its fix-up operands)
      epilogue_item<BN, TR>(p, tc, tmem_base, acc, quad, lane, chunk0, epi_tid, &tmap_out,
                        smem + C::STAGES * C::STAGE, dry, &acc_full[acc], acc_phase);
      if (!dry) {
        tc_fence_before();
        mbar_arrive(&acc_empty[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
    if (p.tma_store && lane == 0) tma_store_wait_all();   // the staging buffers die with the CTA
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------
// CTA-pair (cta_group::2) form of the plain GEMM: a cluster of two CTAs computes a 256 x 256 output tile with one UMMA
// of M = 256. Each CTA loads 128 rows of A and 128 of the 256 B rows per k-block (32 KB per stage instead of 48 KB, so
// six stages fit: 50 % more bytes in flight and half the B traffic per SM); the leader CTA issues the MMAs for the pair,
// both CTAs drain their own 128 accumulator rows. Barrier protocol (all barriers live at the same offsets in both CTAs):
//   full[s]      leader only : expect_tx of BOTH CTAs' bytes, both producers' TMA loads complete on it
//   empty[s]     both        : tcgen05.commit multicast when the MMAs that read stage s retire
//   acc_full[a]  both        : tcgen05.commit multicast when a tile's accumulator is complete
//   acc_empty[a] leader only : 2 x 256 epilogue threads (the peer arrives remotely)
// ------------------------------------------------------------------------------------------------
struct Cfg2 {
  static constexpr int BN = 256;
  static constexpr int STAGE_A = BM * BK * 2;          // this CTA's 128 rows of A
  static constexpr int STAGE_B = (BN / 2) * BK * 2;    // this CTA's 128 rows of B
  static constexpr int STAGE = STAGE_A + STAGE_B;      // 32 KB
  static constexpr int STAGES = 6;
  static constexpr int TMEM_COLS = 512;                // two 256-column accumulator buffers
  static constexpr int STORE_BYTES = 8 * 4096;
  static constexpr int BARS_OFF = STAGES * STAGE + STORE_BYTES;
  static constexpr int SMEM = BARS_OFF + 1024 + 256 + 1024;
};

// TR = swap-AB form (decode at batch 129..256: A = weights, the pair shares the 256-row activation tile, so a stage holds
// 16 KB of weights + 16 KB of activations instead of 16 + 32 and six stages = 96 KB of weights are in flight per SM
// instead of 64 KB). Work items are (256-row tile, k-split) pairs; the plain form always has split_k = 1.
template <bool TR>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_bf16_2cta_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                      const __grid_constant__ CUtensorMap tmap_out, const GemmParams p) {
  using C = Cfg2;
  constexpr int BN = C::BN;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::BARS_OFF);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + C::STAGES;
  uint64_t* acc_full = bars + 2 * C::STAGES;
  uint64_t* acc_empty = bars + 2 * C::STAGES + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * C::STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int rank = (int)cluster_ctarank();        // 0 = leader
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const int m_tiles = (p.M + 2 * BM - 1) / (2 * BM);     // 256-row tiles
  const int n_tiles = p.num_n_tiles;
  const int total = m_tiles * n_tiles * p.split_k;
  const int group = max(1, p.group_m / 2);
  // work item t -> (tile, k-split): k-block range [kb0, kb1) of the split
  auto split_of = [&](int t, int& mn, int& split, int& kb0, int& kb1) {
    mn = t / p.split_k;
    split = t - mn * p.split_k;
    kb0 = (int)(((long long)split * p.k_blocks) / p.split_k);
    kb1 = (int)(((long long)(split + 1) * p.k_blocks) / p.split_k);
  };
  auto tile_of = [&](int t, int& m, int& n) {             // grouped raster over 256-row tiles
    if (p.group_n > 0) {
      // grouped-N: a band of group_n weight panels stays L2-resident (evict-last) while ALL row tiles sweep past it, so
      // the weights are read from HBM once and the activations once per band. For K = 14336 (down_proj) a panel is
      // 7.3 MB: the grouped-M order touched the whole 117 MB weight matrix in every wave.
      const int per_band = p.group_n * m_tiles;
      const int g = t / per_band, r = t - g * per_band;
      const int first_n = g * p.group_n;
      const int gsz = min(n_tiles - first_n, p.group_n);
      n = first_n + r % gsz;
      m = r / gsz;
      return;
    }
    const int per_group = group * n_tiles;
    const int g = t / per_group, r = t - g * per_group;
    const int first_m = g * group;
    const int gsz = min(m_tiles - first_m, group);
    m = first_m + r % gsz;
    n = r / gsz;
  };
  // Item `it` of this PAIR: tc.m is the 256-row tile. Swap-AB form: the generic sequence (whole tiles round-robin, then
  // this pair's share of the stream-K tail; the host filled p.num_m_tiles / dp_items / sk_tiles in pair units).
  auto pair_work = [&](int it, TileCoord& c) -> bool {
    if constexpr (TR) {
      return get_work_bg(p, it, c, pair, n_pairs);
    } else {
      const int t = pair + it * n_pairs;
      if (t >= total) return false;
      int mn;
      split_of(t, mn, c.split, c.kb_begin, c.kb_end);
      tile_of(mn, c.m, c.n);
      c.kind = WORK_TILE; c.sk_tile = 0; c.first_cta = 0; c.cta_stride = 1; c.cnt_slot = 0;
      return true;
    }
  };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    tma_prefetch_desc(&tmap_out);
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&acc_full[a], 1);
      mbar_init(&acc_empty[a], 2 * NUM_EPI_THREADS);
    }
    fence_mbar_init();
  }
  cluster_sync_all();                              // both CTAs' barriers exist before any remote arrive / TMA signal
  if (warp == 1) {
    tmem_alloc_pair(tmem_slot, C::TMEM_COLS);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int pre = 0;
      if constexpr (TR) {
        // PDL prologue (see the single-CTA kernel): the first ring pass of WEIGHT tiles is requested before waiting for
        // the preceding kernel, the activation halves follow after the wait
        TileCoord t0;
        if (pair_work(0, t0)) {
          const int kb0 = t0.kb_begin, kb1 = t0.kb_end, m = t0.m, n = t0.n;
          pre = min(C::STAGES, kb1 - kb0);
          for (int i = 0; i < pre; ++i) {
            if (rank == 0) mbar_arrive_expect_tx(&full_bar[i], 2 * C::STAGE);
            tma_load_2d_pair(smem + i * C::STAGE, &tmap_a, &full_bar[i], (kb0 + i) * BK, (2 * m + rank) * BM, p.hint_a);
          }
          grid_dep_wait();
          for (int i = 0; i < pre; ++i)
            tma_load_2d_pair(smem + i * C::STAGE + C::STAGE_A, &tmap_b, &full_bar[i], (kb0 + i) * BK,
                             n * BN + rank * (BN / 2), p.hint_b);
          if (pre == C::STAGES) { stage = 0; phase = 1; } else { stage = pre; }
        } else {
          grid_dep_wait();
        }
      }
      TileCoord tw;
      for (int it = 0; pair_work(it, tw); ++it) {
        const int kb0 = tw.kb_begin, kb1 = tw.kb_end, m = tw.m, n = tw.n;
        for (int kb = kb0 + (it == 0 ? pre : 0); kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * C::STAGE;
          if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * C::STAGE);   // bytes of both CTAs
          tma_load_2d_pair(sa, &tmap_a, &full_bar[stage], kb * BK, (2 * m + rank) * BM, p.hint_a);
          tma_load_2d_pair(sa + C::STAGE_A, &tmap_b, &full_bar[stage], kb * BK, n * BN + rank * (BN / 2), p.hint_b);
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
      }
      if constexpr (TR) grid_dep_launch();   // late trigger: every load of this CTA is in flight
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(2 * BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      TileCoord tw;
      for (int it = 0; pair_work(it, tw); ++it) {
        const int kb0 = tw.kb_begin, kb1 = tw.kb_end;
        mbar_wait(&acc_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * C::STAGE);
          const uint64_t da = umma_smem_desc_sw128(sa);
          const uint64_t db = umma_smem_desc_sw128(sa + C::STAGE_A);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k)
            umma_bf16_ss_pair(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          umma_commit_pair(&empty_bar[stage]);   // both CTAs' slots
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit_pair(&acc_full[acc]);        // both CTAs' epilogues
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue warps (both CTAs): this CTA's 128 rows of the tile =====================
    const int quad = warp & 3;
    const int chunk0 = (warp - 2) >> 2;
    const int epi_tid = threadIdx.x - 64;
    if constexpr (TR) grid_dep_wait();   // output / partial buffers may still be in use by the preceding kernel
    int acc = 0;
    uint32_t acc_phase = 0;
    TileCoord tc;
    for (int it = 0; pair_work(it, tc); ++it) {
      // this CTA's half of the pair's tile; stream-K bookkeeping per rank: the contributors of a tile are the same-rank
      // CTAs of the preceding pairs, and every rank has its own arrival counter
      tc.m = 2 * tc.m + rank;
      tc.first_cta = 2 * tc.first_cta + rank;
      tc.cta_stride = 2;
      tc.cnt_slot = 2 * tc.sk_tile + rank;
      mbar_wait(&acc_full[acc], acc_phase);
      tc_fence_after();
      epilogue_item<BN, TR>(p, tc, tmem_base, acc, quad, lane, chunk0, epi_tid, &tmap_out, smem + C::STAGES * C::STAGE);
      tc_fence_before();
      mbar_arrive_leader(&acc_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (p.tma_store && lane == 0) tma_store_wait_all();
  }

  tc_fence_before();
  cluster_sync_all();            // the peer's shared memory / TMEM stay alive until the leader's last MMA has retired
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, C::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------
// fused chain kernel (see gemm.h: gemm_chain)
// ------------------------------------------------------------------------------------------------
struct ChainTmaps {
  CUtensorMap a[kMaxChainPhases];
  CUtensorMap b[kMaxChainPhases];
};
struct ChainProgram {
  int n_phases;
  int kind[kMaxChainPhases];
  GemmParams g[kMaxChainPhases];
  ChainNorm n[kMaxChainPhases];
  int* bar;  // [0] arrivals (monotonic within a launch), [1] CTAs that left the kernel (last one re-arms both)
  int l2_depth;               // k-blocks per CTA requested into L2 ahead of a phase barrier (beyond the smem ring)
  unsigned long long* trace;  // optional [grid][kMaxChainPhases][4] globaltimer stamps (tools/trace_chain.py), else null
};

__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void chain_stamp(const ChainProgram& prog, int ph, int slot) {
  if (prog.trace != nullptr) prog.trace[((size_t)blockIdx.x * kMaxChainPhases + ph) * 4 + slot] = global_ns();
}

// spin until `target` CTAs have arrived at the device-wide barrier counter
__device__ __forceinline__ void chain_barrier_wait(const int* bar, int target) {
  if (ld_acquire_gpu(bar) >= target) return;
  const long long t0 = clock64();
  while (ld_acquire_gpu(bar) < target) {
    if (clock64() - t0 > 20000000000LL) {
      printf("opus_b200: chain barrier wait timed out (block %d, target %d)\n", (int)blockIdx.x, target);
      __trap();
    }
  }
}

// split-K reduce + residual + RMSNorm of one row by the 256 epilogue threads (same rounding points as
// rmsnorm_bf16_kernel in bandwidth.cu, so the fused and the unfused decode paths agree bit for bit)
__device__ __forceinline__ void chain_norm_row(const ChainNorm& n, int row, int tid, float* red /* smem [8] */) {
  constexpr int MAXG = 2;  // 8-column groups per thread: cols <= 256 * 8 * 2 = 4096 per pass
  const size_t slice = (size_t)n.rows * n.cols;
  const __nv_bfloat16* res = static_cast<const __nv_bfloat16*>(n.residual);
  const __nv_bfloat16* wv = static_cast<const __nv_bfloat16*>(n.w);
  uint4 raw[MAXG], wq[MAXG], rq[MAXG];
  float4 pa[MAXG][4], pb[MAXG][4];
  // every load of the row (first four partial slices, residual, norm weight) is in flight before the first use
#pragma unroll
  for (int i = 0; i < MAXG; ++i) {
    const int c = (i * NUM_EPI_THREADS + tid) * 8;
    const bool ok = c < n.cols;
    const size_t off = (size_t)row * n.cols + (ok ? c : 0);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const bool pok = ok && u < n.n_partial;
      const float* pp = n.partial + (size_t)(pok ? u : 0) * slice + off;
      pa[i][u] = pok ? __ldcg(reinterpret_cast<const float4*>(pp)) : make_float4(0.f, 0.f, 0.f, 0.f);
      pb[i][u] = pok ? __ldcg(reinterpret_cast<const float4*>(pp + 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    rq[i] = (ok && res != nullptr) ? __ldcg(reinterpret_cast<const uint4*>(res + off)) : make_uint4(0, 0, 0, 0);
    wq[i] = ok ? *reinterpret_cast<const uint4*>(wv + c) : make_uint4(0, 0, 0, 0);
  }
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < MAXG; ++i) {
    const int c = (i * NUM_EPI_THREADS + tid) * 8;
    raw[i] = make_uint4(0, 0, 0, 0);
    if (c < n.cols) {
      const size_t off = (size_t)row * n.cols + c;
      float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int u = 0; u < 4; ++u) {   // left-to-right over the slices, like rmsnorm_bf16_kernel
        if (u < n.n_partial) {
          acc[0] += pa[i][u].x; acc[1] += pa[i][u].y; acc[2] += pa[i][u].z; acc[3] += pa[i][u].w;
          acc[4] += pb[i][u].x; acc[5] += pb[i][u].y; acc[6] += pb[i][u].z; acc[7] += pb[i][u].w;
        }
      }
      for (int u = 4; u < n.n_partial; ++u) {
        const float* pp = n.partial + (size_t)u * slice + off;
        const float4 a = __ldcg(reinterpret_cast<const float4*>(pp)), b = __ldcg(reinterpret_cast<const float4*>(pp + 4));
        acc[0] += a.x; acc[1] += a.y; acc[2] += a.z; acc[3] += a.w;
        acc[4] += b.x; acc[5] += b.y; acc[6] += b.z; acc[7] += b.w;
      }
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = bf16_round(acc[j]);
      if (res != nullptr) {
        const __nv_bfloat162* rb = reinterpret_cast<const __nv_bfloat162*>(&rq[i]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 rf = __bfloat1622float2(rb[j]);
          v[2 * j] = bf16_round(v[2 * j] + rf.x);
          v[2 * j + 1] = bf16_round(v[2 * j + 1] + rf.y);
        }
      }
      raw[i].x = pack_bf16x2(v[0], v[1]); raw[i].y = pack_bf16x2(v[2], v[3]);
      raw[i].z = pack_bf16x2(v[4], v[5]); raw[i].w = pack_bf16x2(v[6], v[7]);
      if (n.h_out != nullptr) *reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(n.h_out) + off) = raw[i];
#pragma unroll
      for (int j = 0; j < 8; ++j) sq += v[j] * v[j];
    }
  }
  sq = warp_sum(sq);
  if ((tid & 31) == 0) red[tid >> 5] = sq;
  named_bar_sync(1, NUM_EPI_THREADS);
  float tot = 0.f;
#pragma unroll
  for (int i = 0; i < NUM_EPI_THREADS / 32; ++i) tot += red[i];
  const float rstd = rsqrtf(tot / n.cols + n.eps);
#pragma unroll
  for (int i = 0; i < MAXG; ++i) {
    const int c = (i * NUM_EPI_THREADS + tid) * 8;
    if (c < n.cols) {
      const __nv_bfloat162* wb = reinterpret_cast<const __nv_bfloat162*>(&wq[i]);
      const __nv_bfloat162* hb = reinterpret_cast<const __nv_bfloat162*>(&raw[i]);
      uint4 o;
      uint32_t* op = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 hf = __bfloat1622float2(hb[j]);
        const float2 wf = __bfloat1622float2(wb[j]);
        op[j] = pack_bf16x2(wf.x * bf16_round(hf.x * rstd), wf.y * bf16_round(hf.y * rstd));
      }
      *reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(n.y) + (size_t)row * n.cols + c) = o;
    }
  }
  named_bar_sync(1, NUM_EPI_THREADS);  // `red` is reused by the next row
}

template <int BN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_chain_tcgen05_kernel(const __grid_constant__ ChainTmaps tm, const __grid_constant__ ChainProgram prog) {
  using C = Cfg<BN, true>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::BARS_OFF);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + C::STAGES;
  uint64_t* acc_full = bars + 2 * C::STAGES;
  uint64_t* acc_empty = bars + 2 * C::STAGES + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * C::STAGES + 4);
  float* red = reinterpret_cast<float*>(tmem_slot + 2);  // [8] RMSNorm block reduction

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int G = gridDim.x;

  if (warp == 0 && lane == 0) {
    for (int ph = 0; ph < prog.n_phases; ++ph) {
      if (prog.kind[ph] == CHAIN_GEMM) { tma_prefetch_desc(&tm.a[ph]); tma_prefetch_desc(&tm.b[ph]); }
    }
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&acc_full[a], 1);
      mbar_init(&acc_empty[a], NUM_EPI_THREADS);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, C::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer: one ring across all phases =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int ph = 0; ph < prog.n_phases; ++ph) {
        if (prog.kind[ph] != CHAIN_GEMM) continue;
        const GemmParams& p = prog.g[ph];
        TileCoord tc;
        int pre = 0;
        const bool have = get_work(p, 0, tc);
        // weights never depend on earlier phases: request the first ring pass of A tiles before the barrier
        if (have) {
          pre = min(C::STAGES, tc.kb_end - tc.kb_begin);
          int st = stage;
          uint32_t phs = phase;
          for (int i = 0; i < pre; ++i) {
            mbar_wait(&empty_bar[st], phs ^ 1);
            mbar_arrive_expect_tx(&full_bar[st], C::STAGE);
            tma_load_2d_hint(smem + st * C::STAGE, &tm.a[ph], &full_bar[st], (tc.kb_begin + i) * BK, tc.m * BM, p.hint_a);
            if (++st == C::STAGES) { st = 0; phs ^= 1; }
          }
        }
        // ... and, beyond the ring, ask for the following k-blocks of that item in L2: HBM keeps streaming while this
        // CTA sits at the barrier (norm phases, stragglers), and the phase then starts from L2-resident weights
        if (have) {
          const int kb_pf_end = min(tc.kb_end, tc.kb_begin + pre + prog.l2_depth);
          for (int kb = tc.kb_begin + pre; kb < kb_pf_end; ++kb) tma_prefetch_l2_2d(&tm.a[ph], kb * BK, tc.m * BM);
        }
        // activations: produced by the preceding kernel (phase 0) or by every CTA's previous phase
        if (ph == 0) grid_dep_wait();
        else chain_barrier_wait(prog.bar, G * ph);
        asm volatile("fence.proxy.async;" ::: "memory");  // generic-proxy writes of other CTAs -> this thread's TMA reads
        chain_stamp(prog, ph, 0);
        if (have) {
          for (int i = 0; i < pre; ++i) {
            tma_load_2d_hint(smem + stage * C::STAGE + C::STAGE_A, &tm.b[ph], &full_bar[stage], (tc.kb_begin + i) * BK,
                             tc.n * BN, p.hint_b);
            if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
          }
        }
        for (int it = 0; get_work(p, it, tc); ++it) {
          for (int kb = tc.kb_begin + (it == 0 ? pre : 0); kb < tc.kb_end; ++kb) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* sa = smem + stage * C::STAGE;
            mbar_arrive_expect_tx(&full_bar[stage], C::STAGE);
            tma_load_2d_hint(sa, &tm.a[ph], &full_bar[stage], kb * BK, tc.m * BM, p.hint_a);
            tma_load_2d_hint(sa + C::STAGE_A, &tm.b[ph], &full_bar[stage], kb * BK, tc.n * BN, p.hint_b);
            if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
      grid_dep_launch();
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int ph = 0; ph < prog.n_phases; ++ph) {
        if (prog.kind[ph] != CHAIN_GEMM) continue;
        const GemmParams& p = prog.g[ph];
        TileCoord tc;
        for (int it = 0; get_work(p, it, tc); ++it) {
          mbar_wait(&acc_empty[acc], acc_phase ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc * BN;
          for (int kb = tc.kb_begin; kb < tc.kb_end; ++kb) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            const uint32_t sa = smem_u32(smem + stage * C::STAGE);
            const uint64_t da = umma_smem_desc_sw128(sa);
            const uint64_t db = umma_smem_desc_sw128(sa + C::STAGE_A);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k)
              umma_bf16_ss(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb > tc.kb_begin || k > 0) ? 1u : 0u);
            umma_commit(&empty_bar[stage]);
            if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
          }
          umma_commit(&acc_full[acc]);
          if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
      }
    }
  } else {
    // ===================== epilogue warps (2..9): GEMM epilogues, norm phases, barrier arrivals =====================
    const int quad = warp & 3;
    const int chunk0 = (warp - 2) >> 2;
    const int epi_tid = threadIdx.x - 64;
    grid_dep_wait();
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int ph = 0; ph < prog.n_phases; ++ph) {
      if (prog.kind[ph] == CHAIN_GEMM) {
        const GemmParams& p = prog.g[ph];
        TileCoord tc;
        for (int it = 0; get_work(p, it, tc); ++it) {
          mbar_wait(&acc_full[acc], acc_phase);
          tc_fence_after();
          if (epi_tid == 0) chain_stamp(prog, ph, it == 0 ? 1 : 3);   // [3] = last item's accumulator ready
          epilogue_item<BN, true>(p, tc, tmem_base, acc, quad, lane, chunk0, epi_tid);
          tc_fence_before();
          mbar_arrive(&acc_empty[acc]);
          if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
      } else {
        // every earlier phase of every CTA must be complete (its partial sums are this phase's input)
        if (ph > 0) {
          if (epi_tid == 0) chain_barrier_wait(prog.bar, G * ph);
          named_bar_sync(1, NUM_EPI_THREADS);
        }
        if (epi_tid == 0) chain_stamp(prog, ph, 3);
        const ChainNorm& n = prog.n[ph];
        for (int row = blockIdx.x; row < n.rows; row += G) chain_norm_row(n, row, epi_tid, red);
      }
      // arrive at the device-wide barrier that closes this phase: the CTA barrier orders every epilogue thread's
      // stores before thread 0's gpu-scope fence + release (cumulative), so one fence per CTA suffices
      named_bar_sync(1, NUM_EPI_THREADS);
      if (epi_tid == 0) {
        __threadfence();
        asm volatile("fence.proxy.async;" ::: "memory");
        chain_stamp(prog, ph, 2);
        red_release_gpu_add(prog.bar, 1);
      }
    }
    // leave: the last CTA out re-arms the counters for the next launch (nobody waits on them any more)
    if (epi_tid == 0) {
      const int left = atomicAdd(prog.bar + 1, 1);
      if (left == G - 1) {
        chain_barrier_wait(prog.bar, G * prog.n_phases);
        prog.bar[0] = 0;
        prog.bar[1] = 0;
        __threadfence();
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  });
  return fn;
}

// [rows, cols] bf16 row-major, row stride ld elements; box = box_rows x 64 columns, 128-byte swizzle.
int make_tmap_bf16(CUtensorMap* tm, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return OPUS_ERR_DRIVER;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {BK, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? OPUS_OK : OPUS_ERR_TMAP;
}

template <int BN>
int launch(const GemmParams& p, const GemmArgs& pf, const void* A, int lda, const void* B, int ldb, int grid,
           cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(gemm_bf16_tcgen05_kernel<BN, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             Cfg<BN, false>::SMEM) != cudaSuccess ||
        cudaFuncSetAttribute(gemm_bf16_tcgen05_kernel<BN, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             Cfg<BN, true>::SMEM) != cudaSuccess)
      return OPUS_ERR_CUDA;
    configured = true;
  }
  CUtensorMap ta, tb;
  int rc = make_tmap_bf16(&ta, A, p.M, p.K, lda, BM);
  if (rc) return rc;
  rc = make_tmap_bf16(&tb, B, p.N, p.K, ldb, BN);
  if (rc) return rc;
  CUtensorMap tp = ta;  // unused unless p.pf_items > 0
  if (p.pf_items > 0) {
    rc = make_tmap_bf16(&tp, pf.pf_w, pf.pf_rows, pf.pf_K, pf.pf_K, BM);
    if (rc) return rc;
  }
  CUtensorMap to = ta;  // unused unless p.tma_store
  if (p.tma_store) {
    // box = 32 rows x 64 columns, one per epilogue warp (SwiGLU halves the output width)
    rc = make_tmap_bf16(&to, p.out, p.M, p.epi == EPI_SWIGLU ? p.N / 2 : p.N, p.ldo, 32);
    if (rc) return rc;
  }
  cudaError_t le;
  if (p.nl_sumsq != nullptr) {
    if constexpr (BN <= 64) {
      static bool norm_configured = false;
      if (!norm_configured) {
        if (cudaFuncSetAttribute(gemm_bf16_tcgen05_kernel<BN, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 Cfg<BN, true, true>::SMEM) != cudaSuccess)
          return OPUS_ERR_CUDA;
        norm_configured = true;
      }
      le = launch_pdl(true, gemm_bf16_tcgen05_kernel<BN, true, true>, dim3(grid), dim3(NUM_THREADS_NORM),
                      Cfg<BN, true, true>::SMEM, stream, ta, tb, tp, to, p);
    } else {
      return OPUS_ERR_ARG;   // prepare_gemm only admits batch tiles <= 64
    }
  } else {
    le = p.transposed ? launch_pdl(true, gemm_bf16_tcgen05_kernel<BN, true>, dim3(grid), dim3(NUM_THREADS), Cfg<BN, true>::SMEM, stream, ta, tb, tp, to, p)
                      : launch_pdl(false, gemm_bf16_tcgen05_kernel<BN, false>, dim3(grid), dim3(NUM_THREADS), Cfg<BN, false>::SMEM, stream, ta, tb, tp, to, p);
  }
  note_launch();
  return (le == cudaSuccess && cudaGetLastError() == cudaSuccess) ? OPUS_OK : OPUS_ERR_CUDA;
}

int g_num_sms = 0;
int num_sms() {
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  return g_num_sms;
}

// Stream-K fix-up workspace: one fp32 accumulator tile (128 x 256 max) per CTA plus one counter per tail tile.
// Allocated once per context on first use (19.4 MB); a context serialises its work on one stream, and consecutive
// launches on that stream reuse it in stream order.
bool ensure_sk_workspace() {
  Context& c = ctx();
  std::lock_guard<std::mutex> lk(c.sk_mu);   // not c.mu: the decode loop holds that one while it captures a step
  if (c.sk.state != 0) return c.sk.state > 0;
  if (!c.tun.streamk) { c.sk.state = -1; return false; }
  const size_t bytes = (size_t)num_sms() * BM * 256 * sizeof(float);
  if (cudaMalloc(&c.sk.ws, bytes) != cudaSuccess || cudaMalloc(&c.sk.cnt, 1024 * sizeof(int)) != cudaSuccess ||
      cudaMemset(c.sk.cnt, 0, 1024 * sizeof(int)) != cudaSuccess) {
    cudaGetLastError();  // e.g. first call inside a stream capture: stay on the plain schedule
    if (c.sk.ws) { cudaFree(c.sk.ws); c.sk.ws = nullptr; }
    if (c.sk.cnt) { cudaFree(c.sk.cnt); c.sk.cnt = nullptr; }
    return false;       // state stays 0: retried on the next eager call
  }
  c.sk.state = 1;
  return true;
}

}  // namespace

// The plain (activations = A) form keeps one summation order for every token row by default, so a token's result does
// not depend on where it sits in the batch (bitwise batch invariance); the stream-K tail is opt-in there. In the swap-AB
// form the tiles partition the FEATURES, every batch row is treated alike, and the tail is on by default.
void gemm_set_streamk_fill(int percent) { ctx().tun.streamk_fill = percent; }
void gemm_set_tma_store(int on) { ctx().tun.tma_store = on != 0; }
void gemm_set_streamk_plain(int on) { ctx().tun.streamk_plain = on; }

int gemm_pick_bn(int N, int transposed) {
  if (transposed) {
    if (N <= 32) return 32;
    if (N <= 64) return 64;
    if (N <= 128) return 128;
    return 256;
  }
  return (N % 256 == 0 || N > 1024) ? 256 : (N >= 128 ? 128 : 64);
}

int gemm_pick_split_k(int M, int N, int K, int bn) {
  const int tiles = ((M + BM - 1) / BM) * ((N + bn - 1) / bn);
  const int kb = (K + BK - 1) / BK;
  const int sms = num_sms();
  if (tiles >= sms / 2) return 1;
  int s = sms / tiles;  // fill the machine once
  if (s > kb / 4) s = kb / 4;  // keep >= 4 k-blocks (256 of K) per split
  if (s > 16) s = 16;
  return s < 1 ? 1 : s;
}

// Swap-AB launches at batch 257..512 (two 256-wide batch tiles) run on the CTA-pair kernel; their work items are
// (256-feature tile, batch tile, k-split). Pick the split with the shortest critical path: whole waves of pairs times the
// k-blocks of one item plus a per-item cost (ring fill, accumulator drain, one more fp32 partial slice to reduce).
int gemm_pick_split_k_wide(int M, int N, int K) {
  const int items1 = ((M + 2 * BM - 1) / (2 * BM)) * ((N + 255) / 256);
  const int kb = (K + BK - 1) / BK;
  const int pairs = num_sms() / 2;
  const int over = ctx().tun.wide_overhead;
  int best = 1;
  long best_cost = -1;
  for (int s = 1; s <= 8 && kb / s >= 4; ++s) {
    const int waves = (items1 * s + pairs - 1) / pairs;
    const long cost = (long)waves * ((kb + s - 1) / s + over + 2 * s);
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = s; }
  }
  return best;
}

size_t gemm_workspace_bytes(int M, int N, int split_k) { return (size_t)split_k * M * N * sizeof(float); }

// Validates `a` and fills the device parameter block (tiling, split-K, stream-K tail, prefetch hint). bn = tile width.
static int prepare_gemm(const GemmArgs& a, GemmParams& p, int& bn) {
  if (a.M <= 0 || a.N <= 0 || a.K <= 0) return OPUS_ERR_ARG;
  if ((a.lda % 8) || (a.ldb % 8) || (a.K % 8)) return OPUS_ERR_ARG;  // TMA: 16-byte aligned rows
  if ((reinterpret_cast<uintptr_t>(a.A) | reinterpret_cast<uintptr_t>(a.B)) & 15) return OPUS_ERR_ARG;
  if (!a.transposed && (a.N % 8)) return OPUS_ERR_ARG;
  if (a.epi == EPI_SWIGLU && ((a.transposed ? a.M : a.N) % 16)) return OPUS_ERR_ARG;
  if (a.transposed && (a.M % 2) && (a.epi == EPI_BF16 || a.epi == EPI_BF16_GELU || a.epi == EPI_BF16_RELU || a.epi == EPI_RES_BF16))
    return OPUS_ERR_ARG;  // paired 4-byte stores need an even feature count
  if (a.transposed && (a.ldo % 2) && a.epi != EPI_PARTIAL_F32 && a.epi != EPI_F32 && a.epi != EPI_RES_F32)
    return OPUS_ERR_ARG;
  if ((a.epi == EPI_RES_F32 || a.epi == EPI_RES_BF16) && a.residual == nullptr) return OPUS_ERR_ARG;
  if (a.epi == EPI_F32 && !a.transposed) return OPUS_ERR_ARG;

  bn = a.block_n > 0 ? a.block_n : gemm_pick_bn(a.N, a.transposed);
  p = GemmParams{};
  p.M = a.M; p.N = a.N; p.K = a.K;
  p.num_m_tiles = (a.M + BM - 1) / BM;
  p.num_n_tiles = (a.N + bn - 1) / bn;
  p.k_blocks = (a.K + BK - 1) / BK;
  p.split_k = a.split_k > 0 ? a.split_k : 1;
  if (p.split_k > p.k_blocks) p.split_k = p.k_blocks;
  if (p.split_k > 1 && a.epi != EPI_PARTIAL_F32 && !a.splitk_fixup) return OPUS_ERR_ARG;
  p.transposed = a.transposed;
  p.epi = a.epi;
  p.out = a.out; p.ldo = a.ldo;
  p.bias = a.bias;
  p.residual = a.residual; p.ldr = a.ldr;
  const Tunables& tun = ctx().tun;
  const int group_override = tun.group_m;
  // grouped-M raster: the A rows of one group (group_m x 128 x K) must stay L2-resident while the group's n-tiles are
  // walked; measured DRAM traffic per launch at M = 32768 (ncu): gate/up (K 4096) 4.96 GB at 16 -> 3.09 GB at 32, but
  // down (K 14336) 4.99 GB at 16 -> 5.65 GB at 32
  // swap-AB with more than one batch tile: walk the batch tiles of one weight panel back to back (n fastest), so the
  // panel is fetched from HBM once and re-read from L2
  p.group_m = a.transposed ? (p.num_n_tiles > 1 ? 1 : p.num_m_tiles)
                           : (group_override > 0 ? group_override : (a.K > 8192 ? 16 : 32));
  if (p.group_m > p.num_m_tiles) p.group_m = p.num_m_tiles;
  // CTA-pair kernel, long K: bands of weight panels instead of groups of row tiles (see tile_of); 0 = grouped-M
  p.group_n = (!a.transposed && a.K > 8192) ? tun.group_n : 0;
  // weights are streamed once in the swap-AB form; activations are re-read by every tile
  p.hint_a = a.transposed ? kCacheEvictFirst : kCacheEvictNormal;
  p.hint_b = a.transposed ? kCacheEvictLast : kCacheEvictNormal;
  {
    // plain form: the A panel of a raster group is re-read by every n-tile of the group, B is streamed once per group
    const int plain_hints = tun.plain_hints;
    if (!a.transposed && plain_hints == 1) { p.hint_a = kCacheEvictLast; p.hint_b = kCacheEvictFirst; }
    if (!a.transposed && plain_hints == 2) { p.hint_a = kCacheEvictLast; p.hint_b = kCacheEvictNormal; }
    if (!a.transposed && plain_hints == 3) { p.hint_a = kCacheEvictNormal; p.hint_b = kCacheEvictFirst; }
    if (!a.transposed && plain_hints == 0 && p.group_n > 0 && tun.group_n_hints) {
      p.hint_a = kCacheEvictFirst;   // a row panel is used by the tiles of one wave only
      p.hint_b = kCacheEvictLast;    // the band's weight panels serve every wave of the band
    }
  }

  if (a.pf_w != nullptr && a.pf_depth > 0 && a.pf_rows > 0 && a.pf_K > 0 && (a.pf_K % 8) == 0 &&
      (reinterpret_cast<uintptr_t>(a.pf_w) & 15) == 0) {
    p.pf_split_k = a.pf_split_k > 0 ? a.pf_split_k : 1;
    p.pf_k_blocks = (a.pf_K + BK - 1) / BK;
    p.pf_depth = a.pf_depth;
    const int items = ((a.pf_rows + BM - 1) / BM) * p.pf_split_k;
    p.pf_items = items < num_sms() ? items : num_sms();  // first wave of the next launch
  }

  p.tma_store = tun.tma_store && !a.transposed && (a.ldo % 8) == 0 && (reinterpret_cast<uintptr_t>(a.out) & 15) == 0 &&
                ((bn >= 64 && (a.epi == EPI_BF16 || a.epi == EPI_BF16_GELU || a.epi == EPI_BF16_RELU)) ||
                 (bn == 256 && a.epi == EPI_SWIGLU && (a.N % 128) == 0));
  if (a.rl_cos != nullptr) {
    if (!(p.tma_store && a.epi == EPI_BF16 && bn == 256 && a.rl_pos != nullptr && a.rl_sin != nullptr && a.rl_bs > 0 &&
          a.N == (a.rl_hq + 2 * a.rl_hkv) * 128))
      return OPUS_ERR_ARG;   // callers check gemm_fuses_rope() first
    p.rl_pos = a.rl_pos; p.rl_slot = a.rl_slot; p.rl_cos = a.rl_cos; p.rl_sin = a.rl_sin;
    p.rl_kcache = a.rl_kcache; p.rl_vcache = a.rl_vcache;
    p.rl_hq = a.rl_hq; p.rl_hkv = a.rl_hkv; p.rl_bs = a.rl_bs;
  }
  if (a.rope_cos != nullptr) {
    if (!(p.tma_store && a.epi == EPI_BF16 && a.rope_pos != nullptr && a.rope_sin != nullptr && (a.rope_cols % 64) == 0 &&
          (a.rope_q_cols % 64) == 0))
      return OPUS_ERR_ARG;   // callers check gemm_fuses_rope() first
    p.rope_pos = a.rope_pos; p.rope_cos = a.rope_cos; p.rope_sin = a.rope_sin;
    p.rope_cols = a.rope_cols; p.rope_q_cols = a.rope_q_cols; p.rope_q_scale = a.rope_q_scale;
  }

  if (a.sumsq_out != nullptr) {
    if (!a.transposed || a.epi != EPI_RES_BF16 || a.sumsq_ld < a.N || (a.M % 32)) return OPUS_ERR_ARG;
    p.sumsq_out = a.sumsq_out; p.sumsq_ld = a.sumsq_ld;
  }
  if (a.norm_sumsq != nullptr) {
    if (!a.transposed || bn > 64 || (a.K % BK) || a.norm_gamma == nullptr || a.norm_slabs <= 0 || a.norm_ld < a.N ||
        (reinterpret_cast<uintptr_t>(a.norm_gamma) & 15))
      return OPUS_ERR_ARG;
    p.nl_sumsq = a.norm_sumsq; p.nl_slabs = a.norm_slabs; p.nl_ld = a.norm_ld;
    p.nl_gamma = a.norm_gamma; p.nl_eps = a.norm_eps;
  }
  const int tiles = p.num_m_tiles * p.num_n_tiles * p.split_k;
  p.dp_items = tiles;
  p.sk_tiles = 0;
  p.epi_warm = a.transposed ? tun.epi_warm : 0;
  p.l2_ahead = a.transposed ? tun.l2_ahead : 0;
  if (a.splitk_fixup && p.split_k > 1) {
    // the CTAs of one tile wait on each other: every work item needs its own resident CTA (one wave)
    if (!a.transposed || tiles > num_sms() || p.num_m_tiles * p.num_n_tiles > 1000 || a.epi == EPI_PARTIAL_F32 ||
        !ensure_sk_workspace())
      return OPUS_ERR_ARG;
    p.sk_fix = 1;
    p.sk_ws = ctx().sk.ws;
    p.sk_cnt = ctx().sk.cnt;
  }
  // wave quantisation: a last wave that fills only part of the machine is cut along K over all CTAs instead
  const int rem = tiles % num_sms();
  const bool sk_ready = ensure_sk_workspace();  // also on launches that do not need it: never first inside a capture
  if (p.split_k == 1 && (a.transposed || tun.streamk_plain || a.streamk_tail) && tiles > num_sms() && rem != 0 &&
      rem * 100 <= tun.streamk_fill * num_sms() && a.epi != EPI_PARTIAL_F32 && sk_ready) {
    p.sk_tiles = rem;
    p.dp_items = tiles - rem;
    p.sk_ws = ctx().sk.ws;
    p.sk_cnt = ctx().sk.cnt;
    // a tail of a few tiles (gate/up at batch 512: 448 tiles = 3 waves + 4): 8 pieces per tile on the first CTAs instead of
    // a sliver on every CTA, so a tile's owner adds 7 accumulator dumps, not 36
    if (!a.transposed && rem * 10 <= num_sms()) p.sk_share = rem * 8;
  }
  return OPUS_OK;
}

bool gemm_fuses_rope(const GemmArgs& a) {
  GemmParams p;
  int bn = 0;
  GemmArgs probe = a;
  probe.rope_cos = nullptr;
  probe.rl_cos = nullptr;
  if (prepare_gemm(probe, p, bn) != OPUS_OK) return false;
  if (a.rl_cos != nullptr)
    return p.tma_store && a.epi == EPI_BF16 && bn == 256 && a.rl_pos != nullptr && a.rl_sin != nullptr && a.rl_bs > 0 &&
           a.N == (a.rl_hq + 2 * a.rl_hkv) * 128;
  return p.tma_store && a.epi == EPI_BF16 && a.rope_pos != nullptr && a.rope_cos != nullptr && a.rope_sin != nullptr &&
         (a.rope_cols % 64) == 0 && (a.rope_q_cols % 64) == 0;
}

namespace {

// Stream-K tail of the CTA-pair swap-AB kernel: a tail of a few tiles (<= 10 % of the pairs, e.g. gate/up at batch 512:
// 224 items = 3 waves + 2) is always cut along K, since its fix-up traffic is a couple of accumulators; larger tails only
// with the tunable pair_streamk (a 256-wide accumulator dump is 128 KB per CTA, measured slower at batch 256).
bool pair_streamk_takes(int rem, int max_pairs) {
  if (rem == 0) return false;
  if (rem * 10 <= max_pairs) return true;
  return ctx().tun.pair_streamk && rem * 100 <= ctx().tun.streamk_fill * max_pairs;
}

template <bool TR>
int launch_2cta(const GemmParams& p_in, const GemmArgs& a, cudaStream_t stream) {
  GemmParams p = p_in;
  if constexpr (TR) {
    // work partition in PAIR units: 256-feature tiles, whole (tile, k-split) items round-robin over the pairs, and a
    // partial last wave cut along K over all pairs (stream-K; same fix-up protocol as the single-CTA kernel, per rank)
    const int max_pairs = num_sms() / 2;
    p.num_m_tiles = (p.M + 2 * BM - 1) / (2 * BM);
    p.group_m = p.num_n_tiles > 1 ? 1 : p.num_m_tiles;
    const int items = p.num_m_tiles * p.num_n_tiles * p.split_k;
    p.dp_items = items;
    p.sk_tiles = 0;
    const int rem = items % max_pairs;
    // opt-in (tunable pair_streamk): at a 256-wide batch tile a dumped accumulator is 128 KB per CTA, and the fix-up
    // traffic costs more than the partial wave it removes (gate/up at batch 256: 6.17 ms per step with two uneven
    // waves, 6.27 with the tail cut along K)
    if (pair_streamk_takes(rem, max_pairs) && p.split_k == 1 && items > max_pairs && rem != 0 && ctx().tun.streamk &&
        a.epi != EPI_PARTIAL_F32 && ensure_sk_workspace()) {
      p.sk_tiles = rem;
      p.dp_items = items - rem;
      p.sk_ws = ctx().sk.ws;
      p.sk_cnt = ctx().sk.cnt;
      // a tail of a few tiles: 8 pieces per tile on the first pairs instead of one sliver on every pair (the tile's owner
      // adds the other pieces' 128 KB accumulator dumps one after the other)
      if (rem * 10 <= max_pairs && !ctx().tun.pair_streamk) p.sk_share = rem * 8 < max_pairs ? rem * 8 : max_pairs;
    }
  }
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(gemm_bf16_2cta_kernel<TR>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg2::SMEM) != cudaSuccess)
      return OPUS_ERR_CUDA;
    configured = true;
  }
  CUtensorMap ta, tb, to;
  int rc = make_tmap_bf16(&ta, a.A, p.M, p.K, a.lda, BM);
  if (rc) return rc;
  rc = make_tmap_bf16(&tb, a.B, p.N, p.K, a.ldb, Cfg2::BN / 2);
  if (rc) return rc;
  to = ta;
  if (p.tma_store) {
    rc = make_tmap_bf16(&to, p.out, p.M, p.epi == EPI_SWIGLU ? p.N / 2 : p.N, p.ldo, 32);
    if (rc) return rc;
  }
  const int items = ((p.M + 2 * BM - 1) / (2 * BM)) * p.num_n_tiles * p.split_k;
  const int max_pairs = num_sms() / 2;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * (TR && items < max_pairs ? items : max_pairs));
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = Cfg2::SMEM;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (TR && pdl_for(true)) ? 2 : 1;   // decode-sized launches chain through PDL like the single-CTA kernel
  const cudaError_t le = cudaLaunchKernelEx(&cfg, gemm_bf16_2cta_kernel<TR>, ta, tb, to, p);
  note_launch();
  return (le == cudaSuccess && cudaGetLastError() == cudaSuccess) ? OPUS_OK : OPUS_ERR_CUDA;
}

// swap-AB launches the pair kernel takes (tunable "gemm_2cta_tr", env OPUS_GEMM_2CTA_TR; default on): batch tile of 256
// (129..256 rows) whose (256-feature tile, k-split) items fill the 74 pairs to >= 85 % of whole waves. The stream-K tail
// stays with the single-CTA kernel (decode gate/up: 112 pair tiles would be 1.5 waves).
bool pair_takes_transposed(const GemmArgs& a, const GemmParams& p, int bn) {
  const int g_2cta_tr = ctx().tun.gemm_2cta_tr;
  if (!g_2cta_tr || !a.transposed || bn != 256 || a.block_n != 0 || a.N <= 128 || a.N > 512) return false;
  if (a.epi == EPI_SWIGLU && g_2cta_tr < 2) return false;     // tunable gemm_2cta_tr = 1: gate/up stays on the single-CTA kernel
  if (a.splitk_fixup || a.sumsq_out != nullptr || a.norm_sumsq != nullptr || a.pf_w != nullptr) return false;
  const int max_pairs = num_sms() / 2;
  const int items = ((p.M + 2 * BM - 1) / (2 * BM)) * p.num_n_tiles * p.split_k;
  if (items < max_pairs) return items * 100 >= 85 * max_pairs;   // a single, well filled wave
  // several waves: whole waves, or a partial last wave that the stream-K tail spreads over all pairs
  const int rem = items % max_pairs;
  const bool sk_ok = pair_streamk_takes(rem, max_pairs) && p.split_k == 1 && a.epi != EPI_PARTIAL_F32 && ctx().tun.streamk;
  const int waves = (items + max_pairs - 1) / max_pairs;
  return rem == 0 || sk_ok || items * 100 >= (a.epi == EPI_SWIGLU ? 75 : 85) * waves * max_pairs;
}
}  // namespace

void gemm_set_2cta_tr(int on) { ctx().tun.gemm_2cta_tr = on < 0 ? 0 : (on > 2 ? 2 : on); }
void gemm_set_2cta(int on) { ctx().tun.gemm_2cta = on < 0 ? 0 : (on > 2 ? 2 : on); }

// D = epi(A * B^T). See gemm.h for the contract.
int gemm_bf16(const GemmArgs& a, cudaStream_t stream) {
  GemmParams p;
  int bn = 0;
  const int rc = prepare_gemm(a, p, bn);
  if (rc != OPUS_OK) return rc;
  const int g_2cta = ctx().tun.gemm_2cta;
  // 1 = every eligible launch, 2 = every eligible launch except the SwiGLU epilogue (measured slower there)
  if (g_2cta && !a.transposed && bn == 256 && p.split_k == 1 && p.sk_tiles == 0 && a.M >= 1024 && a.block_n == 0 &&
      !(g_2cta == 2 && a.epi == EPI_SWIGLU))
    return launch_2cta<false>(p, a, stream);
  if (pair_takes_transposed(a, p, bn)) return launch_2cta<true>(p, a, stream);   // re-partitions the work in pair units
  const int tiles = p.num_m_tiles * p.num_n_tiles * p.split_k;
  const int grid = tiles < num_sms() ? tiles : num_sms();
  switch (bn) {
    case 32: return launch<32>(p, a, a.A, a.lda, a.B, a.ldb, grid, stream);
    case 64: return launch<64>(p, a, a.A, a.lda, a.B, a.ldb, grid, stream);
    case 128: return launch<128>(p, a, a.A, a.lda, a.B, a.ldb, grid, stream);
    case 256: return launch<256>(p, a, a.A, a.lda, a.B, a.ldb, grid, stream);
    default: return OPUS_ERR_ARG;
  }
}

namespace {
template <int BN>
int launch_chain(const ChainTmaps& tm, const ChainProgram& prog, cudaStream_t stream) {
  using C = Cfg<BN, true>;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(gemm_chain_tcgen05_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM) !=
        cudaSuccess)
      return OPUS_ERR_CUDA;
    configured = true;
  }
  // one CTA per SM: the device-wide barriers need every CTA resident (C::SMEM > half an SM, so never two per SM)
  const cudaError_t le = launch_pdl(true, gemm_chain_tcgen05_kernel<BN>, dim3(num_sms()), dim3(NUM_THREADS), C::SMEM,
                                    stream, tm, prog);
  note_launch();
  return (le == cudaSuccess && cudaGetLastError() == cudaSuccess) ? OPUS_OK : OPUS_ERR_CUDA;
}
}  // namespace

// measured: prefetching beyond the ring only adds traffic (tools/bench_decode.py), so the default depth is 0
void gemm_set_chain_l2_depth(int kblocks) { ctx().tun.chain_l2_depth = kblocks < 0 ? 0 : kblocks; }
// enable != 0: allocate the stamp buffer (once) and record every following chain launch (the last one wins);
// out != nullptr: copy [num_sms][kMaxChainPhases][4] stamps to the host (synchronises the device). Returns the SM count.
int gemm_chain_trace(int enable, unsigned long long* out, int cap_words) {
  const int words = num_sms() * kMaxChainPhases * 4;
  unsigned long long*& g_chain_trace = ctx().chain_trace;
  if (enable && g_chain_trace == nullptr) {
    if (cudaMalloc(&g_chain_trace, words * sizeof(unsigned long long)) != cudaSuccess) return OPUS_ERR_CUDA;
    cudaMemset(g_chain_trace, 0, words * sizeof(unsigned long long));
  }
  if (out != nullptr && g_chain_trace != nullptr) {
    cudaDeviceSynchronize();
    const int n = cap_words < words ? cap_words : words;
    if (cudaMemcpy(out, g_chain_trace, n * sizeof(unsigned long long), cudaMemcpyDeviceToHost) != cudaSuccess)
      return OPUS_ERR_CUDA;
  }
  if (!enable && g_chain_trace != nullptr && out == nullptr) { cudaFree(g_chain_trace); g_chain_trace = nullptr; }
  return num_sms();
}

int gemm_chain(const ChainPhase* phases, int n_phases, cudaStream_t stream) {
  if (phases == nullptr || n_phases <= 0 || n_phases > kMaxChainPhases) return OPUS_ERR_ARG;
  if (!ensure_sk_workspace()) return OPUS_ERR_CUDA;
  ChainTmaps tm;
  ChainProgram prog{};
  prog.n_phases = n_phases;
  prog.bar = ctx().sk.cnt + 1000;
  prog.trace = ctx().chain_trace;
  prog.l2_depth = ctx().tun.chain_l2_depth;
  int bn_all = 0;
  for (int i = 0; i < n_phases; ++i) {
    prog.kind[i] = phases[i].kind;
    if (phases[i].kind == CHAIN_NORM) {
      const ChainNorm& n = phases[i].norm;
      if (n.partial == nullptr || n.n_partial <= 0 || n.y == nullptr || n.w == nullptr || n.rows <= 0 ||
          n.cols <= 0 || (n.cols % 8) || n.cols > NUM_EPI_THREADS * 8 * 2)
        return OPUS_ERR_ARG;
      prog.n[i] = n;
      continue;
    }
    if (phases[i].kind != CHAIN_GEMM) return OPUS_ERR_ARG;
    const GemmArgs& a = phases[i].gemm;
    if (!a.transposed || a.pf_w != nullptr) return OPUS_ERR_ARG;
    int bn = 0;
    int rc = prepare_gemm(a, prog.g[i], bn);
    if (rc != OPUS_OK) return rc;
    if (bn_all == 0) bn_all = bn;
    if (bn != bn_all) return OPUS_ERR_ARG;  // one batch size (tile width) for the whole chain
    rc = make_tmap_bf16(&tm.a[i], a.A, a.M, a.K, a.lda, BM);
    if (rc) return rc;
    rc = make_tmap_bf16(&tm.b[i], a.B, a.N, a.K, a.ldb, bn);
    if (rc) return rc;
  }
  if (bn_all == 0) return OPUS_ERR_ARG;
  for (int i = 0; i < n_phases; ++i)   // unused slots still need valid descriptors
    if (prog.kind[i] != CHAIN_GEMM) { tm.a[i] = tm.a[0]; tm.b[i] = tm.b[0]; }
  if (prog.kind[0] != CHAIN_GEMM) {
    // slot 0 must hold a real map for the copies above
    for (int i = 0; i < n_phases; ++i)
      if (prog.kind[i] == CHAIN_GEMM) { tm.a[0] = tm.a[i]; tm.b[0] = tm.b[i]; break; }
    for (int i = 1; i < n_phases; ++i)
      if (prog.kind[i] != CHAIN_GEMM) { tm.a[i] = tm.a[0]; tm.b[i] = tm.b[0]; }
  }
  for (int i = n_phases; i < kMaxChainPhases; ++i) { tm.a[i] = tm.a[0]; tm.b[i] = tm.b[0]; }
  switch (bn_all) {
    case 32: return launch_chain<32>(tm, prog, stream);
    case 64: return launch_chain<64>(tm, prog, stream);
    case 128: return launch_chain<128>(tm, prog, stream);
    case 256: return launch_chain<256>(tm, prog, stream);
    default: return OPUS_ERR_ARG;
  }
}

}  // namespace opus
