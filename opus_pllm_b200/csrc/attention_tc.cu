// Flash attention forward on tcgen05 tensor cores (variable-length packed sequences; bidirectional or causal GQA).
//
// One CTA = one 128-row query tile of one (sequence, head). Warp roles:
//   warp 0     TMA producer: Q tile once, then K/V blocks of 128 keys into a 2-stage ring (SWIZZLE_128B panels of 64 cols)
//   warp 1     MMA issuer (one thread) + TMEM owner:  S = Q K^T  (UMMA 128x128x16, both operands K-major from smem)
//                                                     O_blk = P V (UMMA 128xDx16, A = P K-major, B = V MN-major)
//   warps 2-5  softmax: thread = query row (TMEM lane), so row max / row sum need no shuffles. S is read twice from TMEM
//              (max pass, exp pass), P is written to shared memory in the UMMA K-major swizzled layout, O is kept in
//              REGISTERS (fp32, D values per thread) and updated as O = O * corr + O_blk after every block.
// TMEM: S [128 lanes x 128 cols] + O_blk [128 x D]  (256 columns allocated).
// Nothing T x T is materialised; K/V rows beyond the sequence end (next packed sequence / OOB zero fill) are masked in S.
#include "common.h"
#include "kernels.h"
#include "launch.cuh"
#include "ptx.cuh"

#include <cuda.h>
#include <cstdlib>
#include <mutex>

namespace opus {

namespace {

constexpr int TBM = 128;   // query rows per CTA
constexpr int TBN = 128;   // keys per block
constexpr int PANEL = 128 * 128;  // bytes of one [128 rows x 64 bf16] swizzled panel
constexpr int TC_THREADS = 192;

struct AttnTcParams {
  const int* cu_seqlens;
  __nv_bfloat16* o;
  int ldo;
  int group;          // q heads per kv head
  float scale_log2;   // softmax scale * log2(e)
};

template <int D>
struct TcCfg {
  static constexpr int NP = D / 64;                       // panels per operand tile
  static constexpr int Q_BYTES = NP * PANEL;
  static constexpr int KV_BYTES = NP * PANEL;             // K (or V) block
  static constexpr int P_BYTES = (TBN / 64) * PANEL;
  static constexpr int SMEM = Q_BYTES + 2 * 2 * KV_BYTES + P_BYTES + 1024 + 256;
  static constexpr int TMEM_COLS = 256;                   // S: 128, O_blk: D (<= 128)
};

// instruction descriptor with selectable B major-ness (bit 16: 1 = MN-major)
__host__ __device__ constexpr uint32_t idesc_bf16(uint32_t M, uint32_t N, uint32_t b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (b_mn_major << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

// MN-major operand (rows = K index, 64 contiguous MN elements per 128-byte row, SWIZZLE_128B):
// LBO = byte distance between 64-element MN panels, SBO = byte distance between 8-row (K) groups.
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

template <int D, bool CAUSAL>
__global__ void __launch_bounds__(TC_THREADS, 1)
attn_fwd_tcgen05_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                        const __grid_constant__ CUtensorMap tm_v, const AttnTcParams p) {
  using C = TcCfg<D>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + C::Q_BYTES;            // [2 stages][KV_BYTES]
  uint8_t* sV = sK + 2 * C::KV_BYTES;       // [2 stages][KV_BYTES]
  uint8_t* sP = sV + 2 * C::KV_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + C::P_BYTES);
  uint64_t* q_full = bars;          // 1
  uint64_t* kv_full = bars + 1;     // [2]
  uint64_t* kv_empty = bars + 3;    // [2]
  uint64_t* s_full = bars + 5;
  uint64_t* p_full = bars + 6;
  uint64_t* o_full = bars + 7;
  uint64_t* o_empty = bars + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.z, h = blockIdx.y;
  const int seq_start = p.cu_seqlens[b];
  const int len = p.cu_seqlens[b + 1] - seq_start;
  const int mblk = CAUSAL ? (gridDim.x - 1 - blockIdx.x) : blockIdx.x;
  const int q0 = mblk * TBM;
  if (q0 >= len) return;
  const int kvh = h / p.group;
  const int kv_end = CAUSAL ? min(len, q0 + TBM) : len;
  const int n_blocks = (kv_end + TBN - 1) / TBN;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_q); tma_prefetch_desc(&tm_k); tma_prefetch_desc(&tm_v);
    mbar_init(q_full, 1);
    for (int s = 0; s < 2; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1); }
    mbar_init(s_full, 1);
    mbar_init(p_full, 128);
    mbar_init(o_full, 1);
    mbar_init(o_empty, 128);
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
  const uint32_t tmem_s = tmem_base;          // columns [0, 128)
  const uint32_t tmem_o = tmem_base + TBN;    // columns [128, 128 + D)

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, C::Q_BYTES);
#pragma unroll
      for (int pn = 0; pn < C::NP; ++pn) tma_load_2d(sQ + pn * PANEL, &tm_q, q_full, h * D + pn * 64, seq_start + q0);
      for (int j = 0; j < n_blocks; ++j) {
        const int st = j & 1;
        mbar_wait(&kv_empty[st], ((j >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&kv_full[st], 2 * C::KV_BYTES);
#pragma unroll
        for (int pn = 0; pn < C::NP; ++pn) {
          tma_load_2d(sK + st * C::KV_BYTES + pn * PANEL, &tm_k, &kv_full[st], kvh * D + pn * 64, seq_start + j * TBN);
          tma_load_2d(sV + st * C::KV_BYTES + pn * PANEL, &tm_v, &kv_full[st], kvh * D + pn * 64, seq_start + j * TBN);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = idesc_bf16(TBM, TBN, 0);
      constexpr uint32_t idesc_o = idesc_bf16(TBM, D, 1);
      mbar_wait(q_full, 0);
      for (int j = 0; j < n_blocks; ++j) {
        const int st = j & 1;
        mbar_wait(&kv_full[st], (j >> 1) & 1);
        tc_fence_after();
        // ---- S = Q K^T
        const uint32_t aq = smem_u32(sQ), bk = smem_u32(sK + st * C::KV_BYTES);
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk) {
          const uint32_t off = (kk >> 2) * PANEL + (kk & 3) * 32;
          umma_bf16_ss(tmem_s, umma_smem_desc_sw128(aq + off), umma_smem_desc_sw128(bk + off), idesc_s, kk > 0 ? 1u : 0u);
        }
        umma_commit(s_full);
        // ---- O_blk = P V   (needs P(j) in smem and O_blk(j-1) drained)
        mbar_wait(p_full, j & 1);
        mbar_wait(o_empty, (j & 1) ^ 1);
        tc_fence_after();
        const uint32_t ap = smem_u32(sP), bv = smem_u32(sV + st * C::KV_BYTES);
#pragma unroll
        for (int kk = 0; kk < TBN / 16; ++kk) {
          const uint64_t da = umma_smem_desc_sw128(ap + (kk >> 2) * PANEL + (kk & 3) * 32);
          const uint64_t db = umma_desc_mn_sw128(bv + kk * 2048, PANEL, 1024);
          umma_bf16_ss(tmem_o, da, db, idesc_o, kk > 0 ? 1u : 0u);
        }
        umma_commit(o_full);
        umma_commit(&kv_empty[st]);
      }
    }
  } else {
    // ===================== softmax / accumulate warps: thread = query row =====================
    const int quad = warp & 3;
    const int row = quad * 32 + lane;          // row inside the tile == TMEM lane
    const int qrow = q0 + row;                 // sequence-relative query index
    const uint32_t lane_addr = static_cast<uint32_t>(quad * 32) << 16;
    float o_acc[D];
#pragma unroll
    for (int i = 0; i < D; ++i) o_acc[i] = 0.f;
    float m_run = -INFINITY, l_run = 0.f;
    const uint32_t sp_row = smem_u32(sP) + row * 128;

    for (int j = 0; j < n_blocks; ++j) {
      const int k0 = j * TBN;
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      // pass 1: row maximum over the valid keys of this block
      float mx = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < TBN / 32; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(tmem_s + lane_addr + c * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int key = k0 + c * 32 + i;
          const bool ok = key < len && (!CAUSAL || key <= qrow);
          if (ok) mx = fmaxf(mx, __uint_as_float(r[i]));
        }
      }
      const float m_new = fmaxf(m_run, mx);
      const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
      const float corr = exp2f((m_run - m_use) * p.scale_log2);   // m_run = -inf -> 0
      const float msl = m_use * p.scale_log2;
      m_run = m_new;
      // pass 2: p = exp2(s*scale - m*scale), row sum, bf16 P into the K-major swizzled smem tile
      float ls = 0.f;
#pragma unroll 1
      for (int c = 0; c < TBN / 32; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(tmem_s + lane_addr + c * 32, r);
        tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const int key = k0 + c * 32 + i;
          const bool ok0 = key < len && (!CAUSAL || key <= qrow);
          const bool ok1 = key + 1 < len && (!CAUSAL || key + 1 <= qrow);
          const float p0 = ok0 ? exp2f(__uint_as_float(r[i]) * p.scale_log2 - msl) : 0.f;
          const float p1 = ok1 ? exp2f(__uint_as_float(r[i + 1]) * p.scale_log2 - msl) : 0.f;
          ls += p0 + p1;
          pk[i >> 1] = pack_bf16x2(p0, p1);
        }
        // 32 keys = 4 chunks of 16 bytes; chunk index inside the 64-key panel = (c & 1) * 4 + q
        const uint32_t panel = sp_row + (c >> 1) * PANEL;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int ch = ((c & 1) * 4 + q) ^ (row & 7);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(panel + ch * 16), "r"(pk[q * 4]),
                       "r"(pk[q * 4 + 1]), "r"(pk[q * 4 + 2]), "r"(pk[q * 4 + 3])
                       : "memory");
        }
      }
      l_run = l_run * corr + ls;
      fence_proxy_async();   // generic-proxy smem writes -> visible to the tensor core (async proxy)
      tc_fence_before();
      mbar_arrive(p_full);
      // O = O * corr + O_blk
      mbar_wait(o_full, j & 1);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < D / 32; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(tmem_o + lane_addr + c * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) o_acc[c * 32 + i] = o_acc[c * 32 + i] * corr + __uint_as_float(r[i]);
      }
      tc_fence_before();
      mbar_arrive(o_empty);
    }
    if (qrow < len) {
      const float inv = l_run > 0.f ? 1.0f / l_run : 0.f;
      __nv_bfloat16* dst = p.o + (size_t)(seq_start + qrow) * p.ldo + (size_t)h * D;
#pragma unroll
      for (int g = 0; g < D / 8; ++g) {
        uint4 q;
        q.x = pack_bf16x2(o_acc[g * 8] * inv, o_acc[g * 8 + 1] * inv);
        q.y = pack_bf16x2(o_acc[g * 8 + 2] * inv, o_acc[g * 8 + 3] * inv);
        q.z = pack_bf16x2(o_acc[g * 8 + 4] * inv, o_acc[g * 8 + 5] * inv);
        q.w = pack_bf16x2(o_acc[g * 8 + 6] * inv, o_acc[g * 8 + 7] * inv);
        *reinterpret_cast<uint4*>(dst + g * 8) = q;
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

using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
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

int make_map(CUtensorMap* tm, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return OPUS_ERR_DRIVER;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {64, 128};
  cuuint32_t estr[2] = {1, 1};
  return fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS
             ? OPUS_OK
             : OPUS_ERR_TMAP;
}

template <int D, bool CAUSAL>
int launch_tc(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const AttnTcParams& p, int n_seqs,
              int max_len, int n_q_heads, cudaStream_t st) {
  using C = TcCfg<D>;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(attn_fwd_tcgen05_kernel<D, CAUSAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM) !=
        cudaSuccess)
      return OPUS_ERR_CUDA;
    configured = true;
  }
  dim3 grid((max_len + TBM - 1) / TBM, n_q_heads, n_seqs);
  const cudaError_t le =
      launch_pdl(false, attn_fwd_tcgen05_kernel<D, CAUSAL>, grid, dim3(TC_THREADS), C::SMEM, st, tq, tk, tv, p);
  note_launch();
  return (le == cudaSuccess && cudaGetLastError() == cudaSuccess) ? OPUS_OK : OPUS_ERR_CUDA;
}

}  // namespace

// Same contract as attn_varlen (attention.cu). n_tok = total packed rows (for the tensor maps).
int attn_varlen_tc(const __nv_bfloat16* q, int ldq, const __nv_bfloat16* k, int ldk, const __nv_bfloat16* v, int ldv,
                   __nv_bfloat16* o, int ldo, const int* cu_seqlens, int n_seqs, int n_tok, int max_len, int n_q_heads,
                   int n_kv_heads, int head_dim, int causal, float scale, cudaStream_t st) {
  if (n_seqs == 0 || max_len == 0) return OPUS_OK;
  if ((ldq | ldk | ldv | ldo) % 8) return OPUS_ERR_ARG;
  if ((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v) |
       reinterpret_cast<uintptr_t>(o)) & 15)
    return OPUS_ERR_ARG;
  CUtensorMap tq, tk, tv;
  int rc = make_map(&tq, q, n_tok, (uint64_t)n_q_heads * head_dim, ldq);
  if (rc) return rc;
  rc = make_map(&tk, k, n_tok, (uint64_t)n_kv_heads * head_dim, ldk);
  if (rc) return rc;
  rc = make_map(&tv, v, n_tok, (uint64_t)n_kv_heads * head_dim, ldv);
  if (rc) return rc;
  AttnTcParams p;
  p.cu_seqlens = cu_seqlens;
  p.o = o; p.ldo = ldo;
  p.group = n_q_heads / n_kv_heads;
  p.scale_log2 = scale * 1.4426950408889634f;
  if (head_dim == 128 && causal) return launch_tc<128, true>(tq, tk, tv, p, n_seqs, max_len, n_q_heads, st);
  if (head_dim == 128 && !causal) return launch_tc<128, false>(tq, tk, tv, p, n_seqs, max_len, n_q_heads, st);
  if (head_dim == 64 && causal) return launch_tc<64, true>(tq, tk, tv, p, n_seqs, max_len, n_q_heads, st);
  if (head_dim == 64 && !causal) return launch_tc<64, false>(tq, tk, tv, p, n_seqs, max_len, n_q_heads, st);
  return OPUS_ERR_ARG;
}

}  // namespace opus
