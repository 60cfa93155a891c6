// Flash attention forward on tcgen05 tensor cores (variable-length packed sequences; bidirectional or causal GQA).
//
// One CTA = TWO 128-row query tiles (A, B) of one (sequence, head), processed ping-pong so that the tensor pipe works on
// one tile while the softmax warps of the other tile are busy. Warp roles (320 threads):
//   warp 0      TMA producer: both Q tiles once, then K blocks and V blocks (128 keys) into two independent rings
//               (SWIZZLE_128B panels of 64 columns);
//   warp 1      MMA issuer (one thread) + TMEM owner:
//                 S_X = Q_X K^T        UMMA 128x128x16, both operands K-major from shared memory
//                 O_X (+)= P_X V       UMMA 128xDx16, A = P read from TENSOR MEMORY, B = V MN-major from shared memory
//   warps 2-5   softmax of tile A, warps 6-9 softmax of tile B: thread = query row (TMEM lane), so row max / row sum need
//               no shuffles. One pass: S (128 fp32) -> registers, max, exp2, row sum, bf16 P written back with
//               tcgen05.st over the first 64 columns of S (P aliases S). O stays in TMEM across key blocks; the softmax
//               warps rescale it in place only when a row's running maximum grew by more than 2^8 (lazy rescaling:
//               O and the row sum always share one reference maximum, so the result is exact).
// TMEM (512 columns): S_A [0,128) S_B [128,256) O_A [256,256+D) O_B [256+D,256+2D).
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

constexpr int TBM = 128;   // query rows per tile (two tiles per CTA)
constexpr int TBN = 128;   // keys per block
constexpr int PANEL = 128 * 128;  // bytes of one [128 rows x 64 bf16] swizzled panel
constexpr int TC_THREADS = 320;
constexpr float kRescaleThreshold = 8.0f;  // log2 units

struct AttnTcParams {
  const int* cu_seqlens;
  __nv_bfloat16* o;
  int ldo;
  int group;          // q heads per kv head
  float scale_log2;   // softmax scale * log2(e)
  int n_seqs, n_q_heads, n_mblk, n_items;  // work list: n_mblk tile pairs x n_seqs x n_q_heads
  int skip_tail;      // > 0: a sequence's last work item is left out when it holds <= skip_tail query rows (the caller
                      // runs those rows on the 64-row mma.sync kernel: a 2-row tail would cost a whole 128-row pipeline)
};

template <int D>
struct TcCfg {
  static constexpr int NP = D / 64;                       // panels per operand tile
  static constexpr int TILE_BYTES = NP * PANEL;           // one Q tile / one K block / one V block
  static constexpr int KS = (D == 128) ? 2 : 3;           // K ring stages
  static constexpr int VS = (D == 128) ? 2 : 3;           // V ring stages
  static constexpr int SMEM = (2 + KS + VS) * TILE_BYTES + 1024 + 512;
  static constexpr int TMEM_COLS = 512;
  static constexpr int COL_S = 0, COL_O = 256;
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

// D[tmem] (+)= A[tmem] * B[smem]: the A operand (bf16, K-major, two K elements per 32-bit column) lives in tensor memory.
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

// registers -> TMEM: this warp's 32 lanes x 32 consecutive 32-bit columns
__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// keys of block j that exist, rounded up to the UMMA N / K granularity: the tail block of a sequence is computed narrow
// (S = Q K^T with N = nk, O += P V over nk keys) instead of as a full 128-key block
__device__ __forceinline__ int block_keys(int len, int j) { return min(TBN, ((len - j * TBN + 15) >> 4) << 4); }

// 2^x on the FMA / ALU pipes (Cody-Waite split + degree-3 polynomial, max relative error 1e-4: a twentieth of a bf16
// half-ulp, and P is rounded to bf16 before it is used). x <= ~2^8 by construction (lazy rescaling), any negative value.
__device__ __forceinline__ float exp2_fma(float x) {
  x = fmaxf(x, -126.0f);
  const float t = x + 12582912.0f;             // 1.5 * 2^23: round(x) lands in the low mantissa bits
  const float f = x - (t - 12582912.0f);       // [-0.5, 0.5]
  float p = fmaf(0.05583828315138817f, f, 0.2426394820213318f);
  p = fmaf(p, f, 0.6931367516517639f);
  p = fmaf(p, f, 0.9999245405197144f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));   // * 2^round(x)
}
// Which of 16 consecutive S columns take the polynomial instead of MUFU.EX2: the exponentials of a key block are the
// critical resource of this kernel (16 MUFU lanes per SM against 128 FMA lanes; at head_dim 64 the MUFU time of a block
// is twice its tensor time), so 7 of 16 go to the FMA pipe, which balances the two (MUFU 9/16 per element, FMA
// (2 + 6 * 7/16) / 128).
constexpr uint32_t kExpPolyMask = 0x552Au;

// Softmax of one row over NC (= 128, or 32 for narrow tail blocks) S columns: S -> registers, mask, row maximum, lazy
// reference-maximum update, exp2, row sum, bf16 P written back over S. Returns whether O must be rescaled by `corr`.
template <int NC, bool CAUSAL>
__device__ __forceinline__ bool softmax_block(uint32_t t_s, int k0, int len, int qrow, int qt0, float scale_log2,
                                              bool first, float& m_run, float& l_run, float& corr) {
  uint32_t s[NC];
#pragma unroll
  for (int c = 0; c < NC / 32; ++c) tmem_ld_32x32(t_s + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&s[c * 32]));
  tmem_ld_wait();
  const bool need_mask = (k0 + NC > len) || (CAUSAL && k0 + NC - 1 > qt0);
  if (need_mask) {
#pragma unroll
    for (int i = 0; i < NC; ++i) {
      const int key = k0 + i;
      const bool ok = key < len && (!CAUSAL || key <= qrow);
      if (!ok) s[i] = 0xff800000u;  // -inf
    }
  }
  float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
  for (int i = 0; i < NC; i += 4) {
    mx0 = fmaxf(mx0, __uint_as_float(s[i]));
    mx1 = fmaxf(mx1, __uint_as_float(s[i + 1]));
    mx2 = fmaxf(mx2, __uint_as_float(s[i + 2]));
    mx3 = fmaxf(mx3, __uint_as_float(s[i + 3]));
  }
  const float m_blk = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * scale_log2;
  // lazy rescaling: keep the old reference maximum unless the new one is more than 2^8 larger
  corr = 1.0f;
  bool rescale = false;
  if (m_blk > m_run + kRescaleThreshold) {   // also true for the first block (m_run = -inf)
    corr = exp2f(m_run - m_blk);             // 0 when m_run = -inf
    rescale = !first;
    m_run = m_blk;
  }
  const float m_use = (m_run == -INFINITY) ? 0.f : m_run;
  float ls0 = 0.f, ls1 = 0.f;
  uint32_t pk[NC / 2];
#pragma unroll
  for (int i = 0; i < NC; i += 2) {
    const float x0 = fmaf(__uint_as_float(s[i]), scale_log2, -m_use);
    const float x1 = fmaf(__uint_as_float(s[i + 1]), scale_log2, -m_use);
    const float p0 = ((kExpPolyMask >> (i & 15)) & 1u) ? exp2_fma(x0) : ex2_approx(x0);
    const float p1 = ((kExpPolyMask >> ((i + 1) & 15)) & 1u) ? exp2_fma(x1) : ex2_approx(x1);
    ls0 += p0;
    ls1 += p1;
    pk[i >> 1] = pack_bf16x2(p0, p1);
  }
  l_run = l_run * corr + (ls0 + ls1);
  if constexpr (NC == 128) {
    tmem_st_32x32(t_s, *reinterpret_cast<uint32_t(*)[32]>(&pk[0]));
    tmem_st_32x32(t_s + 32, *reinterpret_cast<uint32_t(*)[32]>(&pk[32]));
  } else {
    tmem_st_32x16(t_s, *reinterpret_cast<uint32_t(*)[16]>(&pk[0]));
  }
  return rescale;
}

// One unit of work: two adjacent 128-row query tiles of one (sequence, head).
struct Item {
  int seq_start, len, h, kvh, q0;
  int n_a, n_b, n_max;  // key blocks of tile A, tile B, and of the pair
  bool valid;
};

template <bool CAUSAL>
__device__ __forceinline__ Item get_item(const AttnTcParams& p, int w) {
  Item it;
  const int per_mblk = p.n_seqs * p.n_q_heads;
  const int mb = w / per_mblk;
  const int rem = w - mb * per_mblk;
  const int b = rem / p.n_q_heads;
  it.h = rem - b * p.n_q_heads;
  it.kvh = it.h / p.group;
  const int mblk = CAUSAL ? (p.n_mblk - 1 - mb) : mb;   // causal: heavy tiles first
  it.q0 = mblk * 2 * TBM;
  it.seq_start = p.cu_seqlens[b];
  it.len = p.cu_seqlens[b + 1] - it.seq_start;
  it.valid = it.q0 < it.len && !(p.skip_tail > 0 && it.len - it.q0 <= p.skip_tail);
  it.n_a = ((CAUSAL ? min(it.len, it.q0 + TBM) : it.len) + TBN - 1) / TBN;
  it.n_b = (it.q0 + TBM < it.len) ? ((CAUSAL ? min(it.len, it.q0 + 2 * TBM) : it.len) + TBN - 1) / TBN : 0;
  it.n_max = it.n_b > 0 ? it.n_b : it.n_a;
  return it;
}

// NARROW: tail key blocks are computed with N = K-extent = the keys that exist (worth it for short sequences such as the
// 258-token encoder batch, where the third block holds 2 keys; it costs ~8 % on long sequences, so the host picks).
template <int D, bool CAUSAL, bool NARROW>
__global__ void __launch_bounds__(TC_THREADS, 1)
attn_fwd_tcgen05_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                        const __grid_constant__ CUtensorMap tm_v, const AttnTcParams p) {
  using C = TcCfg<D>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                                  // [2 tiles][TILE_BYTES]
  uint8_t* sK = sQ + 2 * C::TILE_BYTES;                // [KS][TILE_BYTES]
  uint8_t* sV = sK + C::KS * C::TILE_BYTES;            // [VS][TILE_BYTES]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + C::VS * C::TILE_BYTES);
  uint64_t* q_full = bars;                  // 1   TMA -> MMA
  uint64_t* q_empty = bars + 1;             // 1   MMA -> TMA: every S MMA of the item has read Q
  uint64_t* k_full = bars + 2;              // [KS]
  uint64_t* k_empty = k_full + C::KS;       // [KS]
  uint64_t* v_full = k_empty + C::KS;       // [VS]
  uint64_t* v_empty = v_full + C::VS;       // [VS]
  uint64_t* s_full = v_empty + C::VS;       // [2]  MMA -> softmax: S_X(j) complete
  uint64_t* p_full = s_full + 2;            // [2]  softmax -> MMA: P_X(j) stored (and O_X rescaled)
  uint64_t* o_done = p_full + 2;            // [2]  MMA -> softmax: O_X includes block j
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_done + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_q); tma_prefetch_desc(&tm_k); tma_prefetch_desc(&tm_v);
    mbar_init(q_full, 1);
    mbar_init(q_empty, 1);
    for (int s = 0; s < C::KS; ++s) { mbar_init(&k_full[s], 1); mbar_init(&k_empty[s], 1); }
    for (int s = 0; s < C::VS; ++s) { mbar_init(&v_full[s], 1); mbar_init(&v_empty[s], 1); }
    for (int x = 0; x < 2; ++x) { mbar_init(&s_full[x], 1); mbar_init(&p_full[x], 128); mbar_init(&o_done[x], 1); }
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

  // Persistent CTA: every role walks the same static list of work items; barrier phases are tracked with running use
  // counters, so the O read-out of one item overlaps the loads and the first S MMA of the next.
  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t kc = 0, vc = 0, qc = 0;   // K blocks, V blocks, items issued so far
      for (int w = blockIdx.x; w < p.n_items; w += gridDim.x) {
        const Item it = get_item<CAUSAL>(p, w);
        if (!it.valid) continue;
        mbar_wait(q_empty, (qc & 1) ^ 1);
        mbar_arrive_expect_tx(q_full, 2 * C::TILE_BYTES);
#pragma unroll
        for (int x = 0; x < 2; ++x)
#pragma unroll
          for (int pn = 0; pn < C::NP; ++pn)
            tma_load_2d(sQ + x * C::TILE_BYTES + pn * PANEL, &tm_q, q_full, it.h * D + pn * 64,
                        it.seq_start + it.q0 + x * TBM);
        ++qc;
        for (int j = 0; j < it.n_max; ++j) {
          const uint32_t ks = kc % C::KS, vs = vc % C::VS;
          mbar_wait(&k_empty[ks], ((kc / C::KS) & 1) ^ 1);
          mbar_arrive_expect_tx(&k_full[ks], C::TILE_BYTES);
#pragma unroll
          for (int pn = 0; pn < C::NP; ++pn)
            tma_load_2d(sK + ks * C::TILE_BYTES + pn * PANEL, &tm_k, &k_full[ks], it.kvh * D + pn * 64,
                        it.seq_start + j * TBN);
          ++kc;
          mbar_wait(&v_empty[vs], ((vc / C::VS) & 1) ^ 1);
          mbar_arrive_expect_tx(&v_full[vs], C::TILE_BYTES);
#pragma unroll
          for (int pn = 0; pn < C::NP; ++pn)
            tma_load_2d(sV + vs * C::TILE_BYTES + pn * PANEL, &tm_v, &v_full[vs], it.kvh * D + pn * 64,
                        it.seq_start + j * TBN);
          ++vc;
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc_o = idesc_bf16(TBM, D, 1);
      uint32_t kc = 0, vc = 0, qc = 0, pc[2] = {0, 0};
      auto issue_s = [&](int x, uint32_t ks, int nk) {   // nk = keys of the block (multiple of 16) = UMMA N
        const uint32_t aq = smem_u32(sQ + x * C::TILE_BYTES), bk = smem_u32(sK + ks * C::TILE_BYTES);
        const uint32_t idesc_s = idesc_bf16(TBM, (uint32_t)nk, 0);
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk) {
          const uint32_t off = (kk >> 2) * PANEL + (kk & 3) * 32;
          umma_bf16_ss(tmem_base + C::COL_S + x * TBN, umma_smem_desc_sw128(aq + off), umma_smem_desc_sw128(bk + off),
                       idesc_s, kk > 0 ? 1u : 0u);
        }
        umma_commit(&s_full[x]);
      };
      for (int w = blockIdx.x; w < p.n_items; w += gridDim.x) {
        const Item it = get_item<CAUSAL>(p, w);
        if (!it.valid) continue;
        const int n_x[2] = {it.n_a, it.n_b};
        mbar_wait(q_full, qc & 1);
        ++qc;
        mbar_wait(&k_full[kc % C::KS], (kc / C::KS) & 1);
        tc_fence_after();
        // S_X(0) overwrites P_X of the previous item: ordered after that item's last PV MMA (same issuing thread)
        issue_s(0, kc % C::KS, NARROW ? block_keys(it.len, 0) : TBN);
        if (it.n_b > 0) issue_s(1, kc % C::KS, NARROW ? block_keys(it.len, 0) : TBN);
        umma_commit(&k_empty[kc % C::KS]);
        ++kc;
        if (it.n_max == 1) umma_commit(q_empty);
        for (int j = 0; j < it.n_max; ++j) {
          const uint32_t vs = vc % C::VS, ks1 = kc % C::KS;
          mbar_wait(&v_full[vs], (vc / C::VS) & 1);
          bool k_ready = false;
#pragma unroll
          for (int x = 0; x < 2; ++x) {
            if (j >= n_x[x]) continue;
            mbar_wait(&p_full[x], pc[x] & 1);
            ++pc[x];
            tc_fence_after();
            const uint32_t bv = smem_u32(sV + vs * C::TILE_BYTES);
            const int ksteps = NARROW ? (block_keys(it.len, j) >> 4) : (TBN / 16);
#pragma unroll
            for (int kk = 0; kk < TBN / 16; ++kk) {
              // P: 16 keys = 8 packed columns per step; V: 16 key rows = two 8-row swizzle groups (2048 B) per step
              if (kk < ksteps)
                umma_bf16_ts(tmem_base + C::COL_O + x * D, tmem_base + C::COL_S + x * TBN + kk * 8,
                             umma_desc_mn_sw128(bv + kk * 2048, PANEL, 1024), idesc_o, (j > 0 || kk > 0) ? 1u : 0u);
            }
            umma_commit(&o_done[x]);
            if (j + 1 < n_x[x]) {
              if (!k_ready) {
                mbar_wait(&k_full[ks1], (kc / C::KS) & 1);
                tc_fence_after();
                k_ready = true;
              }
              issue_s(x, ks1, NARROW ? block_keys(it.len, j + 1) : TBN);   // overwrites S_X / P_X(j): ordered after the PV MMAs above
            }
          }
          umma_commit(&v_empty[vs]);
          ++vc;
          if (k_ready) {
            umma_commit(&k_empty[ks1]);
            ++kc;
            if (j + 2 == it.n_max) umma_commit(q_empty);   // the last S MMAs of this item have been issued
          }
        }
      }
    }
  } else {
    // ===================== softmax warps: thread = query row =====================
    const int x = (warp - 2) >> 2;             // tile A (warps 2-5) or B (warps 6-9)
    const int quad = warp & 3;                 // TMEM lane quadrant this warp may access (warp id % 4)
    const int row = quad * 32 + lane;          // row inside the tile == TMEM lane
    const uint32_t lane_addr = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t t_s = tmem_base + lane_addr + C::COL_S + x * TBN;
    const uint32_t t_o = tmem_base + lane_addr + C::COL_O + x * D;
    uint32_t g = 0;                            // key blocks processed by this warp group so far (barrier phases)

    for (int w = blockIdx.x; w < p.n_items; w += gridDim.x) {
      const Item it = get_item<CAUSAL>(p, w);
      if (!it.valid) continue;
      const int n_blocks = x == 0 ? it.n_a : it.n_b;
      if (n_blocks == 0) continue;
      const int len = it.len;
      const int qt0 = it.q0 + x * TBM;           // first query row of the tile (sequence-relative)
      const int qrow = qt0 + row;
      float m_run = -INFINITY, l_run = 0.f;      // m_run in scaled log2 units

      for (int j = 0; j < n_blocks; ++j, ++g) {
        const int k0 = j * TBN;
        mbar_wait(&s_full[x], g & 1);
        tc_fence_after();
        float corr;
        const bool rescale = (NARROW && block_keys(len, j) <= 32)
                                 ? softmax_block<32, CAUSAL>(t_s, k0, len, qrow, qt0, p.scale_log2, j == 0, m_run, l_run, corr)
                                 : softmax_block<128, CAUSAL>(t_s, k0, len, qrow, qt0, p.scale_log2, j == 0, m_run, l_run, corr);
        if (j > 0) {
          // O_X must include block j-1 before it may be rescaled / before PV(j) accumulates on top of it
          mbar_wait(&o_done[x], (g - 1) & 1);
          tc_fence_after();
          if (__any_sync(0xffffffffu, rescale)) {
#pragma unroll 1
            for (int c = 0; c < D / 32; ++c) {
              uint32_t r[32];
              tmem_ld_32x32(t_o + c * 32, r);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * corr);
              tmem_st_32x32(t_o + c * 32, r);
            }
          }
        }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&p_full[x]);
      }
      // O_X is complete once the PV MMA of the last block retires; the next item's PV_X(0) cannot start before this
      // warp group has produced that item's P_X(0), i.e. after this read-out.
      mbar_wait(&o_done[x], (g - 1) & 1);
      tc_fence_after();
      const float inv = l_run > 0.f ? 1.0f / l_run : 0.f;
      __nv_bfloat16* dst = p.o + (size_t)(it.seq_start + qrow) * p.ldo + (size_t)it.h * D;
#pragma unroll 1
      for (int c = 0; c < D / 32; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(t_o + c * 32, r);
        tmem_ld_wait();
        if (qrow < len) {
#pragma unroll
          for (int gq = 0; gq < 4; ++gq) {
            uint4 q;
            q.x = pack_bf16x2(__uint_as_float(r[gq * 8]) * inv, __uint_as_float(r[gq * 8 + 1]) * inv);
            q.y = pack_bf16x2(__uint_as_float(r[gq * 8 + 2]) * inv, __uint_as_float(r[gq * 8 + 3]) * inv);
            q.z = pack_bf16x2(__uint_as_float(r[gq * 8 + 4]) * inv, __uint_as_float(r[gq * 8 + 5]) * inv);
            q.w = pack_bf16x2(__uint_as_float(r[gq * 8 + 6]) * inv, __uint_as_float(r[gq * 8 + 7]) * inv);
            *reinterpret_cast<uint4*>(dst + c * 32 + gq * 8) = q;
          }
        }
      }
      tc_fence_before();
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

template <int D, bool CAUSAL, bool NARROW>
int launch_tc(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const AttnTcParams& p, int n_seqs,
              int max_len, int n_q_heads, cudaStream_t st) {
  using C = TcCfg<D>;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(attn_fwd_tcgen05_kernel<D, CAUSAL, NARROW>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM) !=
        cudaSuccess)
      return OPUS_ERR_CUDA;
    configured = true;
  }
  int sms = 0, dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  dim3 grid(p.n_items < sms ? p.n_items : sms);
  const cudaError_t le =
      launch_pdl(false, attn_fwd_tcgen05_kernel<D, CAUSAL, NARROW>, grid, dim3(TC_THREADS), C::SMEM, st, tq, tk, tv, p);
  note_launch();
  return (le == cudaSuccess && cudaGetLastError() == cudaSuccess) ? OPUS_OK : OPUS_ERR_CUDA;
}

}  // namespace

// Same contract as attn_varlen (attention.cu). n_tok = total packed rows (for the tensor maps).
int attn_varlen_tc(const __nv_bfloat16* q, int ldq, const __nv_bfloat16* k, int ldk, const __nv_bfloat16* v, int ldv,
                   __nv_bfloat16* o, int ldo, const int* cu_seqlens, int n_seqs, int n_tok, int max_len, int n_q_heads,
                   int n_kv_heads, int head_dim, int causal, float scale, cudaStream_t st, int skip_tail) {
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
  p.n_seqs = n_seqs; p.n_q_heads = n_q_heads;
  p.n_mblk = (max_len + 2 * TBM - 1) / (2 * TBM);
  p.n_items = p.n_mblk * n_seqs * n_q_heads;
  p.skip_tail = skip_tail;
  const bool narrow = max_len <= 384;
#define OPUS_TC_CASE(HD, C)                                                                                     \
  if (head_dim == HD && (causal != 0) == C)                                                                     \
    return narrow ? launch_tc<HD, C, true>(tq, tk, tv, p, n_seqs, max_len, n_q_heads, st)                      \
                  : launch_tc<HD, C, false>(tq, tk, tv, p, n_seqs, max_len, n_q_heads, st);
  OPUS_TC_CASE(128, true)
  OPUS_TC_CASE(128, false)
  OPUS_TC_CASE(64, true)
  OPUS_TC_CASE(64, false)
#undef OPUS_TC_CASE
  return OPUS_ERR_ARG;
}

}  // namespace opus
