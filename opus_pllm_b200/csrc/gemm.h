// Internal (C++) interface of the tcgen05 GEMM; the C-ABI wrapper is opus_gemm_bf16 in capi.cu.
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
#include <cstdint>

#include "../../include/opus_b200.h"

namespace opus {

enum EpiMode : int {
  EPI_BF16 = OPUS_EPI_BF16,                // out bf16 = acc (+ bias)
  EPI_BF16_GELU = OPUS_EPI_BF16_GELU,      // out bf16 = gelu_erf(acc + bias)
  EPI_RES_F32 = OPUS_EPI_RES_F32,          // out f32  = residual_f32 + acc (+ bias)
  EPI_RES_BF16 = OPUS_EPI_RES_BF16,        // out bf16 = bf16(residual_bf16 + bf16(acc (+ bias)))
  EPI_SWIGLU = OPUS_EPI_SWIGLU,            // interleaved (gate, up) features -> out bf16 = silu(gate) * up, half width
  EPI_PARTIAL_F32 = OPUS_EPI_PARTIAL_F32,  // split-K partial sums: out f32 [split][rows][ldo]
  EPI_F32 = OPUS_EPI_F32,                  // out f32 = acc (+ bias)   (transposed form only)
  EPI_BF16_RELU = OPUS_EPI_BF16_RELU,      // out bf16 = relu(acc + bias)   (OPT fc1)
};

// Device-visible parameter block.
struct GemmParams {
  int M, N, K;
  int num_m_tiles, num_n_tiles, k_blocks;
  int split_k;
  int transposed;
  int epi;
  int group_m;
  int group_n;   // CTA-pair plain form: > 0 = grouped-N raster with bands of this many 256-wide weight panels
  void* out;
  int ldo;
  const float* bias;
  const void* residual;
  int ldr;
  uint64_t hint_a, hint_b;
  // next-weight L2 prefetch (swap-AB chains): work item t of the NEXT launch = (m tile t / pf_split_k, k-split
  // t % pf_split_k); the first pf_depth k-blocks of items [0, pf_items) are requested into L2 at this CTA's tail.
  int pf_items, pf_split_k, pf_k_blocks, pf_depth;
  // stream-K tail: the last sk_tiles output tiles (a partial wave) are cut along K into gridDim.x equal unit ranges.
  // A CTA whose range does not contain a tile's last k-block dumps its fp32 accumulator to sk_ws[cta]; the CTA that
  // does ("owner") waits on sk_cnt[tile], adds the partials in CTA order (deterministic) and runs the epilogue.
  int tma_store;  // plain form, bf16 (+bias, +GELU) epilogue: tiles leave through shared memory + TMA stores
  // ESM rotary fused into that epilogue (head_dim 64 = one 64-column store box): columns [0, rope_cols) are rotated with
  // the row's position, columns [0, rope_q_cols) are scaled by rope_q_scale first (q <- rope(q * hd^-0.5), k <- rope(k))
  const int* rope_pos;
  const float* rope_cos;   // fp32 [max_pos, 32]
  const float* rope_sin;
  int rope_cols, rope_q_cols;
  float rope_q_scale;
  // Llama rotary + paged KV-cache append fused into the same epilogue (head_dim 128 = two store boxes; BN = 256 = two
  // heads per tile, one per epilogue warp pair): heads [0, rl_hq) are q (rotated), the next rl_hkv are k (rotated, also
  // written to kcache at slot[row]), the last rl_hkv are v (copied to vcache). Same rounding points as
  // rope_llama_kvappend_kernel (HF apply_rotary_pos_emb in bf16).
  const int* rl_pos;
  const int* rl_slot;
  const void* rl_cos;   // bf16 [max_pos, 128], cat(freqs, freqs)
  const void* rl_sin;
  void* rl_kcache;
  void* rl_vcache;
  int rl_hq, rl_hkv, rl_bs;
  int dp_items;   // work items walked round-robin (full tiles x split_k) before the stream-K tail
  int sk_tiles;
  int sk_share;   // CTAs (pairs) sharing the tail; 0 = all of them. A tail of a few tiles is cut into 8 pieces per tile only
  float* sk_ws;
  int* sk_cnt;
  // In-kernel split-K reduction (swap-AB, one wave): the CTA holding a tile's LAST k-split waits for the CTAs holding the
  // other splits (they dump their fp32 accumulators to sk_ws and bump sk_cnt[tile]), adds the partial sums in split order
  // and runs the real epilogue, so no separate reduce kernel follows the GEMM.
  int sk_fix;
  int l2_ahead;   // swap-AB: k-blocks of the weight stream requested into L2 ahead of the shared-memory ring (tunable)
  int epi_warm;   // swap-AB: run the epilogue once "dry" during the main loop to warm the instruction cache (tunable)
  // EPI_RES_BF16 (swap-AB) by-product for a following RMSNorm: sumsq_out[slab][batch row] = sum over the 32 features of
  // slab (= feature / 32) of the squared bf16 values this launch stored. One writer per entry: deterministic.
  float* sumsq_out;
  int sumsq_ld;
  // RMSNorm applied to the activation operand on its way to the tensor cores (swap-AB, batch <= 64): the B operand is
  // the raw residual stream h; dedicated warps rewrite every landed [batch x 64] k-slice in shared memory as
  // bf16(gamma[k] * bf16(h * rstd[row])) (the rounding points of rmsnorm_bf16_kernel / HF LlamaRMSNorm) before the MMA
  // warp may read it. rstd[row] = rsqrt(sum over nl_slabs of nl_sumsq[slab][row] / K + eps).
  const float* nl_sumsq;
  int nl_slabs, nl_ld;
  const void* nl_gamma;  // bf16 [K]
  float nl_eps;
};

// D[M,N] = A[M,K] * B[N,K]^T, both operands bf16 row-major with K contiguous.
//  transposed == 0: out[m*ldo + n], bias[n]; A = activations, B = weights.
//  transposed == 1: out[n*ldo + m], bias[m]; A = weights (M = output features), B = activations (N = batch rows).
//  split_k > 1 requires EPI_PARTIAL_F32; partial s of element (r, c) [r = batch/activation row, c = feature] is at
//  out[(s*rows + r)*ldo + c].
struct GemmArgs {
  const void* A;
  int lda;
  const void* B;
  int ldb;
  int M, N, K;
  int transposed;
  int epi;
  void* out;
  int ldo;
  const float* bias;
  const void* residual;
  int ldr;
  int split_k;  // 0/1 = none
  int block_n;  // 0 = auto
  int streamk_tail;  // plain form: allow the stream-K tail for this launch (decode at batch 257..512; default: tunable streamk_plain)
  // Optional hint: the weight matrix [pf_rows, pf_K] (row stride pf_K) that the NEXT swap-AB launch on this stream will
  // stream with split factor pf_split_k. When pf_w != nullptr every CTA, after issuing its last own load, asks the TMA
  // unit to prefetch the first pf_depth k-blocks of "its" work item of that launch into L2, so HBM stays busy while
  // this kernel drains, the small kernels in between run, and the next GEMM ramps up.
  const void* pf_w;
  int pf_rows, pf_K, pf_split_k, pf_depth;
  // Optional fused ESM rotary (see GemmParams); honoured only when gemm_fuses_rope(args) is true
  const int* rope_pos;
  const float* rope_cos;
  const float* rope_sin;
  int rope_cols, rope_q_cols;
  float rope_q_scale;
  // Optional fused Llama rotary + KV append (see GemmParams); honoured only when gemm_fuses_rope(args) is true
  const int* rl_pos;
  const int* rl_slot;
  const void* rl_cos;
  const void* rl_sin;
  void* rl_kcache;
  void* rl_vcache;
  int rl_hq, rl_hkv, rl_bs;
  // Decode-chain fusions (swap-AB form; see GemmParams): reduce the split-K partial sums inside the kernel (any epilogue
  // other than EPI_PARTIAL_F32 then works with split_k > 1; needs tiles * split_k <= SM count); emit per-slab sums of
  // squares of an EPI_RES_BF16 result; apply RMSNorm (norm_sumsq / norm_gamma / norm_eps) to the activation operand.
  int splitk_fixup;
  float* sumsq_out;        // fp32 [ceil(M / 32)][sumsq_ld], sumsq_ld >= N (batch rows)
  int sumsq_ld;
  const float* norm_sumsq; // fp32 [norm_slabs][norm_ld]; non-null selects the normalising launch (batch <= 64)
  int norm_slabs, norm_ld;
  const void* norm_gamma;  // bf16 [K]
  float norm_eps;
};
// true when gemm_bf16 will apply the rope_* / rl_* fields of `a` in its epilogue (plain form, EPI_BF16, TMA-store path)
bool gemm_fuses_rope(const GemmArgs& a);

int gemm_bf16(const GemmArgs& a, cudaStream_t stream);

// ---- fused chain of swap-AB GEMMs and split-K-reduce + residual + RMSNorm steps in ONE persistent kernel ----
// Decode runs o_proj -> norm -> gate/up(SwiGLU) -> down -> norm -> next qkv (or lm_head) as one launch: phases are
// separated by a device-wide barrier (one CTA per SM, all resident), the TMA producer runs one ring pass ahead into the
// next phase's WEIGHTS, and barriers / TMEM / tensor maps are set up once instead of once per GEMM.
constexpr int kMaxChainPhases = 6;
enum ChainPhaseKind : int { CHAIN_GEMM = 0, CHAIN_NORM = 1 };
struct ChainNorm {
  // h = bf16(residual + bf16(sum_s partial[s]));  h_out = h;  y = w * bf16(h * rsqrt(mean(h^2) + eps))   (rmsnorm_bf16)
  const float* partial;
  int n_partial;
  const void* residual;  // bf16 [rows, cols]
  void* h_out;           // bf16 [rows, cols] (may alias residual)
  const void* w;         // bf16 [cols]
  void* y;               // bf16 [rows, cols]
  int rows, cols;
  float eps;
};
struct ChainPhase {
  int kind;
  GemmArgs gemm;   // CHAIN_GEMM: transposed form only (weights = A); split_k as given (0/1 = none)
  ChainNorm norm;  // CHAIN_NORM
};
// All GEMM phases must share the batch size (B operand rows <= 256). Returns OPUS_OK or an error code.
int gemm_chain(const ChainPhase* phases, int n_phases, cudaStream_t stream);
int gemm_chain_trace(int enable, unsigned long long* out, int cap_words);
void gemm_set_chain_l2_depth(int kblocks);
int gemm_pick_bn(int N, int transposed);
int gemm_pick_split_k(int M, int N, int K, int bn);
int gemm_pick_split_k_wide(int M, int N, int K);   // batch 257..512 on the CTA-pair swap-AB kernel
size_t gemm_workspace_bytes(int M, int N, int split_k);
void gemm_set_streamk_fill(int percent);  // 0 disables the stream-K tail
void gemm_set_streamk_plain(int on);
void gemm_set_tma_store(int on);
void gemm_set_2cta_tr(int on);            // CTA-pair form for swap-AB launches at batch 129..256 (default on)
void gemm_set_2cta(int on);               // CTA-pair (cta_group::2) form for large plain GEMMs          // TMA-store epilogue (and the rotary fusion that rides on it)     // stream-K tail in the plain (non swap-AB) form, off by default

}  // namespace opus
