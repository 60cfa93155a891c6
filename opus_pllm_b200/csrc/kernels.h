// Internal C++ declarations of the bandwidth and attention kernels (launchers). The C-ABI wrappers live in capi.cu.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <climits>
#include <cstddef>
#include <cstdint>

#include "../../include/opus_b200.h"

namespace opus {

// ---- bandwidth.cu ----
int esm_embed(const int* tok, const float* scale, const float* table, float* x, int n_tok, int dim, cudaStream_t st);
// y = LayerNorm(x (+ delta)); when delta != nullptr x is updated in place (x += delta) — delta may alias y.
int layernorm_f32_bf16(float* x, const __nv_bfloat16* delta, const float* gamma, const float* beta, __nv_bfloat16* y,
                       int rows, int cols, float eps, cudaStream_t st);
// y = w * rmsnorm(h), h = x (or bf16(sum of n_partial fp32 partials)) (+ residual); h optionally written to h_out;
// y == nullptr skips the normalisation (pure reduce + residual).
int rmsnorm_bf16(const __nv_bfloat16* x, const float* partial, int n_partial, const __nv_bfloat16* residual,
                 __nv_bfloat16* h_out, const __nv_bfloat16* w, __nv_bfloat16* y, int rows, int cols, float eps,
                 cudaStream_t st);
int splitk_reduce_bf16(const float* partial, int n_partial, const float* bias, __nv_bfloat16* out, int rows, int cols,
                       int ldo, int gelu, cudaStream_t st);
int rope_esm(__nv_bfloat16* qkv, const int* pos, const float* cos_t, const float* sin_t, int n_tok, int n_heads,
             int head_dim, int ld, float q_scale, cudaStream_t st);
int rope_llama_kvappend(__nv_bfloat16* qkv, const float* partial, int n_partial, const int* pos, const int* slot,
                        const __nv_bfloat16* cos_t, const __nv_bfloat16* sin_t, __nv_bfloat16* kcache,
                        __nv_bfloat16* vcache, int n_tok, int n_q_heads, int n_kv_heads, int head_dim, int ld,
                        int block_size, cudaStream_t st, const float* bias = nullptr /* fp32 [ld], split-K path only */);
int final_ln_meanpool(const float* x, const __nv_bfloat16* delta, const int* cu_seqlens, const float* gamma,
                      const float* beta, float* pooled,
                      __nv_bfloat16* pooled_l2, float* hidden_out, int n_seqs, int dim, float eps, cudaStream_t st);
int l2norm_f32_bf16(const float* x, __nv_bfloat16* y, int rows, int dim, cudaStream_t st);
int splice_gather(const int* src, const __nv_bfloat16* embed, const __nv_bfloat16* soft, __nv_bfloat16* out, int n_rows,
                  int dim, cudaStream_t st);
int argmax_eos(const __nv_bfloat16* logits, int ld, int vocab, int n_rows, int* finished, const int* eos_ids, int n_eos,
               int pad_id, int* next_tok, int* out_ids, int out_ld, int step, int* n_unfinished, cudaStream_t st,
               const int* step_ptr = nullptr);
// temperature / top-p sampling with the EOS bookkeeping of argmax_eos (HF do_sample=True); kept_count (nullable) receives
// the number of boundary-key ties kept per row (diagnostics)
int sample_top_p(const __nv_bfloat16* logits, int ld, int vocab, int n_rows, float temperature, float top_p,
                 unsigned long long seed, int* finished, const int* eos_ids, int n_eos, int pad_id, int* next_tok,
                 int* out_ids, int out_ld, int step, int* n_unfinished, cudaStream_t st,
                 const int* step_ptr = nullptr, int* kept_count = nullptr,
                 const unsigned long long* seed_ptr = nullptr /* device; overrides `seed` when non-null */);
// loss[r] = logsumexp(logits[r,:]) - logits[r, target[r]] in fp32 (0 where target[r] < 0)
int cross_entropy_rows(const __nv_bfloat16* logits, int ld, int vocab, const int* target, float* loss, int n_rows,
                       cudaStream_t st);
int embed_gather(const int* tok, const __nv_bfloat16* table, __nv_bfloat16* x, int n_rows, int dim, cudaStream_t st);
// embed_gather + per-32-feature-slab sums of squares of every gathered row: sumsq[slab][row] (norm-fused decode path)
int embed_gather_sumsq(const int* tok, const __nv_bfloat16* table, __nv_bfloat16* x, float* sumsq, int sumsq_ld,
                       int n_rows, int dim, cudaStream_t st);
int decode_advance(int* ctx_len, int* pos, int* slot, const int* block_table, int max_blocks, int block_size, int n,
                   cudaStream_t st, int* step = nullptr);
int lora_merge(__nv_bfloat16* W, const __nv_bfloat16* A, const __nv_bfloat16* B, int out_f, int in_f, int r, float scale,
               cudaStream_t st);

// ---- bandwidth_opt.cu (OPT / Galactica family) ----
// nn.LayerNorm over bf16 rows with the fusions of rmsnorm_bf16 (+ red_bias added to the reduced partial sums)
int layernorm_bf16(const __nv_bfloat16* x, const float* partial, int n_partial, const float* red_bias,
                   const __nv_bfloat16* residual, __nv_bfloat16* h_out, const float* gamma, const float* beta,
                   __nv_bfloat16* y, int rows, int cols, float eps, cudaStream_t st);
// h[i,:] = bf16(h[i,:] + table[pos[i] + offset, :])
int add_pos_embed(__nv_bfloat16* h, const __nv_bfloat16* table, const int* pos, int offset, int table_rows, int n_rows,
                  int dim, cudaStream_t st);

// rows whose emitted tail (ending at column step / *step_ptr of out_ids) equals a stop sequence become finished
int stop_sequences(const int* out_ids, int out_ld, int n_rows, int step, const int* step_ptr, const int* stop_seqs,
                   const int* stop_lens, int n_stop, int stop_ld, int* finished, int* n_unfinished, cudaStream_t st);

// ---- attention.cu ----
int attn_varlen(const __nv_bfloat16* q, int ldq, const __nv_bfloat16* k, int ldk, const __nv_bfloat16* v, int ldv,
                __nv_bfloat16* o, int ldo, const int* cu_seqlens, int n_seqs, int n_tok, int max_len, int n_q_heads,
                int n_kv_heads, int head_dim, int causal, float scale, cudaStream_t st);
// tcgen05 / TMEM / TMA implementation of the same contract (attention_tc.cu); attn_varlen dispatches to it.
int attn_varlen_tc(const __nv_bfloat16* q, int ldq, const __nv_bfloat16* k, int ldk, const __nv_bfloat16* v, int ldv,
                   __nv_bfloat16* o, int ldo, const int* cu_seqlens, int n_seqs, int n_tok, int max_len, int n_q_heads,
                   int n_kv_heads, int head_dim, int causal, float scale, cudaStream_t st, int skip_tail = 0);
int attn_decode_paged(const __nv_bfloat16* q, int ldq, const __nv_bfloat16* kcache, const __nv_bfloat16* vcache,
                      const int* block_table, int max_blocks, const int* ctx_len, __nv_bfloat16* o, int ldo,
                      int n_seqs, int n_q_heads, int n_kv_heads, int head_dim, int block_size, float scale,
                      cudaStream_t st);

// allocates the context's split-KV scratch (call outside stream capture; the decode loop does before it captures)
int attn_decode_warmup();
// Same, with the decode step's split-K reduce + RoPE + KV append done by the attention CTAs themselves: `qkv` is written
// (bf16 q|k|v row of every sequence) from `partial` ([n_partial][n_seqs][ldq] fp32) before the heads are attended.
int attn_decode_paged_fused(__nv_bfloat16* qkv, int ldq, const float* partial, int n_partial, const int* pos,
                            const int* slot, const __nv_bfloat16* cos_t, const __nv_bfloat16* sin_t,
                            __nv_bfloat16* kcache, __nv_bfloat16* vcache, const int* block_table, int max_blocks,
                            const int* ctx_len, __nv_bfloat16* o, int ldo, int n_seqs, int n_q_heads, int n_kv_heads,
                            int head_dim, int block_size, float scale, cudaStream_t st, const float* bias = nullptr);

}  // namespace opus
