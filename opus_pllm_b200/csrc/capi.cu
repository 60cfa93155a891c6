// extern "C" surface of libopus_b200.so (declared in include/opus_b200.h): argument marshalling only.
#include <atomic>
#include <cstdio>
#include <cstring>
#include <string>

#include "common.h"
#include "context.h"
#include "gemm.h"
#include "kernels.h"
#include "models.h"

namespace opus {

namespace {
thread_local std::string t_err;   // per calling thread: concurrent callers never see each other's message
std::atomic<long long> g_launches{0};
}  // namespace

int fail(int rc, const char* what) {
  const cudaError_t ce = cudaPeekAtLastError();
  t_err = std::string(what ? what : "?") + " -> " + std::to_string(rc);
  if (ce != cudaSuccess) t_err += std::string(" [cuda: ") + cudaGetErrorString(ce) + "]";
  return rc;
}
void note_launch(long long n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
long long launch_count(bool reset) {
  return reset ? g_launches.exchange(0, std::memory_order_relaxed) : g_launches.load(std::memory_order_relaxed);
}

}  // namespace opus

using namespace opus;
using bf16 = __nv_bfloat16;

#define ST(s) static_cast<cudaStream_t>(s)
#define RET(expr, name)                          \
  do {                                           \
    const int _rc = (expr);                      \
    return _rc == OPUS_OK ? OPUS_OK : fail(_rc, name); \
  } while (0)

extern "C" {

int opus_abi_version(void) { return OPUS_B200_ABI_VERSION; }

const char* opus_last_error(void) { return t_err.c_str(); }

int opus_ctx_create(opus_ctx** out) {
  if (out == nullptr) return fail(OPUS_ERR_ARG, "opus_ctx_create: null out");
  *out = reinterpret_cast<opus_ctx*>(ctx_create());
  return OPUS_OK;
}
int opus_ctx_destroy(opus_ctx* c) {
  ctx_destroy(reinterpret_cast<Context*>(c));
  return OPUS_OK;
}
int opus_ctx_set_current(opus_ctx* c) {
  ctx_set_current(reinterpret_cast<Context*>(c));
  return OPUS_OK;
}

int opus_device_check(void) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return fail(OPUS_ERR_CUDA, "opus_device_check: no CUDA device");
  int major = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (major != 10) return fail(OPUS_ERR_CUDA, "opus_device_check: device is not sm_100");
  return OPUS_OK;
}

int opus_gemm_bf16(const void* A, int lda, const void* B, int ldb, int M, int N, int K, int transposed, int epilogue,
                   void* out, int ldo, const float* bias, const void* residual, int ldr, int split_k, int block_n,
                   void* stream) {
  if (!A || !B || !out) return fail(OPUS_ERR_ARG, "opus_gemm_bf16: null pointer");
  GemmArgs a{};
  a.A = A; a.lda = lda; a.B = B; a.ldb = ldb;
  a.M = M; a.N = N; a.K = K;
  a.transposed = transposed; a.epi = epilogue;
  a.out = out; a.ldo = ldo; a.bias = bias; a.residual = residual; a.ldr = ldr;
  a.split_k = split_k; a.block_n = block_n;
  RET(gemm_bf16(a, ST(stream)), "opus_gemm_bf16");
}

int opus_gemm_bf16_fused(const void* A, int lda, const void* B, int ldb, int M, int N, int K, int epilogue, void* out,
                         int ldo, const float* bias, const void* residual, int ldr, int split_k, int splitk_fixup,
                         float* sumsq_out, int sumsq_ld, const float* norm_sumsq, int norm_slabs, int norm_ld,
                         const void* norm_gamma, float norm_eps, void* stream) {
  if (!A || !B || !out) return fail(OPUS_ERR_ARG, "opus_gemm_bf16_fused: null pointer");
  GemmArgs a{};
  a.A = A; a.lda = lda; a.B = B; a.ldb = ldb;
  a.M = M; a.N = N; a.K = K;
  a.transposed = 1; a.epi = epilogue;
  a.out = out; a.ldo = ldo; a.bias = bias; a.residual = residual; a.ldr = ldr;
  a.split_k = split_k; a.splitk_fixup = splitk_fixup;
  a.sumsq_out = sumsq_out; a.sumsq_ld = sumsq_ld;
  a.norm_sumsq = norm_sumsq; a.norm_slabs = norm_slabs; a.norm_ld = norm_ld;
  a.norm_gamma = norm_gamma; a.norm_eps = norm_eps;
  RET(gemm_bf16(a, ST(stream)), "opus_gemm_bf16_fused");
}

int opus_gemm_suggest_split_k(int M, int N, int K, int transposed) {
  return gemm_pick_split_k(M, N, K, gemm_pick_bn(N, transposed));
}

int opus_splitk_reduce_bf16(const float* partial, int n_partial, const float* bias, void* out, int rows, int cols,
                            int ldo, int gelu, void* stream) {
  RET(splitk_reduce_bf16(partial, n_partial, bias, static_cast<bf16*>(out), rows, cols, ldo, gelu, ST(stream)),
      "opus_splitk_reduce_bf16");
}

int opus_esm_embed(const int32_t* tok, const float* scale, const float* table, float* x, int n_tok, int dim,
                   void* stream) {
  RET(esm_embed(tok, scale, table, x, n_tok, dim, ST(stream)), "opus_esm_embed");
}

int opus_layernorm_f32_bf16(float* x, const void* delta, const float* gamma, const float* beta, void* y, int rows,
                            int cols, float eps, void* stream) {
  RET(layernorm_f32_bf16(x, static_cast<const bf16*>(delta), gamma, beta, static_cast<bf16*>(y), rows, cols, eps,
                         ST(stream)),
      "opus_layernorm_f32_bf16");
}

int opus_rmsnorm_bf16(const void* x, const float* partial, int n_partial, const void* residual, void* h_out,
                      const void* w, void* y, int rows, int cols, float eps, void* stream) {
  RET(rmsnorm_bf16(static_cast<const bf16*>(x), partial, n_partial, static_cast<const bf16*>(residual),
                   static_cast<bf16*>(h_out), static_cast<const bf16*>(w), static_cast<bf16*>(y), rows, cols, eps,
                   ST(stream)),
      "opus_rmsnorm_bf16");
}

int opus_rope_esm_bf16(void* qkv, const int32_t* pos, const float* cos_t, const float* sin_t, int n_tok, int n_heads,
                       int head_dim, int ld, float q_scale, void* stream) {
  RET(rope_esm(static_cast<bf16*>(qkv), pos, cos_t, sin_t, n_tok, n_heads, head_dim, ld, q_scale, ST(stream)),
      "opus_rope_esm_bf16");
}

int opus_rope_llama_kvappend_bf16(void* qkv, const float* partial, int n_partial, const int32_t* pos,
                                  const int32_t* slot, const void* cos_t, const void* sin_t, void* kcache, void* vcache,
                                  int n_tok, int n_q_heads, int n_kv_heads, int head_dim, int ld, int block_size,
                                  void* stream) {
  RET(rope_llama_kvappend(static_cast<bf16*>(qkv), partial, n_partial, pos, slot, static_cast<const bf16*>(cos_t),
                          static_cast<const bf16*>(sin_t), static_cast<bf16*>(kcache), static_cast<bf16*>(vcache),
                          n_tok, n_q_heads, n_kv_heads, head_dim, ld, block_size, ST(stream)),
      "opus_rope_llama_kvappend_bf16");
}

int opus_final_ln_meanpool(const float* x, const void* delta, const int32_t* cu_seqlens, const float* gamma,
                           const float* beta, float* pooled, void* pooled_l2, float* hidden_out, int n_seqs, int dim,
                           float eps, void* stream) {
  RET(final_ln_meanpool(x, static_cast<const bf16*>(delta), cu_seqlens, gamma, beta, pooled, static_cast<bf16*>(pooled_l2), hidden_out, n_seqs, dim,
                        eps, ST(stream)),
      "opus_final_ln_meanpool");
}

int opus_l2norm_f32_bf16(const float* x, void* y, int rows, int dim, void* stream) {
  RET(l2norm_f32_bf16(x, static_cast<bf16*>(y), rows, dim, ST(stream)), "opus_l2norm_f32_bf16");
}

int opus_splice_gather_bf16(const int32_t* src, const void* embed, const void* soft, void* out, int n_rows, int dim,
                            void* stream) {
  RET(splice_gather(src, static_cast<const bf16*>(embed), static_cast<const bf16*>(soft), static_cast<bf16*>(out),
                    n_rows, dim, ST(stream)),
      "opus_splice_gather_bf16");
}

int opus_argmax_eos(const void* logits, int ld, int vocab, int n_rows, int32_t* finished, const int32_t* eos_ids,
                    int n_eos, int pad_id, int32_t* next_tok, int32_t* out_ids, int out_ld, int step,
                    int32_t* n_unfinished, void* stream) {
  RET(argmax_eos(static_cast<const bf16*>(logits), ld, vocab, n_rows, finished, eos_ids, n_eos, pad_id, next_tok,
                 out_ids, out_ld, step, n_unfinished, ST(stream), nullptr),
      "opus_argmax_eos");
}

int opus_sample_top_p(const void* logits, int ld, int vocab, int n_rows, float temperature, float top_p, uint64_t seed,
                      int32_t* finished, const int32_t* eos_ids, int n_eos, int pad_id, int32_t* next_tok,
                      int32_t* out_ids, int out_ld, int step, int32_t* n_unfinished, int32_t* kept_count,
                      void* stream) {
  RET(sample_top_p(static_cast<const bf16*>(logits), ld, vocab, n_rows, temperature, top_p, seed, finished, eos_ids,
                   n_eos, pad_id, next_tok, out_ids, out_ld, step, n_unfinished, ST(stream), nullptr, kept_count),
      "opus_sample_top_p");
}

int opus_cross_entropy_bf16(const void* logits, int ld, int vocab, const int32_t* target, float* loss, int n_rows,
                            void* stream) {
  RET(cross_entropy_rows(static_cast<const bf16*>(logits), ld, vocab, target, loss, n_rows, ST(stream)),
      "opus_cross_entropy_bf16");
}

int opus_embed_gather_bf16(const int32_t* tok, const void* table, void* x, int n_rows, int dim, void* stream) {
  RET(embed_gather(tok, static_cast<const bf16*>(table), static_cast<bf16*>(x), n_rows, dim, ST(stream)),
      "opus_embed_gather_bf16");
}

int opus_layernorm_bf16(const void* x, const float* partial, int n_partial, const float* red_bias, const void* residual,
                        void* h_out, const float* gamma, const float* beta, void* y, int rows, int cols, float eps,
                        void* stream) {
  RET(layernorm_bf16(static_cast<const bf16*>(x), partial, n_partial, red_bias, static_cast<const bf16*>(residual),
                     static_cast<bf16*>(h_out), gamma, beta, static_cast<bf16*>(y), rows, cols, eps, ST(stream)),
      "opus_layernorm_bf16");
}

int opus_add_pos_embed_bf16(void* h, const void* table, const int32_t* pos, int offset, int table_rows, int n_rows,
                            int dim, void* stream) {
  RET(add_pos_embed(static_cast<bf16*>(h), static_cast<const bf16*>(table), pos, offset, table_rows, n_rows, dim,
                    ST(stream)),
      "opus_add_pos_embed_bf16");
}

int opus_stop_sequences(const int32_t* out_ids, int out_ld, int n_rows, int step, const int32_t* step_ptr,
                        const int32_t* stop_seqs, const int32_t* stop_lens, int n_stop, int stop_ld, int32_t* finished,
                        int32_t* n_unfinished, void* stream) {
  RET(stop_sequences(out_ids, out_ld, n_rows, step, step_ptr, stop_seqs, stop_lens, n_stop, stop_ld, finished,
                     n_unfinished, ST(stream)),
      "opus_stop_sequences");
}

int opus_lora_merge_bf16(void* W, const void* A, const void* B, int out_features, int in_features, int r, float scale,
                         void* stream) {
  RET(lora_merge(static_cast<bf16*>(W), static_cast<const bf16*>(A), static_cast<const bf16*>(B), out_features,
                 in_features, r, scale, ST(stream)),
      "opus_lora_merge_bf16");
}

int opus_attn_varlen_bf16(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, void* o, int ldo,
                          const int32_t* cu_seqlens, int n_seqs, int n_tok, int max_len, int n_q_heads, int n_kv_heads,
                          int head_dim, int causal, float scale, void* stream) {
  RET(attn_varlen(static_cast<const bf16*>(q), ldq, static_cast<const bf16*>(k), ldk, static_cast<const bf16*>(v), ldv,
                  static_cast<bf16*>(o), ldo, cu_seqlens, n_seqs, n_tok, max_len, n_q_heads, n_kv_heads, head_dim, causal,
                  scale, ST(stream)),
      "opus_attn_varlen_bf16");
}

int opus_attn_decode_paged_bf16(const void* q, int ldq, const void* kcache, const void* vcache,
                                const int32_t* block_table, int max_blocks, const int32_t* ctx_len, void* o, int ldo,
                                int n_seqs, int n_q_heads, int n_kv_heads, int head_dim, int block_size, float scale,
                                void* stream) {
  RET(attn_decode_paged(static_cast<const bf16*>(q), ldq, static_cast<const bf16*>(kcache),
                        static_cast<const bf16*>(vcache), block_table, max_blocks, ctx_len, static_cast<bf16*>(o), ldo,
                        n_seqs, n_q_heads, n_kv_heads, head_dim, block_size, scale, ST(stream)),
      "opus_attn_decode_paged_bf16");
}

int opus_esm2_forward(const opus_esm2_model* model, const opus_esm2_workspace* ws, const int32_t* tokens,
                      const float* tok_scale, const int32_t* pos, const int32_t* cu_seqlens, int n_seqs, int n_tok,
                      int max_len, float* pooled, void* pooled_l2, float* hidden_out, void* stream) {
  return esm2_forward(model, ws, tokens, tok_scale, pos, cu_seqlens, n_seqs, n_tok, max_len, pooled, pooled_l2,
                      hidden_out, ST(stream));
}

int opus_projector_forward(const opus_projector_model* model, const void* x_l2, int n, void* cstp_out, void* h0,
                           void* out, float* splitk_ws, size_t splitk_ws_bytes, void* stream) {
  return projector_forward(model, x_l2, n, cstp_out, h0, out, splitk_ws, splitk_ws_bytes, ST(stream));
}

int opus_llama_prefill(const opus_llama_model* model, const opus_kv_cache* cache, const opus_llama_workspace* ws,
                       const void* embeds, const int32_t* pos, const int32_t* slot, const int32_t* cu_seqlens,
                       const int32_t* last_rows, int n_seqs, int n_tok, int max_len, void* stream) {
  return llama_prefill(model, cache, ws, embeds, pos, slot, cu_seqlens, last_rows, n_seqs, n_tok, max_len, ST(stream));
}

int opus_llama_decode_step(const opus_llama_model* model, const opus_kv_cache* cache, const opus_llama_workspace* ws,
                           const opus_decode_state* state, int n_seqs, void* stream) {
  return llama_decode_step(model, cache, ws, state, n_seqs, ST(stream));
}

int opus_llama_select(const opus_llama_model* model, const opus_llama_workspace* ws, const opus_decode_state* state,
                      int n_seqs, void* stream) {
  RET(llama_select(model, ws, state, n_seqs, ST(stream)), "opus_llama_select");
}

int opus_llama_decode_loop(const opus_llama_model* model, const opus_kv_cache* cache, const opus_llama_workspace* ws,
                           const opus_decode_state* state, int n_seqs, int n_steps, int check_every, int use_graph,
                           void* stream) {
  return llama_decode_loop(model, cache, ws, state, n_seqs, n_steps, check_every, use_graph, ST(stream));
}

int opus_release_graphs(void) { return release_graphs(); }

int opus_set_tunable(const char* name, int value) { return set_tunable(name, value); }

int opus_chain_trace(int enable, unsigned long long* out, int cap_words) {
  return gemm_chain_trace(enable, out, cap_words);
}

int opus_trace_begin(void* stream) { return trace_begin(ST(stream)); }

int opus_trace_end(char* buf, int cap) { return trace_end(buf, cap); }

long long opus_launch_count(int reset) { return launch_count(reset != 0); }

}  // extern "C"
