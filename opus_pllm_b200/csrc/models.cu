// Host-side orchestration of the three composite forwards (ESM-2 encoder, projectors, Llama prefill / decode) on top of
// the kernels in gemm_tcgen05.cu / bandwidth.cu / attention.cu. Pure launch sequencing: no allocation, no sync.
#include <atomic>
#include <map>
#include <mutex>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "common.h"
#include "context.h"
#include "gemm.h"
#include "kernels.h"
#include "models.h"

namespace opus {

#define OPUS_TRY(expr)                                   \
  do {                                                   \
    const int _rc = (expr);                              \
    if (_rc != OPUS_OK) return fail(_rc, #expr);         \
    if (g_trace_on) trace_mark(#expr, st);               \
  } while (0)

// ---- optional in-situ kernel timeline (debug / profiling aid; off by default, never active inside graph capture) ----
bool g_trace_on = false;
namespace {
struct TraceMark { const char* label; cudaEvent_t ev; };
std::vector<TraceMark> g_trace;
}  // namespace
void trace_mark(const char* label, cudaStream_t st) {
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) return;
  cudaEvent_t ev;
  if (cudaEventCreate(&ev) != cudaSuccess) return;
  cudaEventRecord(ev, st);
  g_trace.push_back({label, ev});
}
int trace_begin(cudaStream_t st) {
  for (auto& m : g_trace) cudaEventDestroy(m.ev);
  g_trace.clear();
  g_trace_on = true;
  trace_mark("begin", st);
  return OPUS_OK;
}
// writes "label<TAB>microseconds\n" per mark (time since the previous mark) into buf; returns bytes written
int trace_end(char* buf, int cap) {
  g_trace_on = false;
  if (g_trace.empty()) return 0;
  cudaEventSynchronize(g_trace.back().ev);
  int n = 0;
  for (size_t i = 1; i < g_trace.size(); ++i) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, g_trace[i - 1].ev, g_trace[i].ev);
    char label[48];
    snprintf(label, sizeof(label), "%.40s", g_trace[i].label);
    for (char* c = label; *c; ++c) if (*c == '(') { *c = 0; break; }
    const int w = snprintf(buf + n, cap > n ? cap - n : 0, "%s\t%.2f\n", label, ms * 1e3f);
    if (w < 0 || n + w >= cap) break;
    n += w;
  }
  for (auto& m : g_trace) cudaEventDestroy(m.ev);
  g_trace.clear();
  return n;
}

namespace {

using bf16 = __nv_bfloat16;

// Next-weight L2 prefetch hint for a chain of swap-AB GEMMs (see GemmArgs::pf_w).
struct Prefetch {
  const void* w = nullptr;
  int rows = 0, K = 0, split_k = 1, depth = 0;
};
void apply_prefetch(GemmArgs& a, const Prefetch* pf) {
  if (pf == nullptr || pf->w == nullptr || pf->depth <= 0) return;
  a.pf_w = pf->w; a.pf_rows = pf->rows; a.pf_K = pf->K; a.pf_split_k = pf->split_k; a.pf_depth = pf->depth;
}
// k-blocks to prefetch per work item for the decode chain (default 0 = off; OPUS_PF_<QKV|O|GU|DOWN|LM>=n at context
// creation, opus_set_tunable("pf_qkv", n) etc. at run time; tools/bench_decode.py sweeps them).
enum { PF_QKV = 0, PF_O = 1, PF_GU = 2, PF_DOWN = 3, PF_LM = 4 };
int pf_depth_for(int which) { return ctx().tun.pf_depth[which]; }

// fused decode chain (gemm_chain): opt-in with OPUS_DECODE_FUSED=1 / opus_set_tunable("decode_fused", 1).
bool decode_fused() { return ctx().tun.decode_fused != 0; }

// activations [rows, K] x weight [N, K]^T -> out, choosing the weight-streaming (swap-AB) form for small `rows`.
int linear(const void* x, int rows, const void* w, int N, int K, int epi, void* out, int ldo, const float* bias,
           const void* residual, int ldr, float* ws, size_t ws_bytes, cudaStream_t st, const Prefetch* pf = nullptr,
           bool decode = false) {
  GemmArgs a{};
  apply_prefetch(a, pf);
  a.K = K;
  a.epi = epi;
  a.out = out; a.ldo = ldo;
  a.bias = bias;
  a.residual = residual; a.ldr = ldr;
  if (rows <= 256) {
    a.transposed = 1;
    a.A = w; a.lda = K; a.M = N;
    a.B = x; a.ldb = K; a.N = rows;
  } else {
    a.transposed = 0;
    a.A = x; a.lda = K; a.M = rows;
    a.B = w; a.ldb = K; a.N = N;
    // decode at batch 257..512 (gate/up: 448 tiles = 3 waves + 4): the partial last wave is cut along K. The encoder and
    // the prefill keep one summation order for every token row (a token's result does not depend on its batch).
    a.streamk_tail = decode && rows <= 512;
  }
  (void)ws; (void)ws_bytes;
  return gemm_bf16(a, st);
}

// swap-AB GEMM with split-K partials into `partial` ([s][rows][N] fp32). Returns the split count via *splits.
int splitk_for(int rows, int N, int K, size_t partial_bytes) {
  const int bn = gemm_pick_bn(rows, 1);
  int s = (rows > 256 && rows <= 512 && ctx().tun.gemm_2cta_tr) ? gemm_pick_split_k_wide(N, rows, K)
                                                                : gemm_pick_split_k(N, rows, K, bn);
  while (s > 1 && gemm_workspace_bytes(rows, N, s) > partial_bytes) --s;
  return s;
}

int linear_splitk(const void* x, int rows, const void* w, int N, int K, float* partial, size_t partial_bytes,
                  int* splits, cudaStream_t st, const Prefetch* pf = nullptr) {
  const int s = splitk_for(rows, N, K, partial_bytes);
  if (gemm_workspace_bytes(rows, N, s) > partial_bytes) return OPUS_ERR_ARG;
  GemmArgs a{};
  apply_prefetch(a, pf);
  a.transposed = 1;
  a.A = w; a.lda = K; a.M = N;
  a.B = x; a.ldb = K; a.N = rows;
  a.K = K;
  a.epi = EPI_PARTIAL_F32;
  a.out = partial; a.ldo = N;
  a.split_k = s;
  *splits = s;
  return gemm_bf16(a, st);
}

}  // namespace

// ------------------------------------------------------------------------------------------------ ESM-2
int esm2_forward(const opus_esm2_model* m, const opus_esm2_workspace* ws, const int* tokens, const float* tok_scale,
                 const int* pos, const int* cu_seqlens, int n_seqs, int n_tok, int max_len, float* pooled,
                 void* pooled_l2, float* hidden_out, cudaStream_t st) {
  if (!m || !ws || !m->layers) return fail(OPUS_ERR_ARG, "esm2_forward: null model/workspace");
  if (n_tok <= 0 || n_seqs <= 0) return OPUS_OK;
  const int d = m->dim, hd = d / m->n_heads, ffn = m->ffn_dim;
  if (hd != 64) return fail(OPUS_ERR_ARG, "esm2_forward: head_dim must be 64");
  OPUS_TRY(esm_embed(tokens, tok_scale, m->embed, ws->x, n_tok, d, st));
  bf16* xn = static_cast<bf16*>(ws->xn);
  bf16* qkv = static_cast<bf16*>(ws->qkv);
  bf16* attn = static_cast<bf16*>(ws->attn);
  bf16* ffb = static_cast<bf16*>(ws->ffn);
  // The fp32 residual stream x is only touched by the LayerNorm kernels: out_proj / fc2 write their bf16 result
  // ("delta") into the xn buffer (free once the GEMM that read it has run) and the NEXT LayerNorm folds it into x.
  // That keeps the GEMM epilogues light (bf16 stores, no fp32 read-modify-write with row-per-thread access) and
  // matches the reference's autocast rounding (half-precision Linear output added to the fp32 stream).
  const bf16* pending = nullptr;
  for (int l = 0; l < m->n_layers; ++l) {
    const opus_esm2_layer& L = m->layers[l];
    OPUS_TRY(layernorm_f32_bf16(ws->x, pending, L.ln1_g, L.ln1_b, xn, n_tok, d, m->ln_eps, st));
    {
      // q|k|v projection with the rotary embedding applied in the GEMM epilogue (a 64-wide head is exactly one store box
      // of the row-owning epilogue thread); the separate RoPE kernel remains for shapes the fused path does not take
      GemmArgs a{};
      a.transposed = 0;
      a.A = xn; a.lda = d; a.M = n_tok;
      a.B = L.wqkv; a.ldb = d; a.N = 3 * d;
      a.K = d;
      a.epi = EPI_BF16;
      a.out = qkv; a.ldo = 3 * d;
      a.bias = L.bqkv;
      a.rope_pos = pos; a.rope_cos = m->rope_cos; a.rope_sin = m->rope_sin;
      a.rope_cols = 2 * d; a.rope_q_cols = d; a.rope_q_scale = 0.125f;
      if (n_tok > 256 && gemm_fuses_rope(a)) {
        OPUS_TRY(gemm_bf16(a, st));
      } else {
        OPUS_TRY(linear(xn, n_tok, L.wqkv, 3 * d, d, EPI_BF16, qkv, 3 * d, L.bqkv, nullptr, 0, nullptr, 0, st));
        OPUS_TRY(rope_esm(qkv, pos, m->rope_cos, m->rope_sin, n_tok, m->n_heads, hd, 3 * d, 0.125f, st));
      }
    }
    OPUS_TRY(attn_varlen(qkv, 3 * d, qkv + d, 3 * d, qkv + 2 * d, 3 * d, attn, d, cu_seqlens, n_seqs, n_tok, max_len,
                         m->n_heads, m->n_heads, hd, 0, 1.0f, st));
    OPUS_TRY(linear(attn, n_tok, L.wo, d, d, EPI_BF16, xn, d, L.bo, nullptr, 0, nullptr, 0, st));
    OPUS_TRY(layernorm_f32_bf16(ws->x, xn, L.ln2_g, L.ln2_b, xn, n_tok, d, m->ln_eps, st));
    OPUS_TRY(linear(xn, n_tok, L.w1, ffn, d, EPI_BF16_GELU, ffb, ffn, L.b1, nullptr, 0, nullptr, 0, st));
    OPUS_TRY(linear(ffb, n_tok, L.w2, d, ffn, EPI_BF16, xn, d, L.b2, nullptr, 0, nullptr, 0, st));
    pending = xn;
  }
  OPUS_TRY(final_ln_meanpool(ws->x, pending, cu_seqlens, m->lnf_g, m->lnf_b, pooled, static_cast<bf16*>(pooled_l2),
                             hidden_out, n_seqs, d, m->ln_eps, st));
  return OPUS_OK;
}

// ------------------------------------------------------------------------------------------------ projectors
int projector_forward(const opus_projector_model* m, const void* x_l2, int n, void* cstp_out, void* h0, void* out,
                      float* ws, size_t ws_bytes, cudaStream_t st) {
  if (!m) return fail(OPUS_ERR_ARG, "projector_forward: null model");
  if (n <= 0) return OPUS_OK;
  const void* cur = x_l2;
  int cur_dim = m->in_dim;
  if (m->w_cstp != nullptr) {
    OPUS_TRY(linear(cur, n, m->w_cstp, m->cstp_dim, cur_dim, EPI_BF16, cstp_out, m->cstp_dim, m->b_cstp, nullptr, 0, ws,
                    ws_bytes, st));
    cur = cstp_out;
    cur_dim = m->cstp_dim;
  }
  if (m->w2 == nullptr) {  // 'linear' projector type
    OPUS_TRY(linear(cur, n, m->w0, m->hidden_dim, cur_dim, EPI_BF16, out, m->hidden_dim, m->b0, nullptr, 0, ws,
                    ws_bytes, st));
    return OPUS_OK;
  }
  OPUS_TRY(linear(cur, n, m->w0, m->hidden_dim, cur_dim, EPI_BF16_GELU, h0, m->hidden_dim, m->b0, nullptr, 0, ws,
                  ws_bytes, st));
  OPUS_TRY(linear(h0, n, m->w2, m->hidden_dim, m->hidden_dim, EPI_BF16, out, m->hidden_dim, m->b2, nullptr, 0, ws,
                  ws_bytes, st));
  return OPUS_OK;
}

namespace {
// softmax scale: 1/sqrt of the model's REAL head width. Models whose heads are narrower than the 128 columns the
// attention / cache kernels work on (Qwen2-0.5B, OPT-125m ... 2.7B, Galactica-1.3B: 64 or 80) are loaded with every head
// zero-padded to 128 columns (llama.py / opt.py); head_dim_real then carries the width the scores are scaled by.
inline float attn_scale(const opus_llama_model* m) {
  return 1.0f / sqrtf((float)(m->head_dim_real > 0 ? m->head_dim_real : m->head_dim));
}
// OPT / Galactica family (defined below, after llama_select)
int opt_prefill(const opus_llama_model* m, const opus_kv_cache* kv, const opus_llama_workspace* ws, const void* embeds,
                const int* pos, const int* slot, const int* cu_seqlens, const int* last_rows, int n_seqs, int n_tok,
                int max_len, cudaStream_t st);
int opt_decode_step(const opus_llama_model* m, const opus_kv_cache* kv, const opus_llama_workspace* ws,
                    const opus_decode_state* s, int B, cudaStream_t st);
}  // namespace

// ------------------------------------------------------------------------------------------------ Llama prefill
int llama_prefill(const opus_llama_model* m, const opus_kv_cache* kv, const opus_llama_workspace* ws,
                  const void* embeds, const int* pos, const int* slot, const int* cu_seqlens, const int* last_rows,
                  int n_seqs, int n_tok, int max_len, cudaStream_t st) {
  if (!m || !kv || !ws || !m->layers) return fail(OPUS_ERR_ARG, "llama_prefill: null argument");
  if (n_tok <= 0 || n_seqs <= 0) return OPUS_OK;
  if (m->arch == OPUS_ARCH_OPT)
    return opt_prefill(m, kv, ws, embeds, pos, slot, cu_seqlens, last_rows, n_seqs, n_tok, max_len, st);
  const int d = m->dim, hd = m->head_dim, Hq = m->n_q_heads, Hkv = m->n_kv_heads, ffn = m->ffn_dim;
  const int qkv_n = (Hq + 2 * Hkv) * hd;
  if (hd != 128) return fail(OPUS_ERR_ARG, "llama_prefill: head_dim must be 128");
  bf16* h = static_cast<bf16*>(ws->h);
  bf16* xn = static_cast<bf16*>(ws->xn);
  bf16* qkv = static_cast<bf16*>(ws->qkv);
  bf16* attn = static_cast<bf16*>(ws->attn);
  bf16* act = static_cast<bf16*>(ws->act);
  if (embeds != ws->h) {
    if (cudaMemcpyAsync(h, embeds, (size_t)n_tok * d * sizeof(bf16), cudaMemcpyDeviceToDevice, st) != cudaSuccess)
      return fail(OPUS_ERR_CUDA, "llama_prefill: copy embeds");
    note_launch();
  }
  const size_t layer_stride = (size_t)kv->num_blocks * Hkv * kv->block_size * hd;
  const float scale = attn_scale(m);
  for (int l = 0; l < m->n_layers; ++l) {
    const opus_llama_layer& L = m->layers[l];
    bf16* kc = static_cast<bf16*>(kv->k) + (size_t)l * layer_stride;
    bf16* vc = static_cast<bf16*>(kv->v) + (size_t)l * layer_stride;
    OPUS_TRY(rmsnorm_bf16(h, nullptr, 0, nullptr, nullptr, static_cast<const bf16*>(L.ln1_w), xn, n_tok, d, m->rms_eps,
                          st));
    {
      // q|k|v projection with RoPE and the paged KV append applied in the GEMM epilogue (K = 4096 leaves the epilogue
      // idle most of the time); the separate kernel remains for shapes the fused path does not take
      GemmArgs a{};
      a.transposed = 0;
      a.A = xn; a.lda = d; a.M = n_tok;
      a.B = L.wqkv; a.ldb = d; a.N = qkv_n;
      a.K = d;
      a.epi = EPI_BF16;
      a.out = qkv; a.ldo = qkv_n;
      a.bias = L.bqkv;
      a.rl_pos = pos; a.rl_slot = slot; a.rl_cos = m->rope_cos; a.rl_sin = m->rope_sin;
      a.rl_kcache = kc; a.rl_vcache = vc;
      a.rl_hq = Hq; a.rl_hkv = Hkv; a.rl_bs = kv->block_size;
      if (n_tok > 256 && gemm_fuses_rope(a)) {
        OPUS_TRY(gemm_bf16(a, st));
      } else {
        OPUS_TRY(linear(xn, n_tok, L.wqkv, qkv_n, d, EPI_BF16, qkv, qkv_n, L.bqkv, nullptr, 0, nullptr, 0, st));
        OPUS_TRY(rope_llama_kvappend(qkv, nullptr, 0, pos, slot, static_cast<const bf16*>(m->rope_cos),
                                     static_cast<const bf16*>(m->rope_sin), kc, vc, n_tok, Hq, Hkv, hd, qkv_n,
                                     kv->block_size, st));
      }
    }
    OPUS_TRY(attn_varlen(qkv, qkv_n, qkv + Hq * hd, qkv_n, qkv + (Hq + Hkv) * hd, qkv_n, attn, Hq * hd, cu_seqlens,
                         n_seqs, n_tok, max_len, Hq, Hkv, hd, 1, scale, st));
    OPUS_TRY(linear(attn, n_tok, L.wo, d, Hq * hd, EPI_RES_BF16, h, d, nullptr, h, d, nullptr, 0, st));
    OPUS_TRY(rmsnorm_bf16(h, nullptr, 0, nullptr, nullptr, static_cast<const bf16*>(L.ln2_w), xn, n_tok, d, m->rms_eps,
                          st));
    OPUS_TRY(linear(xn, n_tok, L.wgu, 2 * ffn, d, EPI_SWIGLU, act, ffn, nullptr, nullptr, 0, nullptr, 0, st));
    OPUS_TRY(linear(act, n_tok, L.wdown, d, ffn, EPI_RES_BF16, h, d, nullptr, h, d, nullptr, 0, st));
  }
  // last token of every sequence -> final norm -> lm_head
  bf16* last_h = static_cast<bf16*>(ws->last_h);
  OPUS_TRY(embed_gather(last_rows, h, last_h, n_seqs, d, st));
  OPUS_TRY(rmsnorm_bf16(last_h, nullptr, 0, nullptr, nullptr, static_cast<const bf16*>(m->norm_w), last_h, n_seqs, d,
                        m->rms_eps, st));
  OPUS_TRY(linear(last_h, n_seqs, m->lm_head, m->vocab, d, EPI_BF16, ws->logits, m->vocab, nullptr, nullptr, 0, nullptr,
                  0, st));
  return OPUS_OK;
}

int llama_select(const opus_llama_model* m, const opus_llama_workspace* ws, const opus_decode_state* s, int n_seqs,
                 cudaStream_t st) {
  if (s->do_sample)
    OPUS_TRY(sample_top_p(static_cast<const bf16*>(ws->logits), m->vocab, m->vocab, n_seqs, s->temperature, s->top_p,
                          s->seed, s->finished, s->eos_ids, s->n_eos, s->pad_id, s->next_tok, s->out_ids, s->out_ld, -1,
                          s->n_unfinished, st, s->step, nullptr,
                          reinterpret_cast<const unsigned long long*>(s->seed_ptr)));
  else
    OPUS_TRY(argmax_eos(static_cast<const bf16*>(ws->logits), m->vocab, m->vocab, n_seqs, s->finished, s->eos_ids,
                        s->n_eos, s->pad_id, s->next_tok, s->out_ids, s->out_ld, -1, s->n_unfinished, st, s->step));
  if (s->n_stop > 0 && s->stop_seqs != nullptr)
    OPUS_TRY(stop_sequences(s->out_ids, s->out_ld, n_seqs, -1, s->step, s->stop_seqs, s->stop_lens, s->n_stop,
                            s->stop_ld, s->finished, s->n_unfinished, st));
  return OPUS_OK;
}

// ------------------------------------------------------------------------------------------------ OPT / Galactica
// Sibling decoder family of language_model/opus_opt.py (HF OPTDecoder, do_layer_norm_before): the same GEMM / attention /
// paged-cache kernels as the Llama path with LayerNorm instead of RMSNorm, biased linears, a plain fc1 -> act -> fc2 MLP
// and learned positions. The "rotary" tables of an OPT model hold cos = 1 / sin = 0, so the fused RoPE + KV-append
// epilogues pass q and k through unchanged (x * 1 + rot(x) * 0 is exact in bf16) and only do the cache append.
namespace {
int opt_check(const opus_llama_model* m) {
  if (m->pos_embed == nullptr || m->pos_rows <= 0 || m->norm_g == nullptr)
    return fail(OPUS_ERR_ARG, "OPT family: pos_embed / final_layer_norm missing");
  if (m->n_q_heads != m->n_kv_heads) return fail(OPUS_ERR_ARG, "OPT family: multi-head attention only");
  if (m->head_dim != 128) return fail(OPUS_ERR_ARG, "OPT family: head_dim must be 128");
  return OPUS_OK;
}
inline int opt_fc1_epi(const opus_llama_model* m) { return m->opt_act == 1 ? EPI_BF16_GELU : EPI_BF16_RELU; }

int opt_prefill(const opus_llama_model* m, const opus_kv_cache* kv, const opus_llama_workspace* ws, const void* embeds,
                const int* pos, const int* slot, const int* cu_seqlens, const int* last_rows, int n_seqs, int n_tok,
                int max_len, cudaStream_t st) {
  OPUS_TRY(opt_check(m));
  const int d = m->dim, hd = m->head_dim, H = m->n_q_heads, ffn = m->ffn_dim;
  const int aw = H * hd;          // width of the head space (> dim when narrow heads are padded to 128 columns)
  const int qkv_n = 3 * aw;
  bf16* h = static_cast<bf16*>(ws->h);
  bf16* xn = static_cast<bf16*>(ws->xn);
  bf16* qkv = static_cast<bf16*>(ws->qkv);
  bf16* attn = static_cast<bf16*>(ws->attn);
  bf16* act = static_cast<bf16*>(ws->act);
  if (embeds != ws->h) {
    if (cudaMemcpyAsync(h, embeds, (size_t)n_tok * d * sizeof(bf16), cudaMemcpyDeviceToDevice, st) != cudaSuccess)
      return fail(OPUS_ERR_CUDA, "opt_prefill: copy embeds");
    note_launch();
  }
  OPUS_TRY(add_pos_embed(h, static_cast<const bf16*>(m->pos_embed), pos, 2, m->pos_rows, n_tok, d, st));
  const size_t layer_stride = (size_t)kv->num_blocks * H * kv->block_size * hd;
  const float scale = attn_scale(m);
  for (int l = 0; l < m->n_layers; ++l) {
    const opus_llama_layer& L = m->layers[l];
    bf16* kc = static_cast<bf16*>(kv->k) + (size_t)l * layer_stride;
    bf16* vc = static_cast<bf16*>(kv->v) + (size_t)l * layer_stride;
    OPUS_TRY(layernorm_bf16(h, nullptr, 0, nullptr, nullptr, nullptr, L.ln1_g, L.ln1_b, xn, n_tok, d, m->rms_eps, st));
    {
      GemmArgs a{};
      a.transposed = 0;
      a.A = xn; a.lda = d; a.M = n_tok;
      a.B = L.wqkv; a.ldb = d; a.N = qkv_n;
      a.K = d;
      a.epi = EPI_BF16;
      a.out = qkv; a.ldo = qkv_n;
      a.bias = L.bqkv;
      a.rl_pos = pos; a.rl_slot = slot; a.rl_cos = m->rope_cos; a.rl_sin = m->rope_sin;
      a.rl_kcache = kc; a.rl_vcache = vc;
      a.rl_hq = H; a.rl_hkv = H; a.rl_bs = kv->block_size;
      if (n_tok > 256 && gemm_fuses_rope(a)) {
        OPUS_TRY(gemm_bf16(a, st));
      } else {
        OPUS_TRY(linear(xn, n_tok, L.wqkv, qkv_n, d, EPI_BF16, qkv, qkv_n, L.bqkv, nullptr, 0, nullptr, 0, st));
        OPUS_TRY(rope_llama_kvappend(qkv, nullptr, 0, pos, slot, static_cast<const bf16*>(m->rope_cos),
                                     static_cast<const bf16*>(m->rope_sin), kc, vc, n_tok, H, H, hd, qkv_n,
                                     kv->block_size, st));
      }
    }
    OPUS_TRY(attn_varlen(qkv, qkv_n, qkv + aw, qkv_n, qkv + 2 * aw, qkv_n, attn, aw, cu_seqlens, n_seqs, n_tok, max_len, H,
                         H, hd, 1, scale, st));
    OPUS_TRY(linear(attn, n_tok, L.wo, d, aw, EPI_RES_BF16, h, d, L.bo, h, d, nullptr, 0, st));
    OPUS_TRY(layernorm_bf16(h, nullptr, 0, nullptr, nullptr, nullptr, L.ln2_g, L.ln2_b, xn, n_tok, d, m->rms_eps, st));
    OPUS_TRY(linear(xn, n_tok, L.wgu, ffn, d, opt_fc1_epi(m), act, ffn, L.b1, nullptr, 0, nullptr, 0, st));
    OPUS_TRY(linear(act, n_tok, L.wdown, d, ffn, EPI_RES_BF16, h, d, L.b2, h, d, nullptr, 0, st));
  }
  bf16* last_h = static_cast<bf16*>(ws->last_h);
  OPUS_TRY(embed_gather(last_rows, h, last_h, n_seqs, d, st));
  OPUS_TRY(layernorm_bf16(last_h, nullptr, 0, nullptr, nullptr, nullptr, m->norm_g, m->norm_b, last_h, n_seqs, d,
                          m->rms_eps, st));
  OPUS_TRY(linear(last_h, n_seqs, m->lm_head, m->vocab, d, EPI_BF16, ws->logits, m->vocab, nullptr, nullptr, 0, nullptr,
                  0, st));
  return OPUS_OK;
}

int opt_decode_step(const opus_llama_model* m, const opus_kv_cache* kv, const opus_llama_workspace* ws,
                    const opus_decode_state* s, int B, cudaStream_t st) {
  OPUS_TRY(opt_check(m));
  const int d = m->dim, hd = m->head_dim, H = m->n_q_heads, ffn = m->ffn_dim;
  const int aw = H * hd;          // width of the head space (> dim when narrow heads are padded to 128 columns)
  const int qkv_n = 3 * aw;
  bf16* h = static_cast<bf16*>(ws->h);
  bf16* xn = static_cast<bf16*>(ws->xn);
  bf16* qkv = static_cast<bf16*>(ws->qkv);
  bf16* attn = static_cast<bf16*>(ws->attn);
  bf16* act = static_cast<bf16*>(ws->act);
  const size_t layer_stride = (size_t)kv->num_blocks * H * kv->block_size * hd;
  const float scale = attn_scale(m);
  OPUS_TRY(decode_advance(s->ctx_len, s->pos, s->slot, s->block_table, s->max_blocks, kv->block_size, B, st, s->step));
  OPUS_TRY(embed_gather(s->next_tok, static_cast<const bf16*>(m->embed), h, B, d, st));
  OPUS_TRY(add_pos_embed(h, static_cast<const bf16*>(m->pos_embed), s->pos, 2, m->pos_rows, B, d, st));
  OPUS_TRY(layernorm_bf16(h, nullptr, 0, nullptr, nullptr, nullptr, m->layers[0].ln1_g, m->layers[0].ln1_b, xn, B, d,
                          m->rms_eps, st));
  for (int l = 0; l < m->n_layers; ++l) {
    const opus_llama_layer& L = m->layers[l];
    bf16* kc = static_cast<bf16*>(kv->k) + (size_t)l * layer_stride;
    bf16* vc = static_cast<bf16*>(kv->v) + (size_t)l * layer_stride;
    int sp = 1;
    // q|k|v: weight-streaming split-K GEMM; the attention CTAs reduce the partials of their heads (+ bias) and append K/V
    OPUS_TRY(linear_splitk(xn, B, L.wqkv, qkv_n, d, ws->partial, ws->partial_bytes, &sp, st));
    OPUS_TRY(attn_decode_paged_fused(qkv, qkv_n, ws->partial, sp, s->pos, s->slot,
                                     static_cast<const bf16*>(m->rope_cos), static_cast<const bf16*>(m->rope_sin), kc,
                                     vc, s->block_table, s->max_blocks, s->ctx_len, attn, aw, B, H, H, hd,
                                     kv->block_size, scale, st, L.bqkv));
    OPUS_TRY(linear_splitk(attn, B, L.wo, d, aw, ws->partial, ws->partial_bytes, &sp, st));
    OPUS_TRY(layernorm_bf16(nullptr, ws->partial, sp, L.bo, h, h, L.ln2_g, L.ln2_b, xn, B, d, m->rms_eps, st));
    OPUS_TRY(linear(xn, B, L.wgu, ffn, d, opt_fc1_epi(m), act, ffn, L.b1, nullptr, 0, nullptr, 0, st, nullptr, true));
    OPUS_TRY(linear_splitk(act, B, L.wdown, d, ffn, ws->partial, ws->partial_bytes, &sp, st));
    const bool last = l + 1 == m->n_layers;
    OPUS_TRY(layernorm_bf16(nullptr, ws->partial, sp, L.b2, h, h, last ? m->norm_g : m->layers[l + 1].ln1_g,
                            last ? m->norm_b : m->layers[l + 1].ln1_b, xn, B, d, m->rms_eps, st));
  }
  OPUS_TRY(linear(xn, B, m->lm_head, m->vocab, d, EPI_BF16, ws->logits, m->vocab, nullptr, nullptr, 0, nullptr, 0, st,
                  nullptr, true));
  OPUS_TRY(llama_select(m, ws, s, B, st));
  return OPUS_OK;
}
}  // namespace

// ------------------------------------------------------------------------------------------------ Llama decode step
int llama_decode_step(const opus_llama_model* m, const opus_kv_cache* kv, const opus_llama_workspace* ws,
                      const opus_decode_state* s, int B, cudaStream_t st) {
  if (!m || !kv || !ws || !s) return fail(OPUS_ERR_ARG, "llama_decode_step: null argument");
  if (B <= 0) return OPUS_OK;
  if (m->arch == OPUS_ARCH_OPT) return opt_decode_step(m, kv, ws, s, B, st);
  const int d = m->dim, hd = m->head_dim, Hq = m->n_q_heads, Hkv = m->n_kv_heads, ffn = m->ffn_dim;
  const int qkv_n = (Hq + 2 * Hkv) * hd;
  bf16* h = static_cast<bf16*>(ws->h);
  bf16* xn = static_cast<bf16*>(ws->xn);
  bf16* qkv = static_cast<bf16*>(ws->qkv);
  bf16* attn = static_cast<bf16*>(ws->attn);
  bf16* act = static_cast<bf16*>(ws->act);
  const size_t layer_stride = (size_t)kv->num_blocks * Hkv * kv->block_size * hd;
  const float scale = attn_scale(m);

  OPUS_TRY(decode_advance(s->ctx_len, s->pos, s->slot, s->block_table, s->max_blocks, kv->block_size, B, st, s->step));
  if (B <= 64 && ctx().tun.decode_norm_fused && ctx().tun.decode_rope_fused && (d % 64) == 0) {
    // Norm-fused form (5 launches per layer instead of 7): the RMSNorm kernels disappear.
    //   * o_proj / down reduce their split-K partial sums INSIDE the GEMM (the CTA holding a tile's last k-split adds the
    //     others' dumps in split order: same sums as the reduce kernel), apply the residual epilogue, write the residual
    //     stream h and, per 32-feature slab and batch row, the sum of squares of what they stored;
    //   * q|k|v and gate/up take the raw h as their activation operand: four extra warps rewrite every landed k-slice in
    //     shared memory as bf16(gamma * bf16(h * rstd)) -- the rounding points of rmsnorm_bf16_kernel -- before the
    //     tensor core reads it; rstd comes from the slab sums (fixed summation order).
    // The only difference from the kernel-per-op path is the order in which the squares of a row are added up.
    const int slabs = d / 32, sld = 64;
    if (ws->partial_bytes < gemm_workspace_bytes(B, qkv_n, 1) + (size_t)slabs * sld * sizeof(float))
      return fail(OPUS_ERR_ARG, "llama_decode_step: split-K workspace too small for the norm-fused path");
    float* sumsq = ws->partial + ws->partial_bytes / sizeof(float) - (size_t)slabs * sld;
    const size_t part_bytes = ws->partial_bytes - (size_t)slabs * sld * sizeof(float);
    OPUS_TRY(embed_gather_sumsq(s->next_tok, static_cast<const bf16*>(m->embed), h, sumsq, sld, B, d, st));
    auto normed = [&](GemmArgs& a, const void* gamma) {
      a.norm_sumsq = sumsq; a.norm_slabs = slabs; a.norm_ld = sld; a.norm_gamma = gamma; a.norm_eps = m->rms_eps;
    };
    auto reduced_residual = [&](const void* x, const void* w, int K, int split) {
      GemmArgs a{};
      a.transposed = 1;
      a.A = w; a.lda = K; a.M = d;
      a.B = x; a.ldb = K; a.N = B;
      a.K = K;
      a.epi = EPI_RES_BF16;
      a.out = h; a.ldo = d;
      a.residual = h; a.ldr = d;
      a.split_k = split; a.splitk_fixup = 1;
      a.sumsq_out = sumsq; a.sumsq_ld = sld;
      return gemm_bf16(a, st);
    };
    const int sp_o = splitk_for(B, d, Hq * hd, part_bytes), sp_down = splitk_for(B, d, ffn, part_bytes);
    for (int l = 0; l < m->n_layers; ++l) {
      const opus_llama_layer& L = m->layers[l];
      bf16* kc = static_cast<bf16*>(kv->k) + (size_t)l * layer_stride;
      bf16* vc = static_cast<bf16*>(kv->v) + (size_t)l * layer_stride;
      const int sp = splitk_for(B, qkv_n, d, part_bytes);
      {
        GemmArgs a{};
        a.transposed = 1;
        a.A = L.wqkv; a.lda = d; a.M = qkv_n;
        a.B = h; a.ldb = d; a.N = B;
        a.K = d;
        a.epi = EPI_PARTIAL_F32;
        a.out = ws->partial; a.ldo = qkv_n;
        a.split_k = sp;
        normed(a, L.ln1_w);
        OPUS_TRY(gemm_bf16(a, st));
      }
      OPUS_TRY(attn_decode_paged_fused(qkv, qkv_n, ws->partial, sp, s->pos, s->slot,
                                       static_cast<const bf16*>(m->rope_cos), static_cast<const bf16*>(m->rope_sin), kc,
                                       vc, s->block_table, s->max_blocks, s->ctx_len, attn, Hq * hd, B, Hq, Hkv, hd,
                                       kv->block_size, scale, st, L.bqkv));
      OPUS_TRY(reduced_residual(attn, L.wo, Hq * hd, sp_o));
      {
        GemmArgs a{};
        a.transposed = 1;
        a.A = L.wgu; a.lda = d; a.M = 2 * ffn;
        a.B = h; a.ldb = d; a.N = B;
        a.K = d;
        a.epi = EPI_SWIGLU;
        a.out = act; a.ldo = ffn;
        normed(a, L.ln2_w);
        OPUS_TRY(gemm_bf16(a, st));
      }
      OPUS_TRY(reduced_residual(act, L.wdown, ffn, sp_down));
    }
    OPUS_TRY(rmsnorm_bf16(h, nullptr, 0, nullptr, nullptr, static_cast<const bf16*>(m->norm_w), xn, B, d, m->rms_eps, st));
    OPUS_TRY(linear(xn, B, m->lm_head, m->vocab, d, EPI_BF16, ws->logits, m->vocab, nullptr, nullptr, 0, nullptr, 0, st));
    OPUS_TRY(llama_select(m, ws, s, B, st));
    return OPUS_OK;
  }
  OPUS_TRY(embed_gather(s->next_tok, static_cast<const bf16*>(m->embed), h, B, d, st));
  OPUS_TRY(rmsnorm_bf16(h, nullptr, 0, nullptr, nullptr, static_cast<const bf16*>(m->layers[0].ln1_w), xn, B, d,
                        m->rms_eps, st));
  if (B <= 256 && decode_fused()) {
    // Fused form: per layer  rope+append -> paged attention -> ONE chain kernel
    //   { o_proj (split-K) | reduce+residual+RMSNorm | gate/up+SwiGLU | down (split-K) | reduce+residual+RMSNorm |
    //     next layer's qkv (split-K)  or  lm_head }
    // i.e. 3 launches per layer instead of 8, with the weight stream running across the phase boundaries.
    const int sp_qkv = splitk_for(B, qkv_n, d, ws->partial_bytes);
    const int sp_o = splitk_for(B, d, Hq * hd, ws->partial_bytes);
    const int sp_down = splitk_for(B, d, ffn, ws->partial_bytes);
    auto gemm_phase = [&](ChainPhase& ph, const void* w, int N, int K, const void* x, int epi, void* out, int ldo,
                          int split) {
      ph = ChainPhase{};
      ph.kind = CHAIN_GEMM;
      GemmArgs& a = ph.gemm;
      a.transposed = 1;
      a.A = w; a.lda = K; a.M = N;
      a.B = x; a.ldb = K; a.N = B;
      a.K = K;
      a.epi = epi;
      a.out = out; a.ldo = ldo;
      a.split_k = split;
    };
    auto norm_phase = [&](ChainPhase& ph, int n_partial, const void* w) {
      ph = ChainPhase{};
      ph.kind = CHAIN_NORM;
      ph.norm.partial = ws->partial; ph.norm.n_partial = n_partial;
      ph.norm.residual = h; ph.norm.h_out = h;
      ph.norm.w = w; ph.norm.y = xn;
      ph.norm.rows = B; ph.norm.cols = d; ph.norm.eps = m->rms_eps;
    };
    {
      int sp = 1;
      OPUS_TRY(linear_splitk(xn, B, m->layers[0].wqkv, qkv_n, d, ws->partial, ws->partial_bytes, &sp, st));
    }
    for (int l = 0; l < m->n_layers; ++l) {
      const opus_llama_layer& L = m->layers[l];
      bf16* kc = static_cast<bf16*>(kv->k) + (size_t)l * layer_stride;
      bf16* vc = static_cast<bf16*>(kv->v) + (size_t)l * layer_stride;
      if (ctx().tun.decode_rope_fused) {
        OPUS_TRY(attn_decode_paged_fused(qkv, qkv_n, ws->partial, sp_qkv, s->pos, s->slot,
                                         static_cast<const bf16*>(m->rope_cos), static_cast<const bf16*>(m->rope_sin),
                                         kc, vc, s->block_table, s->max_blocks, s->ctx_len, attn, Hq * hd, B, Hq, Hkv,
                                         hd, kv->block_size, scale, st, L.bqkv));
      } else {
        OPUS_TRY(rope_llama_kvappend(qkv, ws->partial, sp_qkv, s->pos, s->slot, static_cast<const bf16*>(m->rope_cos),
                                     static_cast<const bf16*>(m->rope_sin), kc, vc, B, Hq, Hkv, hd, qkv_n,
                                     kv->block_size, st, L.bqkv));
        OPUS_TRY(attn_decode_paged(qkv, qkv_n, kc, vc, s->block_table, s->max_blocks, s->ctx_len, attn, Hq * hd, B, Hq,
                                   Hkv, hd, kv->block_size, scale, st));
      }
      ChainPhase ph[6];
      const bool last = l + 1 == m->n_layers;
      gemm_phase(ph[0], L.wo, d, Hq * hd, attn, EPI_PARTIAL_F32, ws->partial, d, sp_o);
      norm_phase(ph[1], sp_o, L.ln2_w);
      gemm_phase(ph[2], L.wgu, 2 * ffn, d, xn, EPI_SWIGLU, act, ffn, 1);
      gemm_phase(ph[3], L.wdown, d, ffn, act, EPI_PARTIAL_F32, ws->partial, d, sp_down);
      norm_phase(ph[4], sp_down, last ? m->norm_w : m->layers[l + 1].ln1_w);
      if (last) gemm_phase(ph[5], m->lm_head, m->vocab, d, xn, EPI_BF16, ws->logits, m->vocab, 1);
      else gemm_phase(ph[5], m->layers[l + 1].wqkv, qkv_n, d, xn, EPI_PARTIAL_F32, ws->partial, qkv_n, sp_qkv);
      OPUS_TRY(gemm_chain(ph, 6, st));
    }
    OPUS_TRY(llama_select(m, ws, s, B, st));
    return OPUS_OK;
  }
  // Each weight-streaming GEMM asks for the head of the NEXT GEMM's weight stream at its tail (L2 prefetch), so HBM
  // does not idle while it drains, the small kernels between run, and the next GEMM ramps up.
  const bool swap = B <= 256;
  Prefetch pf_o, pf_gu, pf_down, pf_next;
  if (swap) {
    pf_o.rows = d; pf_o.K = Hq * hd; pf_o.split_k = splitk_for(B, d, Hq * hd, ws->partial_bytes); pf_o.depth = pf_depth_for(PF_O);
    pf_gu.rows = 2 * ffn; pf_gu.K = d; pf_gu.split_k = 1; pf_gu.depth = pf_depth_for(PF_GU);
    pf_down.rows = d; pf_down.K = ffn; pf_down.split_k = splitk_for(B, d, ffn, ws->partial_bytes); pf_down.depth = pf_depth_for(PF_DOWN);
  }
  for (int l = 0; l < m->n_layers; ++l) {
    const opus_llama_layer& L = m->layers[l];
    bf16* kc = static_cast<bf16*>(kv->k) + (size_t)l * layer_stride;
    bf16* vc = static_cast<bf16*>(kv->v) + (size_t)l * layer_stride;
    int sp = 1;
    pf_o.w = L.wo; pf_gu.w = L.wgu; pf_down.w = L.wdown;
    if (swap && l + 1 < m->n_layers) {
      pf_next.w = m->layers[l + 1].wqkv; pf_next.rows = qkv_n; pf_next.K = d;
      pf_next.split_k = splitk_for(B, qkv_n, d, ws->partial_bytes); pf_next.depth = pf_depth_for(PF_QKV);
    } else if (swap) {
      pf_next.w = m->lm_head; pf_next.rows = m->vocab; pf_next.K = d; pf_next.split_k = 1; pf_next.depth = pf_depth_for(PF_LM);
    }
    OPUS_TRY(linear_splitk(xn, B, L.wqkv, qkv_n, d, ws->partial, ws->partial_bytes, &sp, st, &pf_o));
    if (ctx().tun.decode_rope_fused) {
      // split-K reduce + RoPE + KV append happen inside the attention CTAs (one launch less per layer)
      OPUS_TRY(attn_decode_paged_fused(qkv, qkv_n, ws->partial, sp, s->pos, s->slot,
                                       static_cast<const bf16*>(m->rope_cos), static_cast<const bf16*>(m->rope_sin), kc,
                                       vc, s->block_table, s->max_blocks, s->ctx_len, attn, Hq * hd, B, Hq, Hkv, hd,
                                       kv->block_size, scale, st, L.bqkv));
    } else {
      OPUS_TRY(rope_llama_kvappend(qkv, ws->partial, sp, s->pos, s->slot, static_cast<const bf16*>(m->rope_cos),
                                   static_cast<const bf16*>(m->rope_sin), kc, vc, B, Hq, Hkv, hd, qkv_n,
                                   kv->block_size, st, L.bqkv));
      OPUS_TRY(attn_decode_paged(qkv, qkv_n, kc, vc, s->block_table, s->max_blocks, s->ctx_len, attn, Hq * hd, B, Hq,
                                 Hkv, hd, kv->block_size, scale, st));
    }
    OPUS_TRY(linear_splitk(attn, B, L.wo, d, Hq * hd, ws->partial, ws->partial_bytes, &sp, st, &pf_gu));
    OPUS_TRY(rmsnorm_bf16(nullptr, ws->partial, sp, h, h, static_cast<const bf16*>(L.ln2_w), xn, B, d, m->rms_eps, st));
    OPUS_TRY(linear(xn, B, L.wgu, 2 * ffn, d, EPI_SWIGLU, act, ffn, nullptr, nullptr, 0, nullptr, 0, st, &pf_down, true));
    OPUS_TRY(linear_splitk(act, B, L.wdown, d, ffn, ws->partial, ws->partial_bytes, &sp, st, &pf_next));
    const bf16* next_w = static_cast<const bf16*>(l + 1 < m->n_layers ? m->layers[l + 1].ln1_w : m->norm_w);
    OPUS_TRY(rmsnorm_bf16(nullptr, ws->partial, sp, h, h, next_w, xn, B, d, m->rms_eps, st));
  }
  Prefetch pf_first;  // the next decode step starts with layer 0's qkv weights
  if (swap) {
    pf_first.w = m->layers[0].wqkv; pf_first.rows = qkv_n; pf_first.K = d;
    pf_first.split_k = splitk_for(B, qkv_n, d, ws->partial_bytes); pf_first.depth = pf_depth_for(PF_QKV);
  }
  OPUS_TRY(linear(xn, B, m->lm_head, m->vocab, d, EPI_BF16, ws->logits, m->vocab, nullptr, nullptr, 0, nullptr, 0, st,
                  &pf_first, true));
  OPUS_TRY(llama_select(m, ws, s, B, st));
  return OPUS_OK;
}

// ------------------------------------------------------------------------------------------------ decode loop + graph
int set_tunable(const char* name, int value) {
  static const char* pf_names[5] = {"pf_qkv", "pf_o", "pf_gu", "pf_down", "pf_lm"};
  if (name == nullptr) return fail(OPUS_ERR_ARG, "set_tunable: null name");
  Tunables& t = ctx().tun;
  bool known = true;
  if (std::strcmp(name, "chain_l2_depth") == 0) gemm_set_chain_l2_depth(value);
  else if (std::strcmp(name, "decode_rope_fused") == 0) t.decode_rope_fused = value != 0;
  else if (std::strcmp(name, "decode_fused") == 0) t.decode_fused = value != 0;
  else if (std::strcmp(name, "gemm_2cta") == 0) gemm_set_2cta(value);
  else if (std::strcmp(name, "gemm_2cta_tr") == 0) gemm_set_2cta_tr(value);
  else if (std::strcmp(name, "tma_store") == 0) gemm_set_tma_store(value);
  else if (std::strcmp(name, "streamk_plain") == 0) gemm_set_streamk_plain(value);
  else if (std::strcmp(name, "streamk_fill") == 0) gemm_set_streamk_fill(value);
  else if (std::strcmp(name, "attn_mode") == 0) t.attn_mode = value < 0 ? 0 : (value > 2 ? 2 : value);
  else if (std::strcmp(name, "attn_tail") == 0) t.attn_tail = value != 0;
  else if (std::strcmp(name, "epi_warm") == 0) t.epi_warm = value != 0;
  else if (std::strcmp(name, "attn_split") == 0) t.attn_split = value;
  else if (std::strcmp(name, "pair_streamk") == 0) t.pair_streamk = value != 0;
  else if (std::strcmp(name, "group_n") == 0) t.group_n = value < 0 ? 0 : value;
  else if (std::strcmp(name, "group_n_hints") == 0) t.group_n_hints = value != 0;
  else if (std::strcmp(name, "group_m") == 0) t.group_m = value < 0 ? 0 : value;
  else if (std::strcmp(name, "l2_ahead") == 0) t.l2_ahead = value < 0 ? 0 : value;
  else if (std::strcmp(name, "wide_overhead") == 0) t.wide_overhead = value < 0 ? 0 : value;
  else if (std::strcmp(name, "decode_norm_fused") == 0) t.decode_norm_fused = value != 0;
  else {
    known = false;
    for (int i = 0; i < 5; ++i)
      if (std::strcmp(name, pf_names[i]) == 0) { t.pf_depth[i] = value < 0 ? 0 : value; known = true; }
  }
  if (!known) return fail(OPUS_ERR_ARG, "set_tunable: unknown name");
  return release_graphs();   // captured decode graphs bake the old value in
}

int release_graphs() {
  Context& c = ctx();
  std::lock_guard<std::mutex> lk(c.mu);
  for (auto& kv : c.graphs)
    if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
  c.graphs.clear();
  return OPUS_OK;
}

namespace {
// Everything a captured decode step bakes into its kernel arguments: the struct contents (pointers AND scalars such as
// pad_id, temperature, top_p, the immediate seed, max_blocks, n_eos, the stop-sequence fields) plus the batch size and the
// per-layer weight table. Two calls that differ in any of them get different graphs; the values that legitimately change
// between replays (positions, step counter, and the seed when state->seed_ptr is used) live in device memory.
template <typename T>
void key_add(std::string& k, const T& v) { k.append(reinterpret_cast<const char*>(&v), sizeof(T)); }
std::string graph_key(const opus_llama_model* m, const opus_kv_cache* kv, const opus_llama_workspace* ws,
                      const opus_decode_state* s, int B) {
  std::string k;
  k.reserve(sizeof(*m) + sizeof(*kv) + sizeof(*ws) + sizeof(*s) + 8);
  key_add(k, *m); key_add(k, *kv); key_add(k, *ws); key_add(k, *s); key_add(k, B);
  return k;
}
}  // namespace

int llama_decode_loop(const opus_llama_model* m, const opus_kv_cache* kv, const opus_llama_workspace* ws,
                      const opus_decode_state* s, int B, int n_steps, int check_every, int use_graph,
                      cudaStream_t st) {
  if (!m || !kv || !ws || !s) return fail(OPUS_ERR_ARG, "llama_decode_loop: null argument");
  if (n_steps <= 0) return 0;
  attn_decode_warmup();   // context scratch of the split-KV attention: never allocated first inside the capture below
  GraphEntry entry;
  if (use_graph) {
    Context& c = ctx();
    std::lock_guard<std::mutex> lk(c.mu);
    const std::string key = graph_key(m, kv, ws, s, B);
    auto it = c.graphs.find(key);
    if (it == c.graphs.end()) {
      cudaGraph_t graph = nullptr;
      const long long before = launch_count(false);
      // The caller's stream may be the legacy default stream, which cannot be captured: record the step on a private
      // non-blocking stream (nothing executes during capture) and replay the instantiated graph on the caller's stream.
      if (c.cap_stream == nullptr &&
          cudaStreamCreateWithFlags(&c.cap_stream, cudaStreamNonBlocking) != cudaSuccess)
        return fail(OPUS_ERR_CUDA, "decode_loop: create capture stream");
      if (cudaStreamBeginCapture(c.cap_stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess)
        return fail(OPUS_ERR_CUDA, "decode_loop: begin capture");
      const int rc = llama_decode_step(m, kv, ws, s, B, c.cap_stream);
      const cudaError_t ce = cudaStreamEndCapture(c.cap_stream, &graph);
      if (rc != OPUS_OK) {
        if (graph) cudaGraphDestroy(graph);
        return rc;
      }
      if (ce != cudaSuccess || graph == nullptr) return fail(OPUS_ERR_CUDA, "decode_loop: end capture");
      GraphEntry e;
      e.launches = launch_count(false) - before;
      note_launch(-e.launches);  // capture did not execute anything
      if (cudaGraphInstantiate(&e.exec, graph, 0) != cudaSuccess) {
        cudaGraphDestroy(graph);
        return fail(OPUS_ERR_CUDA, "decode_loop: instantiate");
      }
      cudaGraphDestroy(graph);
      if (c.graphs.size() >= 16) {   // bounded cache: callers that churn states (new out_ids per call) do not leak graphs
        for (auto& g : c.graphs)
          if (g.second.exec) cudaGraphExecDestroy(g.second.exec);
        c.graphs.clear();
      }
      it = c.graphs.emplace(key, e).first;
    }
    entry = it->second;
  }
  int done = 0;
  for (int i = 0; i < n_steps; ++i) {
    if (use_graph) {
      if (cudaGraphLaunch(entry.exec, st) != cudaSuccess) return fail(OPUS_ERR_CUDA, "decode_loop: graph launch");
      note_launch(entry.launches);
    } else {
      const int rc = llama_decode_step(m, kv, ws, s, B, st);
      if (rc != OPUS_OK) return rc;
    }
    ++done;
    if (check_every > 0 && (done % check_every) == 0 && i + 1 < n_steps) {
      int left = 1;
      if (cudaMemcpyAsync(&left, s->n_unfinished, sizeof(int), cudaMemcpyDeviceToHost, st) != cudaSuccess ||
          cudaStreamSynchronize(st) != cudaSuccess)
        return fail(OPUS_ERR_CUDA, "decode_loop: read n_unfinished");
      if (left <= 0) break;
    }
  }
  return done;
}

}  // namespace opus
