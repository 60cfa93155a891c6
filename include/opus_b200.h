/*
 * opus_b200 — C ABI of the B200-native OPUS-PLLM generation hot path (libopus_b200.so).
 *
 * The reference (Fanchuana/OPUS-PLLM) has no FFI layer: its hot path is Python calling torch / fair-esm / transformers.
 * The entry points below are what a maintainer binds (ctypes) to replace those calls; each one
 * cites the reference call site it stands in for (paths relative to multi_modality_model/ in the reference repo).
 *
 * Conventions
 *   - every function returns 0 (OPUS_OK) or a negative OPUS_ERR_* code; opus_last_error() gives a message;
 *   - pointers are raw DEVICE pointers unless a parameter is documented as host memory; the caller owns all buffers
 *     (no hidden allocation except the per-context stream-K scratch described under "Contexts"), `stream` is a
 *     cudaStream_t passed as void*; no entry point synchronises the stream except opus_llama_decode_loop when early EOS
 *     stopping is requested (check_every > 0, documented there) and the two profiling aids;
 *   - bf16 tensors are passed as void*; matrices are row-major with explicit leading dimensions in ELEMENTS;
 *   - weights of every nn.Linear keep the torch layout [out_features, in_features] (K contiguous).
 *   - There is NO CPU fallback: on a machine without an sm_100a GPU every compute entry point fails.
 */
#ifndef OPUS_B200_H_
#define OPUS_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OPUS_B200_ABI_VERSION 4

enum {
  OPUS_OK = 0,
  OPUS_ERR_ARG = -1,    /* invalid argument (shape / alignment / null pointer) */
  OPUS_ERR_CUDA = -2,   /* CUDA runtime error at launch */
  OPUS_ERR_DRIVER = -3, /* driver entry point (cuTensorMapEncodeTiled) unavailable */
  OPUS_ERR_TMAP = -4,   /* tensor-map encoding rejected */
  OPUS_ERR_STATE = -5   /* object used in the wrong state */
};

/* GEMM epilogues (opus_gemm_bf16) */
enum {
  OPUS_EPI_BF16 = 0,        /* out bf16 = acc (+ bias)                                  */
  OPUS_EPI_BF16_GELU = 1,   /* out bf16 = gelu_erf(acc + bias)        (ESM fc1, switch projector layer 0) */
  OPUS_EPI_RES_F32 = 2,     /* out f32  = residual_f32 + acc (+ bias) (ESM out_proj / fc2, fp32 residual stream) */
  OPUS_EPI_RES_BF16 = 3,    /* out bf16 = bf16(residual + bf16(acc))  (Llama o_proj / down_proj) */
  OPUS_EPI_SWIGLU = 4,      /* features interleaved (gate_j, up_j) -> out bf16[.., j] = silu(gate)*up */
  OPUS_EPI_PARTIAL_F32 = 5, /* split-K partial sums, out f32 [split][rows][ldo] */
  OPUS_EPI_F32 = 6,         /* out f32 = acc (+ bias), transposed form only */
  OPUS_EPI_BF16_RELU = 7    /* out bf16 = relu(acc + bias)            (OPT fc1: HF OPTDecoderLayer, activation "relu") */
};

int opus_abi_version(void);
/* Message of the last failing call made BY THE CALLING THREAD (thread-local; valid until that thread's next failure). */
const char* opus_last_error(void);

/* ------------------------------------------------------------------------------------------------------------------
 * Contexts
 *
 * All mutable library state lives in a context: the run-time tunables (opus_set_tunable), the cache of captured decode
 * graphs (opus_llama_decode_loop), and two lazily allocated device scratch areas (the stream-K fix-up workspace of the
 * GEMM, 19.4 MB, and the chain-trace buffer). Entry points act on the calling thread's CURRENT context; threads that
 * never call opus_ctx_set_current share the process-default context. A context serialises its work on one stream at a
 * time; use one context per concurrently used stream / host thread. The OPUS_* environment variables (OPUS_PDL,
 * OPUS_ATTN, OPUS_ATTN_TAIL, OPUS_GEMM_2CTA, OPUS_GEMM_2CTA_TR, OPUS_GEMM_GROUP_M, OPUS_GEMM_HINTS, OPUS_TMA_STORE,
 * OPUS_STREAMK, OPUS_DECODE_FUSED, OPUS_PF, OPUS_PF_*, and the opt-in experiments OPUS_GEMM_GROUP_N,
 * OPUS_GEMM_GROUP_N_HINTS, OPUS_ATTN_SPLIT, OPUS_L2_AHEAD, OPUS_PAIR_STREAMK, OPUS_EPI_WARM, OPUS_DECODE_NORM_FUSED,
 * OPUS_WIDE_OVERHEAD) are read once, when a context is created: launch paths never consult the environment.
 * ------------------------------------------------------------------------------------------------------------------ */
typedef struct opus_ctx opus_ctx;
int opus_ctx_create(opus_ctx** out);
/* Destroys the graphs and scratch of `ctx` (the caller makes sure no work of this context is still in flight). */
int opus_ctx_destroy(opus_ctx* ctx);
/* Makes `ctx` the calling thread's current context; NULL returns the thread to the process-default context. */
int opus_ctx_set_current(opus_ctx* ctx);
/* 0 when the current device is an sm_100 part this library was built for, OPUS_ERR_CUDA otherwise. */
int opus_device_check(void);

/* ------------------------------------------------------------------------------------------------------------------
 * Contractions
 * ------------------------------------------------------------------------------------------------------------------ */

/* D[M,N] = epilogue(A[M,K] * B[N,K]^T), bf16 operands, fp32 accumulation on tcgen05 tensor cores (TMA-fed, TMEM acc).
 *   transposed = 0: A = activations, B = weight [N=out_features, K];  out[m*ldo + n], bias[n], residual[m*ldr + n]
 *   transposed = 1: A = weight [M=out_features, K], B = activations [N=batch rows, K] (swap-AB, weight streaming for
 *                   batch <= 256);                                     out[n*ldo + m], bias[m], residual[n*ldr + m]
 *   split_k > 1 needs OPUS_EPI_PARTIAL_F32; reduce with opus_splitk_reduce_bf16 / opus_rmsnorm_bf16.
 * Stands in for every torch.nn.Linear on the path (cuBLAS in the reference): fair-esm q/k/v/out/fc1/fc2 reached from
 * cstp_v3/modelling.py:48, CSTPBase.protein_forward cstp_v3/modelling.py:396-400, the switch projector
 * multi_modality_v1/model/protein_mlp/builder.py:21-24, HF Llama linears reached from
 * multi_modality_v1/model/language_model/opus_llama.py:82-93.  Requires K % 8 == 0, lda/ldb % 8 == 0, 16-byte aligned
 * A and B, and N % 8 == 0 when transposed = 0. */
int opus_gemm_bf16(const void* A, int lda, const void* B, int ldb, int M, int N, int K, int transposed, int epilogue,
                   void* out, int ldo, const float* bias, const void* residual, int ldr, int split_k, int block_n,
                   void* stream);
/* Swap-AB GEMM (transposed = 1 semantics of opus_gemm_bf16: A = weight [M = out_features, K], B = activations [N = batch
 * rows <= 256, K]) with the fusions of the norm-fused decode step (HF LlamaDecoderLayer reached from opus_llama.py:127-132):
 *   splitk_fixup != 0 : split_k > 1 is reduced INSIDE the kernel (the CTA holding a tile's last k-split adds the other
 *                       splits' accumulators in split order), so any epilogue works with split_k > 1; needs
 *                       ceil(M/128) * split_k <= SM count. Replaces the separate reduce (opus_rmsnorm_bf16 with partials).
 *   sumsq_out != NULL : OPUS_EPI_RES_BF16 only. sumsq_out[slab * sumsq_ld + n] = sum over the 32 features of slab
 *                       (= feature / 32) of the squares of the bf16 values stored for batch row n (fp32; one writer each).
 *   norm_sumsq != NULL: batch <= 64, K % 64 == 0. The activation operand is the raw residual stream h; every k-slice is
 *                       rewritten in shared memory as bf16(norm_gamma[k] * bf16(h[n,k] * rstd[n])) before the tensor core
 *                       reads it, rstd[n] = rsqrt(sum_slab norm_sumsq[slab * norm_ld + n] / K + norm_eps): HF
 *                       LlamaRMSNorm with its rounding points, without a norm kernel or a normalised copy in HBM. */
int opus_gemm_bf16_fused(const void* A, int lda, const void* B, int ldb, int M, int N, int K, int epilogue, void* out,
                         int ldo, const float* bias, const void* residual, int ldr, int split_k, int splitk_fixup,
                         float* sumsq_out, int sumsq_ld, const float* norm_sumsq, int norm_slabs, int norm_ld,
                         const void* norm_gamma, float norm_eps, void* stream);
/* Suggested split-K factor for a (rows of A = M, rows of B = N, K) problem; 1 = none. */
int opus_gemm_suggest_split_k(int M, int N, int K, int transposed);
/* out bf16 [rows, cols] = sum_s partial[s][rows][cols] (+ bias[col]) (optionally erf-GELU). */
int opus_splitk_reduce_bf16(const float* partial, int n_partial, const float* bias, void* out, int rows, int cols,
                            int ldo, int gelu, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Bandwidth kernels
 * ------------------------------------------------------------------------------------------------------------------ */

/* x[i,:] = table[tok[i],:] * scale[i]  — ESM-2 token embedding with token-dropout rescale (fair-esm ESM2.forward,
 * called at cstp_v3/modelling.py:48).  tok: int32 [n_tok]; table fp32 [vocab, dim]; x fp32 [n_tok, dim]. */
int opus_esm_embed(const int32_t* tok, const float* scale, const float* table, float* x, int n_tok, int dim,
                   void* stream);
/* y bf16 = LayerNorm(x fp32 (+ delta)) * gamma + beta, row-wise (fair-esm pre-LN blocks). If delta (bf16, nullable) is
 * given, the pending branch output is first folded into the fp32 residual stream in place (x += delta); delta may
 * alias y. */
int opus_layernorm_f32_bf16(float* x, const void* delta, const float* gamma, const float* beta, void* y, int rows,
                            int cols, float eps, void* stream);
/* HF LlamaRMSNorm with optional fused residual add and split-K reduction:
 *   h = x  or  bf16(sum_s partial[s])      (exactly one of x / partial non-null)
 *   h = bf16(h + residual) if residual;  h_out <- h if h_out;  y <- w * bf16(h * rsqrt(mean(h^2)+eps)) if y. */
int opus_rmsnorm_bf16(const void* x, const float* partial, int n_partial, const void* residual, void* h_out,
                      const void* w, void* y, int rows, int cols, float eps, void* stream);
/* ESM rotary on fused [n_tok, ld] bf16 activations laid out q heads | k heads | v heads: q <- rope(q*q_scale),
 * k <- rope(k). cos/sin fp32 [max_pos, head_dim/2]; pos int32 [n_tok]. */
int opus_rope_esm_bf16(void* qkv, const int32_t* pos, const float* cos_t, const float* sin_t, int n_tok, int n_heads,
                       int head_dim, int ld, float q_scale, void* stream);
/* Llama rotary (bf16 rounding points of HF apply_rotary_pos_emb) fused with the paged KV-cache append.
 * qkv bf16 [n_tok, ld] = q heads | k heads | v heads; q,k rotated in place; k,v of token i written to cache slot
 * slot[i] (= block*block_size + offset; < 0 skips). Cache layout [blocks][n_kv_heads][block_size][head_dim] bf16.
 * If `partial` != NULL the row is first reduced from n_partial fp32 split-K partials [s][n_tok][ld].
 * Replaces HF DynamicCache.update (torch.cat per layer per step). cos/sin bf16 [max_pos, head_dim] built like HF's
 * LlamaRotaryEmbedding, emb = cat(freqs, freqs): only columns [0, head_dim/2) are read, the upper half is its copy. */
int opus_rope_llama_kvappend_bf16(void* qkv, const float* partial, int n_partial, const int32_t* pos,
                                  const int32_t* slot, const void* cos_t, const void* sin_t, void* kcache, void* vcache,
                                  int n_tok, int n_q_heads, int n_kv_heads, int head_dim, int ld, int block_size,
                                  void* stream);
/* Final LayerNorm of every residue + mean over residues [1, len-1) of each packed sequence + L2 normalise
 * (cstp_v3/modelling.py:53-55 and F.normalize at :398). pooled fp32 [n_seqs, dim]; pooled_l2 bf16 (nullable);
 * hidden_out fp32 [n_tok, dim] (nullable) receives the per-residue normalised states (representations[33]);
 * delta bf16 [n_tok, dim] (nullable) is a pending branch output added to x before the norm. */
int opus_final_ln_meanpool(const float* x, const void* delta, const int32_t* cu_seqlens, const float* gamma,
                           const float* beta, float* pooled, void* pooled_l2, float* hidden_out, int n_seqs, int dim,
                           float eps, void* stream);
int opus_l2norm_f32_bf16(const float* x, void* y, int rows, int dim, void* stream);
/* Soft-token splice gather (multi_modality_v1/model/opus_arch.py:176-270): out[i,:] = src[i] >= 0 ? embed[src[i]] :
 * src[i] == INT32_MIN ? 0 : soft[-src[i]-1]. */
int opus_splice_gather_bf16(const int32_t* src, const void* embed, const void* soft, void* out, int n_rows, int dim,
                            void* stream);
/* Greedy token selection with EOS bookkeeping (HF GenerationMixin._sample, do_sample=False). logits bf16 [n_rows, ld].
 * step: column of out_ids to write; if step_ptr != NULL the column is *step_ptr (device) instead. */
int opus_argmax_eos(const void* logits, int ld, int vocab, int n_rows, int32_t* finished, const int32_t* eos_ids,
                    int n_eos, int pad_id, int32_t* next_tok, int32_t* out_ids, int out_ld, int step,
                    int32_t* n_unfinished, void* stream);
/* Temperature + nucleus sampling with the same bookkeeping (HF GenerationMixin._sample, do_sample=True: scores / T, then
 * TopPLogitsWarper(top_p, min_tokens_to_keep=1), then one draw per row). Deterministic in (seed, row, step).
 * kept_count (nullable, [n_rows]) receives the size of each row's nucleus. */
int opus_sample_top_p(const void* logits, int ld, int vocab, int n_rows, float temperature, float top_p, uint64_t seed,
                      int32_t* finished, const int32_t* eos_ids, int n_eos, int pad_id, int32_t* next_tok,
                      int32_t* out_ids, int out_ld, int step, int32_t* n_unfinished, int32_t* kept_count,
                      void* stream);
/* Teacher-forced scoring (HF LlamaForCausalLM.forward(labels=...), reached from language_model/opus_llama.py:41-93):
 * loss[r] = logsumexp(fp32(logits[r,:])) - logits[r, target[r]]; 0 where target[r] < 0 (ignore_index). */
int opus_cross_entropy_bf16(const void* logits, int ld, int vocab, const int32_t* target, float* loss, int n_rows,
                            void* stream);
/* Stop-sequence check after a token selection step (column `step`, or *step_ptr when non-NULL, of out_ids has just been
 * written): unfinished rows whose emitted tail matches one of the sequences are marked finished (n_unfinished
 * decremented). Device-side counterpart of KeywordsStoppingCriteria (multi_modality_v1/mm_utils.py:43-75), per row. */
int opus_stop_sequences(const int32_t* out_ids, int out_ld, int n_rows, int step, const int32_t* step_ptr,
                        const int32_t* stop_seqs, const int32_t* stop_lens, int n_stop, int stop_ld, int32_t* finished,
                        int32_t* n_unfinished, void* stream);
int opus_embed_gather_bf16(const int32_t* tok, const void* table, void* x, int n_rows, int dim, void* stream);
/* OPT / Galactica family (multi_modality_v1/model/language_model/opus_opt.py:82-93,127-132 hand the work to HF
 * OPTForCausalLM; family chosen at model/builder.py:71-82). nn.LayerNorm over a bf16 residual stream with
 * the fusions of opus_rmsnorm_bf16 plus the bias of the linear whose split-K partials are reduced:
 *   h = x  or  bf16(sum_s partial[s] + red_bias)   (exactly one of x / partial non-null; red_bias fp32 [cols] nullable)
 *   h = bf16(h + residual) if residual;  h_out <- h if h_out;
 *   y <- bf16((h - mean) * rsqrt(var + eps) * gamma + beta) if y   (gamma, beta fp32 [cols]; beta nullable). */
int opus_layernorm_bf16(const void* x, const float* partial, int n_partial, const float* red_bias, const void* residual,
                        void* h_out, const float* gamma, const float* beta, void* y, int rows, int cols, float eps,
                        void* stream);
/* OPTLearnedPositionalEmbedding: h[i,:] = bf16(h[i,:] + table[pos[i] + offset, :]) in place (offset = 2 in HF);
 * table bf16 [table_rows, dim], pos int32 [n_rows]. */
int opus_add_pos_embed_bf16(void* h, const void* table, const int32_t* pos, int offset, int table_rows, int n_rows,
                            int dim, void* stream);
/* W += scale * (B @ A): peft merge_and_unload (multi_modality_v1/model/builder.py:107-109). */
int opus_lora_merge_bf16(void* W, const void* A, const void* B, int out_features, int in_features, int r, float scale,
                         void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Attention
 * ------------------------------------------------------------------------------------------------------------------ */

/* Flash-style attention over packed variable-length sequences: tokens of sequence b are rows
 * [cu_seqlens[b], cu_seqlens[b+1]). causal = 0: bidirectional (ESM encoder; padded batches pass their valid spans so
 * the key-padding mask is implicit); causal = 1: causal GQA (Llama prefill). head_dim in {64, 128}. n_tok = number of
 * packed rows of q/k/v (bounds of the TMA tensor maps). Batches whose longest sequence has >= 96 tokens run on the
 * tcgen05/TMEM kernel, shorter ones on the mma.sync kernel (tunable "attn_mode" / OPUS_ATTN=tc|mma forces either). */
int opus_attn_varlen_bf16(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, void* o, int ldo,
                          const int32_t* cu_seqlens, int n_seqs, int n_tok, int max_len, int n_q_heads, int n_kv_heads,
                          int head_dim, int causal, float scale, void* stream);
/* One query token per sequence against the paged KV cache (block_size 16, head_dim 128). ctx_len[b] counts the cached
 * tokens including the one appended this step. The cache must have been zero-initialised. */
int opus_attn_decode_paged_bf16(const void* q, int ldq, const void* kcache, const void* vcache,
                                const int32_t* block_table, int max_blocks, const int32_t* ctx_len, void* o, int ldo,
                                int n_seqs, int n_q_heads, int n_kv_heads, int head_dim, int block_size, float scale,
                                void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Composite forwards (host-side orchestration in C++, every kernel on the caller's stream)
 * ------------------------------------------------------------------------------------------------------------------ */

typedef struct {
  const float* ln1_g; const float* ln1_b;
  const void* wqkv;  const float* bqkv; /* bf16 [3*dim, dim] rows = q | k | v; fp32 [3*dim] */
  const void* wo;    const float* bo;   /* bf16 [dim, dim] */
  const float* ln2_g; const float* ln2_b;
  const void* w1;    const float* b1;   /* bf16 [ffn, dim] */
  const void* w2;    const float* b2;   /* bf16 [dim, ffn] */
} opus_esm2_layer;

typedef struct {
  int32_t n_layers, dim, n_heads, ffn_dim, vocab, rope_max_pos;
  float ln_eps;
  const float* embed;            /* fp32 [vocab, dim] */
  const opus_esm2_layer* layers; /* HOST array [n_layers] */
  const float* lnf_g; const float* lnf_b;
  const float* rope_cos; const float* rope_sin; /* fp32 [rope_max_pos, head_dim/2] */
} opus_esm2_model;

typedef struct {
  float* x;   /* fp32 [n_tok, dim] residual stream */
  void* xn;   /* bf16 [n_tok, dim] */
  void* qkv;  /* bf16 [n_tok, 3*dim] */
  void* attn; /* bf16 [n_tok, dim] */
  void* ffn;  /* bf16 [n_tok, ffn_dim] */
} opus_esm2_workspace;

/* ESM-2 forward + final LN + mean pool (+ L2 normalise) over packed sequences — replaces
 * ProteinSeqEmbeddingExtractor.get_protein_seq_embeddings (cstp_v3/modelling.py:37-57) after tokenisation.
 * tokens/tok_scale/pos: [n_tok]; cu_seqlens: [n_seqs+1]; pooled fp32 [n_seqs, dim]. */
int opus_esm2_forward(const opus_esm2_model* model, const opus_esm2_workspace* ws, const int32_t* tokens,
                      const float* tok_scale, const int32_t* pos, const int32_t* cu_seqlens, int n_seqs, int n_tok,
                      int max_len, float* pooled, void* pooled_l2, float* hidden_out, void* stream);

typedef struct {
  int32_t in_dim, cstp_dim, hidden_dim;     /* 1280, 5120, 8*H */
  const void* w_cstp; const float* b_cstp;  /* bf16 [cstp_dim, in_dim]; NULL w_cstp = identity projector */
  const void* w0;     const float* b0;      /* bf16 [hidden, cstp_dim] */
  const void* w2;     const float* b2;      /* bf16 [hidden, hidden]; NULL = 'linear' projector type */
} opus_projector_model;

/* CSTP projection + switch ("modality refinement") projector: x_l2 bf16 [n, in_dim] (already L2-normalised) ->
 * out bf16 [n, hidden_dim] (= [n, 8, H]). cstp_out [n, cstp_dim], h0 [n, hidden_dim] are bf16 scratch;
 * splitk_ws fp32 scratch of splitk_ws_bytes. Replaces opus_arch.py:115-131. */
int opus_projector_forward(const opus_projector_model* model, const void* x_l2, int n, void* cstp_out, void* h0,
                           void* out, float* splitk_ws, size_t splitk_ws_bytes, void* stream);

typedef struct {
  const void* ln1_w; /* bf16 [dim] */
  const void* wqkv;  /* bf16 [(Hq+2*Hkv)*hd, dim] rows = q | k | v */
  const void* wo;    /* bf16 [dim, Hq*hd] */
  const void* ln2_w;
  const void* wgu;   /* bf16 [2*ffn, dim], rows interleaved gate_0, up_0, gate_1, up_1, ... */
  const void* wdown; /* bf16 [dim, ffn] */
  const float* bqkv; /* fp32 [(Hq+2*Hkv)*hd] q|k|v projection bias (Qwen2 / OPT families) or NULL (Llama) */
  /* OPT / Galactica family only (arch == OPUS_ARCH_OPT; NULL otherwise). There ln1_w / ln2_w are unused: the LayerNorm
   * gains live in ln1_g / ln2_g as fp32, wgu holds fc1 [ffn, dim] (plain rows, no gate) and wdown holds fc2. */
  const float* ln1_g; const float* ln1_b; /* fp32 [dim] self_attn_layer_norm weight / bias (bias NULL = none) */
  const float* ln2_g; const float* ln2_b; /* fp32 [dim] final_layer_norm of the layer */
  const float* bo;   /* fp32 [dim]  out_proj bias or NULL */
  const float* b1;   /* fp32 [ffn]  fc1 bias or NULL */
  const float* b2;   /* fp32 [dim]  fc2 bias or NULL */
} opus_llama_layer;

typedef struct {
  int32_t n_layers, dim, n_q_heads, n_kv_heads, head_dim, ffn_dim, vocab, rope_max_pos;
  float rms_eps;
  const void* embed;              /* bf16 [vocab, dim] */
  const opus_llama_layer* layers; /* HOST array */
  const void* norm_w;             /* bf16 [dim] */
  const void* lm_head;            /* bf16 [vocab, dim] */
  const void* rope_cos; const void* rope_sin; /* bf16 [rope_max_pos, head_dim] (OPT: cos = 1, sin = 0, i.e. no rotation) */
  /* Decoder family. OPUS_ARCH_LLAMA (0): RMSNorm + RoPE + SwiGLU blocks (Llama, Qwen2; opus_llama.py / opus_qwen.py).
   * OPUS_ARCH_OPT (1): HF OPTDecoder with do_layer_norm_before (opus_opt.py; OPT and Galactica): learned positions
   * h = embeds + pos_embed[pos + 2], LayerNorm (eps = rms_eps), biased linears, fc1 -> ReLU | erf-GELU -> fc2, MHA. */
  int32_t arch;
  int32_t opt_act;                /* OPT family: 0 = ReLU, 1 = erf-GELU */
  const void* pos_embed;          /* OPT family: bf16 [pos_rows, dim] (HF embed_positions.weight, offset 2 included) */
  int32_t pos_rows;
  int32_t head_dim_real;          /* 0 or the model's head width when heads are stored zero-padded to head_dim columns
                                   * (the softmax scale is 1/sqrt(head_dim_real)); see llama.py / opt.py */
  const float* norm_g; const float* norm_b; /* OPT family: decoder.final_layer_norm fp32 [dim] */
} opus_llama_model;
enum { OPUS_ARCH_LLAMA = 0, OPUS_ARCH_OPT = 1 };

typedef struct {
  void* k; void* v;          /* bf16 [n_layers][num_blocks][n_kv_heads][block_size][head_dim], zero-initialised */
  int32_t num_blocks, block_size;
} opus_kv_cache;

typedef struct {
  void* h;      /* bf16 [rows, dim] residual stream */
  void* xn;     /* bf16 [rows, dim] */
  void* qkv;    /* bf16 [rows, (Hq+2Hkv)*hd] */
  void* attn;   /* bf16 [rows, Hq*hd] */
  void* act;    /* bf16 [rows, ffn] */
  float* partial; size_t partial_bytes; /* split-K scratch (decode) */
  void* last_h; /* bf16 [n_seqs, dim] */
  void* logits; /* bf16 [n_seqs, vocab] */
} opus_llama_workspace;

/* Per-batch decode state, all device int32. */
typedef struct {
  int32_t* next_tok;    /* [n_seqs] token to feed next step */
  int32_t* ctx_len;     /* [n_seqs] tokens in cache */
  int32_t* pos;         /* [n_seqs] scratch */
  int32_t* slot;        /* [n_seqs] scratch */
  int32_t* block_table; /* [n_seqs, max_blocks] */
  int32_t max_blocks;
  int32_t* finished;    /* [n_seqs] */
  int32_t* n_unfinished;/* [1] */
  int32_t* step;        /* [1] column of out_ids the next argmax writes */
  int32_t* out_ids;     /* [n_seqs, out_ld] */
  int32_t out_ld;
  const int32_t* eos_ids; int32_t n_eos; int32_t pad_id;
  /* token selection: do_sample == 0 -> argmax (greedy); else temperature / top-p sampling, u = hash(seed, row, step) */
  uint64_t seed; float temperature; float top_p; int32_t do_sample; int32_t reserved_;
  /* optional stop sequences (the "###" keyword of mm_utils.py:43-75 KeywordsStoppingCriteria, evaluated on the device):
   * a row also finishes when its last stop_lens[j] emitted tokens equal stop_seqs[j][0 .. stop_lens[j]) for some j.
   * stop_seqs int32 [n_stop, stop_ld] (NULL / n_stop == 0 = none). Checked by one extra small kernel per step. */
  const int32_t* stop_seqs; const int32_t* stop_lens; int32_t n_stop; int32_t stop_ld;
  /* optional: the sampling seed in DEVICE memory (uint64 [1]); when non-NULL it overrides `seed`, so the caller can
   * change the seed between replays of one captured decode graph (per call, per scheduling round) */
  const uint64_t* seed_ptr;
} opus_decode_state;

/* Prefill over packed prompt embeddings (already spliced): embeds bf16 [n_tok, dim] is copied into ws->h; K/V of every
 * token go to cache slot slot[i]; logits of the LAST token of each sequence land in ws->logits [n_seqs, vocab].
 * Replaces the first HF LlamaForCausalLM.forward of generate (opus_llama.py:127-132). Does not select tokens. */
int opus_llama_prefill(const opus_llama_model* model, const opus_kv_cache* cache, const opus_llama_workspace* ws,
                       const void* embeds, const int32_t* pos, const int32_t* slot, const int32_t* cu_seqlens,
                       const int32_t* last_rows, int n_seqs, int n_tok, int max_len, void* stream);
/* One greedy decode step for n_seqs sequences: advance state, embed next_tok, 32 layers against the paged cache,
 * lm_head, argmax + EOS bookkeeping (writes out_ids[:, *step]). Graph-capturable (no host-dependent values). */
int opus_llama_decode_step(const opus_llama_model* model, const opus_kv_cache* cache, const opus_llama_workspace* ws,
                           const opus_decode_state* state, int n_seqs, void* stream);
/* Select the first token from ws->logits (after prefill): argmax + EOS bookkeeping into column *state->step. */
int opus_llama_select(const opus_llama_model* model, const opus_llama_workspace* ws, const opus_decode_state* state,
                      int n_seqs, void* stream);
/* Greedy decode loop: n_steps calls of opus_llama_decode_step replayed from a CUDA graph captured on `stream`
 * (the graph is cached inside the current context, keyed by the contents of all four structs and n_seqs). If check_every > 0 the loop reads
 * *n_unfinished every check_every steps (this synchronises the stream) and stops early when it reaches 0.
 * Returns the number of steps executed (>= 0) or a negative error. */
int opus_llama_decode_loop(const opus_llama_model* model, const opus_kv_cache* cache, const opus_llama_workspace* ws,
                           const opus_decode_state* state, int n_seqs, int n_steps, int check_every, int use_graph,
                           void* stream);
/* Drop the current context's cached CUDA graphs (call before freeing buffers they reference). A graph is keyed by
 * everything a captured step bakes in (model / cache / workspace / state structs by value, batch size), so changing a
 * scalar such as pad_id or top_p can never replay a stale graph; at most 16 graphs are kept per context. */
int opus_release_graphs(void);
/* Run-time tunables of the composite forwards (measurement aid; defaults are the tuned values). Names: "pf_qkv",
 * "pf_o", "pf_gu", "pf_down", "pf_lm" = k-blocks (64 K-elements each) per work item that a decode GEMM prefetches into
 * L2 for the NEXT weight matrix of the chain (0 = off); "streamk_fill" = largest partial-wave fill (percent of the SMs)
 * for which a swap-AB GEMM cuts its last wave along K (0 = off); "streamk_plain" = 1 enables the same for the plain
 * form (off by default: it would make a token's rounding depend on its position in the batch); "decode_fused" = 1
 * runs o_proj -> norm -> gate/up -> down -> norm -> next qkv / lm_head of a decode step as one persistent chain kernel,
 * 0 (default) = one kernel per GEMM / norm; "chain_l2_depth" = k-blocks the chain kernel prefetches into L2 per phase;
 * "gemm_2cta" = 0 single-CTA GEMM everywhere, 1 CTA-pair (cta_group::2) form for every large plain GEMM, 2 (default) the
 * pair form except under the SwiGLU epilogue; "gemm_2cta_tr" = 0 keeps swap-AB launches at batch 129..512 on the
 * single-CTA kernel, 1 = CTA-pair kernel where its work items fill the pairs except under SwiGLU, 2 (default) = gate/up too
 * (batch 129..256; at batch 257..512 gate/up and lm_head run in the plain form with a stream-K tail);
 * "wide_overhead" = per-item cost, in k-blocks, of the split-K choice at batch 257..512 (default 8);
 * "decode_rope_fused" = 0 runs the decode step's split-K reduce + RoPE + KV append as its own kernel instead of inside
 * the paged-attention CTAs (default 1);
 * "tma_store" = 0 sends the plain bf16 / GELU GEMM epilogues back to direct row-per-thread stores (and the encoder's rotary
 * embedding back to its own kernel).
 * "attn_mode" = 0 automatic, 1 mma.sync attention, 2 tcgen05 attention; "attn_tail" = 0 keeps short query tails inside the
 * tcgen05 attention kernel.
 * Opt-in experiments, all measured slower or equal on B200 and off by default: "decode_norm_fused"
 * (RMSNorm folded into the decode GEMMs, batch <= 64), "attn_split" (split-KV decode attention: -1 automatic, 2, 4),
 * "pair_streamk" (stream-K tail of any size in the CTA-pair swap-AB kernel), "l2_ahead" (k-blocks requested into L2 ahead of
 * the shared-memory ring), "epi_warm" (dry epilogue pass), "group_m" / "group_n" / "group_n_hints" (raster overrides of the
 * plain GEMMs).
 * Unknown names fail with OPUS_ERR_ARG. Acts on the current context and drops its cached graphs. */
int opus_set_tunable(const char* name, int value);
/* Profiling aid: between opus_trace_begin(stream) and opus_trace_end the composite forwards record a CUDA event after
 * every kernel launch (not inside graph capture); opus_trace_end synchronises and writes "label<TAB>microseconds\n"
 * lines (time since the previous launch completed) into the HOST buffer buf, returning the bytes written. */
/* Profiling aid for the fused decode chain kernel: enable != 0 makes every following chain launch record globaltimer
 * stamps per (CTA, phase): [0] producer passed the phase's barrier, [1] first accumulator ready, [2] phase finished,
 * [3] norm phase started. out != NULL copies the last launch's [n_sms][6][4] stamps to the host (synchronises).
 * Returns the SM count or a negative error. */
int opus_chain_trace(int enable, unsigned long long* out, int cap_words);
int opus_trace_begin(void* stream);
int opus_trace_end(char* buf, int cap);
/* Number of kernel launches issued by this library since the last call (bench.py's gpu_launches counter). */
long long opus_launch_count(int reset);

#ifdef __cplusplus
}
#endif
#endif /* OPUS_B200_H_ */
