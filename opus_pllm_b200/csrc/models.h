// Composite forwards (models.cu).
#pragma once
#include <cuda_runtime.h>

#include "../../include/opus_b200.h"

namespace opus {
int esm2_forward(const opus_esm2_model* m, const opus_esm2_workspace* ws, const int* tokens, const float* tok_scale,
                 const int* pos, const int* cu_seqlens, int n_seqs, int n_tok, int max_len, float* pooled,
                 void* pooled_l2, float* hidden_out, cudaStream_t st);
int projector_forward(const opus_projector_model* m, const void* x_l2, int n, void* cstp_out, void* h0, void* out,
                      float* ws, size_t ws_bytes, cudaStream_t st);
int llama_prefill(const opus_llama_model* m, const opus_kv_cache* kv, const opus_llama_workspace* ws,
                  const void* embeds, const int* pos, const int* slot, const int* cu_seqlens, const int* last_rows,
                  int n_seqs, int n_tok, int max_len, cudaStream_t st);
int llama_select(const opus_llama_model* m, const opus_llama_workspace* ws, const opus_decode_state* s, int n_seqs,
                 cudaStream_t st);
int llama_decode_step(const opus_llama_model* m, const opus_kv_cache* kv, const opus_llama_workspace* ws,
                      const opus_decode_state* s, int B, cudaStream_t st);
int llama_decode_loop(const opus_llama_model* m, const opus_kv_cache* kv, const opus_llama_workspace* ws,
                      const opus_decode_state* s, int B, int n_steps, int check_every, int use_graph, cudaStream_t st);
int release_graphs();
int set_tunable(const char* name, int value);
int trace_begin(cudaStream_t st);
int trace_end(char* buf, int cap);
}  // namespace opus
