// Library context: every piece of mutable library state (run-time tunables, cached decode graphs, the stream-K fix-up
// workspace) lives in a Context instead of in process globals, so independent callers do not share knobs or graphs.
// Environment variables (OPUS_*) are read exactly once, when a context is created; the launch paths never call getenv.
// A thread works against its *current* context (opus_ctx_set_current, thread-local); threads that never set one share
// the process-default context, created on first use. See include/opus_b200.h, "Contexts".
#pragma once
#include <cuda_runtime.h>

#include <map>
#include <mutex>
#include <string>

namespace opus {

struct Tunables {
  int pf_depth[5];        // OPUS_PF_{QKV,O,GU,DOWN,LM}: k-blocks of the next weight matrix a decode GEMM prefetches into L2
  int decode_rope_fused;  // split-K reduce + RoPE + KV append inside the decode attention kernel
  int decode_fused;       // OPUS_DECODE_FUSED: persistent chain kernel for the decode GEMMs
  int chain_l2_depth;     // k-blocks the chain kernel prefetches into L2 per phase
  int gemm_2cta;          // OPUS_GEMM_2CTA: 0 off, 1 every eligible plain GEMM, 2 all but the SwiGLU epilogue
  int gemm_2cta_tr;       // OPUS_GEMM_2CTA_TR: CTA-pair form for swap-AB launches at batch 129..256 (0 off, 1 all but SwiGLU, 2 all)
  int tma_store;          // OPUS_TMA_STORE: plain bf16 / GELU epilogues through shared memory + TMA stores
  int streamk;            // OPUS_STREAMK: 0 disables the stream-K tail
  int streamk_plain;      // stream-K tail in the plain (non swap-AB) form
  int streamk_fill;       // largest partial-wave fill (percent) that takes the stream-K tail
  int group_m;            // OPUS_GEMM_GROUP_M: raster group override (0 = automatic)
  int group_n;            // OPUS_GEMM_GROUP_N: weight-panel band of the grouped-N raster (CTA-pair kernel, K > 8192; 0 = off)
  int group_n_hints;      // OPUS_GEMM_GROUP_N_HINTS: evict-first activations / evict-last weights under that raster
  int plain_hints;        // OPUS_GEMM_HINTS: L2 eviction hints of the plain form (0 = none)
  int attn_mode;          // OPUS_ATTN: 0 automatic, 1 mma.sync kernel, 2 tcgen05 kernel
  int attn_tail;          // OPUS_ATTN_TAIL: short query tails leave the tcgen05 kernel
  int pdl;                // OPUS_PDL: 0 off, 1 decode-sized launches, 2 every launch
  int l2_ahead;           // OPUS_L2_AHEAD: k-blocks a swap-AB GEMM requests into L2 ahead of its shared-memory ring
  int pair_streamk;       // OPUS_PAIR_STREAMK: stream-K tail in the CTA-pair swap-AB kernel (default off: measured slower)
  int attn_split;         // OPUS_ATTN_SPLIT: split-KV parts of the decode attention (0 / 1 off = default, -1 automatic, 2, 4)
  int epi_warm;           // OPUS_EPI_WARM: swap-AB GEMMs run their epilogue once "dry" to warm the instruction cache
  int wide_overhead;      // OPUS_WIDE_OVERHEAD: per-item cost (in k-blocks) of the split-K choice at batch 257..512
  int decode_norm_fused;  // decode: RMSNorm folded into the GEMMs (in-kernel split-K reduce + norm-on-load), batch <= 64
  static Tunables from_env();
};

struct GraphEntry {
  cudaGraphExec_t exec = nullptr;
  long long launches = 0;
};

struct SkWorkspace {
  float* ws = nullptr;
  int* cnt = nullptr;
  int state = 0;  // 0 = not tried, 1 = ready, -1 = unavailable
};

struct Context {
  Tunables tun = Tunables::from_env();
  std::mutex mu;                             // guards graphs / cap_stream (held across a graph capture)
  std::mutex sk_mu;                          // guards sk: taken by GEMM launches, also from inside a capture
  std::map<std::string, GraphEntry> graphs;  // decode graphs keyed by everything a captured step bakes in
  cudaStream_t cap_stream = nullptr;
  SkWorkspace sk;
  unsigned long long* chain_trace = nullptr;
  float* attn_ws = nullptr;                  // split-KV decode attention: partial results (guarded by sk_mu)
  int* attn_cnt = nullptr;                   //   and per-unit arrival counters
  ~Context();
};

Context& ctx();                 // the calling thread's current context
Context* ctx_create();
void ctx_destroy(Context* c);   // must not be current on any thread
void ctx_set_current(Context* c);  // nullptr = back to the process default

}  // namespace opus
