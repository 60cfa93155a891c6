// Context lifetime and the one place that reads OPUS_* environment variables (see context.h).
#include "context.h"

#include <cstdlib>

namespace opus {

namespace {
int env_int(const char* name, int dflt) {
  const char* e = std::getenv(name);
  return (e != nullptr && e[0] != 0) ? atoi(e) : dflt;
}
bool env_off(const char* name) {
  const char* e = std::getenv(name);
  return e != nullptr && e[0] == '0';
}
thread_local Context* t_current = nullptr;
}  // namespace

Tunables Tunables::from_env() {
  Tunables t{};
  static const char* pf_names[5] = {"OPUS_PF_QKV", "OPUS_PF_O", "OPUS_PF_GU", "OPUS_PF_DOWN", "OPUS_PF_LM"};
  const bool pf_on = !env_off("OPUS_PF");
  for (int i = 0; i < 5; ++i) t.pf_depth[i] = pf_on ? env_int(pf_names[i], 0) : 0;  // measured: no gain at batch 64
  t.decode_rope_fused = 1;
  t.decode_fused = env_int("OPUS_DECODE_FUSED", 0) == 1;
  t.chain_l2_depth = 0;
  {
    const char* e = std::getenv("OPUS_GEMM_2CTA");
    t.gemm_2cta = (e != nullptr && e[0] >= '0' && e[0] <= '2') ? e[0] - '0' : 2;
  }
  t.gemm_2cta_tr = env_int("OPUS_GEMM_2CTA_TR", 2);   // 2: gate/up (SwiGLU) too; measured 6.70 -> 6.17 ms per step at batch 256
  t.tma_store = env_off("OPUS_TMA_STORE") ? 0 : 1;
  t.streamk = env_off("OPUS_STREAMK") ? 0 : 1;
  t.streamk_plain = 0;
  t.streamk_fill = 90;
  t.group_m = env_int("OPUS_GEMM_GROUP_M", 0);
  t.plain_hints = env_int("OPUS_GEMM_HINTS", 0);
  t.group_n = env_int("OPUS_GEMM_GROUP_N", 0);
  t.group_n_hints = env_int("OPUS_GEMM_GROUP_N_HINTS", 1);
  {
    const char* e = std::getenv("OPUS_ATTN");
    t.attn_mode = (e == nullptr) ? 0 : (e[0] == 't' ? 2 : 1);
  }
  t.attn_tail = env_off("OPUS_ATTN_TAIL") ? 0 : 1;
  {
    const char* e = std::getenv("OPUS_PDL");
    t.pdl = (e == nullptr) ? 1 : (e[0] - '0');
  }
  // measured at batch 64 / ctx 512: 2 parts + last-arriver merge 4.70 ms per step against 4.25 unsplit (the merge is one more
  // dependent round trip through a saturated memory system), so the split is opt-in
  t.attn_split = env_int("OPUS_ATTN_SPLIT", 0);
  t.l2_ahead = env_int("OPUS_L2_AHEAD", 0);
  t.pair_streamk = env_int("OPUS_PAIR_STREAMK", 0) == 1;
  t.epi_warm = env_int("OPUS_EPI_WARM", 0) == 1;   // measured: no gain (tools/hop_probe.py WARM_AB=1), off by default
  t.wide_overhead = env_int("OPUS_WIDE_OVERHEAD", 8);
  t.decode_norm_fused = env_int("OPUS_DECODE_NORM_FUSED", 0) == 1;
  return t;
}

Context::~Context() {
  for (auto& kv : graphs)
    if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
  if (cap_stream) cudaStreamDestroy(cap_stream);
  if (sk.ws) cudaFree(sk.ws);
  if (sk.cnt) cudaFree(sk.cnt);
  if (chain_trace) cudaFree(chain_trace);
  if (attn_ws) cudaFree(attn_ws);
  if (attn_cnt) cudaFree(attn_cnt);
}

Context& ctx() {
  if (t_current != nullptr) return *t_current;
  static Context* process_default = new Context();   // lives for the process (no static-destruction-order CUDA calls)
  return *process_default;
}
Context* ctx_create() { return new Context(); }
void ctx_destroy(Context* c) {
  if (c == nullptr) return;
  if (t_current == c) t_current = nullptr;
  delete c;
}
void ctx_set_current(Context* c) { t_current = c; }

}  // namespace opus
