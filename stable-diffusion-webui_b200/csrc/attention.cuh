// Argument block of the tcgen05 flash-attention kernel (attention.cu).
#pragma once
#include "common.cuh"

namespace sdxe {

struct alignas(64) AttnArgs {
  // 4D per-head views (common.cuh make_tmap_heads): {d (true head dim: boxes reaching past it are zero-filled), token,
  // head, batch}, box 64 x 128 x 1 x 1, 128B swizzle
  CUtensorMap tmQ;
  CUtensorMap tmK;
  CUtensorMap tmV;
  int B, H, Nq, Nk;
  int dqk_slabs;    // dqk_pad / 64  (1..8)
  int dv_slabs;     // dv_pad / 64   (1..4)
  int dv;           // valid value columns per head that are stored (multiple of 8)
  int dqk;          // valid q/k columns (<= dqk_slabs * 64); columns beyond are zero padding
  int q_resident;   // set by the launcher
  int num_slots;    // set by the launcher
  float scale_log2; // softmax scale * log2(e)
  void* out;        // [B*Nq, ldo] 16-bit; head h writes columns out_col0 + h*dv ...
  int ldo;
  int out_col0;
  unsigned long long* trace;  // timeline of CTA (0,0), only in builds with -DSDXE_ATT_TRACE=1; null otherwise
};

int attention_launch(const AttnArgs& a, bool bf16, cudaStream_t stream);
int attention_init();
// two-query-tile variant (attention2.cu), used automatically by attention_launch when eligible
bool attention2_eligible(const AttnArgs& a);
int attention2_launch(const AttnArgs& a, bool bf16, cudaStream_t stream);
int attention2_init();
// persistent, software-pipelined short-KV (cross-) attention: Nk <= 128, head dim <= 64 (attention_x.cu)
bool attentionx_eligible(const AttnArgs& a);
int attentionx_launch(const AttnArgs& a, bool bf16, cudaStream_t stream);
int attentionx_init();

}  // namespace sdxe
