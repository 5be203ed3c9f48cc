// Two-query-tile flash attention for sm_100a, one softmax THREAD per query row (head dims up to 128, Nq >= 256).
// Same contract as attention.cu (AttnArgs) and the same CTA shape as attention2.cu (256 query rows, every K / V slab
// serves two S = Q K^T and two O += P V products; TMEM S_A | S_B | O_A | O_B), but the softmax is organised the FA-4 way:
//
//  * 8 softmax warps: warp (tile, TMEM lane quarter), a thread owns a whole 128-column score row. It copies the row
//    into registers first thing (4 x tcgen05.ld 32x32b.x32) and hands the TMEM buffer straight back (s_free), so the
//    MMA warp computes S(i+1) of that tile WHILE its softmax of block i runs: the score tile of the next block is
//    always waiting. (attention2 keeps S in TMEM through the block: ncu showed 20 % of its softmax-warp time waiting
//    for the S round trip and 10 % for the PV round trip.)
//  * the row max, the lazy-rescale decision (only when the max moves by more than 2^8) and the row sum are private to
//    the thread: no shared-memory exchange, no named barrier; the 128 exponentials of a row are independent
//    instruction streams (ILP instead of more warps).
//  * P is stored only after the row's exponentials are done and PV(i-1) has completed, so the single P buffer per tile
//    never stalls the exponentials; O is rescaled in TMEM (per-lane factor) only when some row of the warp needs it.
//
// Barriers, one phase per key block (parity i & 1):
//   s_full[T]   MMA -> softmax_T : S_T(i) complete in TMEM
//   s_free[T]   softmax_T -> MMA : all four warps hold S_T(i) in registers (S_T may be overwritten)
//   p_ready[T]  softmax_T -> MMA : P_T(i) in smem, O_T rescaled if needed
//   pv_done[T]  MMA -> softmax_T : O_T includes block i, P_T free
#include "attention.cuh"
#include <algorithm>

namespace sdxe {

static constexpr int SLAB4 = 16384;
static constexpr int ATT4_THREADS = 320;  // warp 0 TMA, warp 1 MMA, warps 2-9 softmax (tile x TMEM lane quarter)

template <bool BF16>
__global__ void __launch_bounds__(ATT4_THREADS, 1) attention4_kernel(const __grid_constant__ AttnArgs a) {
  using T = T16<BF16>;
  using TT = typename T::type;
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t pad = ((raw + 1023u) & ~1023u) - raw;
  uint8_t* smem = smem_raw + pad;
  const uint32_t sbase = raw + pad;

  const int NS = a.num_slots;
  const int QS = a.dqk_slabs, VS = a.dv_slabs;
  const uint32_t sQ = sbase;                          // [2 tiles][QS] slabs
  const uint32_t sRing = sQ + 2 * QS * SLAB4;
  const uint32_t sP = sRing + NS * SLAB4;             // [2 tiles][2 slabs]
  const uint32_t bar_base = sP + 4 * SLAB4;
  auto slot_full = [&](int s) { return bar_base + 8u * s; };
  auto slot_empty = [&](int s) { return bar_base + 8u * (NS + s); };
  const uint32_t q_full = bar_base + 8u * (2 * NS);
  auto s_full = [&](int t) { return bar_base + 8u * (2 * NS + 1 + t); };
  auto s_free = [&](int t) { return bar_base + 8u * (2 * NS + 3 + t); };
  auto p_ready = [&](int t) { return bar_base + 8u * (2 * NS + 5 + t); };
  auto pv_done = [&](int t) { return bar_base + 8u * (2 * NS + 7 + t); };
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + (bar_base - sbase) + 8 * (2 * NS + 9));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 256;
  const int bh = blockIdx.y;
  const int hb_b = bh / a.H, hb_h = bh - hb_b * a.H;  // (batch, head) coordinates of the 4D per-head tensor maps
  const int nblk = (a.Nk + 127) / 128;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NS; ++s) { mbar_init(slot_full(s), 1); mbar_init(slot_empty(s), 1); }
    mbar_init(q_full, 1);
    for (int t = 0; t < 2; ++t) {
      mbar_init(s_full(t), 1);
      mbar_init(s_free(t), 4);
      mbar_init(p_ready(t), 4);
      mbar_init(pv_done(t), 1);
    }
    fence_mbar_init();
    tma_prefetch_desc(&a.tmQ);
    tma_prefetch_desc(&a.tmK);
    tma_prefetch_desc(&a.tmV);
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_ptr_smem), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_wait();  // q / k / v of the producing GEMM are complete; the output buffer is free

  if (warp == 0) {
    // ---------------------------------------------------------------- producer (converged warp, elected issue)
    if (elect_one()) {
      mbar_expect_tx(q_full, 2 * QS * SLAB4);
      for (int t = 0; t < 2; ++t)
        for (int c = 0; c < QS; ++c) tma_load_4d(sQ + (t * QS + c) * SLAB4, &a.tmQ, q_full, c * 64, q0 + t * 128, hb_h, hb_b);
    }
    __syncwarp();
    int slot = 0;
    uint32_t phase = 0;
    auto push = [&](const CUtensorMap* tm, int c0, int r0) {
      mbar_wait(slot_empty(slot), phase ^ 1u);
      if (elect_one()) {
        mbar_expect_tx(slot_full(slot), SLAB4);
        tma_load_4d(sRing + slot * SLAB4, tm, slot_full(slot), c0, r0, hb_h, hb_b);
      }
      __syncwarp();
      if (++slot == NS) { slot = 0; phase ^= 1u; }
    };
    // ring order == consumption order: K_0, (K_1, V_0), (K_2, V_1), ..., V_{n-1}
    for (int i = 0; i <= nblk; ++i) {
      if (i < nblk)
        for (int c = 0; c < QS; ++c) push(&a.tmK, c * 64, i * 128);
      if (i >= 1)
        for (int vs = 0; vs < VS; ++vs) push(&a.tmV, vs * 64, (i - 1) * 128);
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer (converged warp, elected issue)
    const uint32_t idesc_s = umma_idesc(BF16 ? 1 : 0, 128, 128, 0, 0);
    const uint32_t idesc_pv = umma_idesc(BF16 ? 1 : 0, 128, 64, 0, 1);
    const int ksteps_last = (a.dqk - (QS - 1) * 64 + 15) / 16;
    const int n_last = (a.dv - (VS - 1) * 64 + 15) / 16 * 16;
    const uint32_t idesc_pv_last = umma_idesc(BF16 ? 1 : 0, 128, n_last, 0, 1);
    int slot = 0;
    uint32_t phase = 0;
    auto pop = [&](int& slot_id) -> uint32_t {  // wait for the next slab in ring order; caller releases it later
      mbar_wait(slot_full(slot), phase);
      slot_id = slot;
      const uint32_t addr = sRing + slot * SLAB4;
      if (++slot == NS) { slot = 0; phase ^= 1u; }
      return addr;
    };
    uint64_t qd[2][2], pd[2][2];
#pragma unroll
    for (int t = 0; t < 2; ++t) {
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        qd[t][c] = umma_desc_sw128(sQ + (t * QS + c) * SLAB4, 16, 1024);
        pd[t][c] = umma_desc_sw128(sP + (uint32_t)(t * 2 + c) * SLAB4, 16, 1024);
      }
    }
    uint64_t kd[2] = {0, 0}, vd[2] = {0, 0};
    int k_slot[2] = {0, 0}, v_slot[2] = {0, 0};
    auto issue_s = [&](int t) {
      const uint32_t d_s = tmem_base + (uint32_t)(t * 128);
      if (elect_one()) {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          if (c < QS) {
            const int ks = (c == QS - 1) ? ksteps_last : 4;
#pragma unroll
            for (int k = 0; k < 4; ++k)
              if (k < ks) tc_mma_f16(d_s, qd[t][c] + 2 * k, kd[c] + 2 * k, idesc_s, (c | k) != 0 ? 1u : 0u);
          }
        }
        tc_commit(s_full(t));
      }
      __syncwarp();
    };
    auto issue_pv = [&](int t, int j) {
      if (elect_one()) {
#pragma unroll
        for (int vs = 0; vs < 2; ++vs) {
          if (vs < VS) {
            const uint32_t d_o = tmem_base + 256u + (uint32_t)(t * 128 + vs * 64);
            const uint32_t id = (vs == VS - 1) ? idesc_pv_last : idesc_pv;
#pragma unroll
            for (int k = 0; k < 8; ++k)  // 16 key rows per step: +2048 B in V (= +128 in the addr>>4 field), +32 B in P
              tc_mma_f16(d_o, pd[t][k >> 2] + 2 * (k & 3), vd[vs] + 128 * k, id, (j | k) != 0 ? 1u : 0u);
          }
        }
        tc_commit(pv_done(t));
      }
      __syncwarp();
    };
    auto release = [&](const int* slots, int n) {  // free ring slabs once the MMAs issued so far have read them
      if (elect_one()) {
#pragma unroll
        for (int c = 0; c < 2; ++c)
          if (c < n) tc_commit(slot_empty(slots[c]));
      }
      __syncwarp();
    };
    mbar_wait(q_full, 0);
#pragma unroll
    for (int c = 0; c < 2; ++c)
      if (c < QS) kd[c] = umma_desc_sw128(pop(k_slot[c]), 16, 1024);
    tc_fence_after();
    issue_s(0);
    issue_s(1);
    release(k_slot, QS);
    for (int i = 0; i < nblk; ++i) {
      const bool more = i + 1 < nblk;
      if (more) {
#pragma unroll
        for (int c = 0; c < 2; ++c)
          if (c < QS) kd[c] = umma_desc_sw128(pop(k_slot[c]), 16, 1024);
      }
#pragma unroll
      for (int vs = 0; vs < 2; ++vs)
        if (vs < VS) vd[vs] = umma_desc_sw128(pop(v_slot[vs]), SLAB4, 1024);
      // four events per block, served in arrival order: S_t(i+1) as soon as tile t's rows are in registers,
      // O_t += P_t V as soon as P_t(i) is in shared memory
      bool s0 = !more, s1 = !more, p0 = false, p1 = false;
      const uint32_t par = (uint32_t)(i & 1);
      for (uint32_t spin = 0; !(s0 && s1 && p0 && p1); ++spin) {
        if (!s0 && __any_sync(0xffffffffu, mbar_test(s_free(0), par))) { tc_fence_after(); issue_s(0); s0 = true; }
        if (!s1 && __any_sync(0xffffffffu, mbar_test(s_free(1), par))) { tc_fence_after(); issue_s(1); s1 = true; }
        if (!p0 && __any_sync(0xffffffffu, mbar_test(p_ready(0), par))) { tc_fence_after(); issue_pv(0, i); p0 = true; }
        if (!p1 && __any_sync(0xffffffffu, mbar_test(p_ready(1), par))) { tc_fence_after(); issue_pv(1, i); p1 = true; }
        if (spin > (1u << 28)) {
          printf("sdxe: attention4 MMA warp watchdog block(%d,%d) i %d state %d%d%d%d\n", blockIdx.x, blockIdx.y, i, s0, s1, p0, p1);
          __trap();
        }
      }
      if (more) release(k_slot, QS);
      release(v_slot, VS);
    }
  } else {
    // ---------------------------------------------------------------- softmax / epilogue: one thread per query row
    const int sw = warp - 2;
    const int quarter = warp & 3;  // TMEM lane quarter (hardware: warp id mod 4)
    const int t = sw >> 2;         // tile
    const int row = quarter * 32 + lane;
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    const uint32_t t_s = tmem_base + (uint32_t)(t * 128) + lane_base;
    const uint32_t t_o = tmem_base + 256u + (uint32_t)(t * 128) + lane_base;
    const uint32_t p_row = sP + (uint32_t)(t * 2) * SLAB4 + (uint32_t)row * 128u;
    const int o_chunks = (a.dv + 31) / 32;
    const float sl2 = a.scale_log2;
    float m_run = -INFINITY, l_run = 0.f;
    for (int i = 0; i < nblk; ++i) {
      mbar_wait(s_full(t), (uint32_t)(i & 1));
      tc_fence_after();
      uint32_t r[128];
      tmem_ld32(t_s, r);
      tmem_ld32(t_s + 32, r + 32);
      tmem_ld32(t_s + 64, r + 64);
      tmem_ld32(t_s + 96, r + 96);
      tc_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_free(t));  // the whole score row lives in registers: S_t may be recomputed
      if ((i + 1) * 128 > a.Nk) {  // only the last block has invalid key columns
#pragma unroll
        for (int j = 0; j < 128; ++j)
          if (i * 128 + j >= a.Nk) r[j] = 0xff800000u;  // -inf
      }
      float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
      for (int j = 0; j < 128; j += 8) {
        mx0 = fmaxf(fmaxf(mx0, __uint_as_float(r[j + 0])), __uint_as_float(r[j + 1]));
        mx1 = fmaxf(fmaxf(mx1, __uint_as_float(r[j + 2])), __uint_as_float(r[j + 3]));
        mx2 = fmaxf(fmaxf(mx2, __uint_as_float(r[j + 4])), __uint_as_float(r[j + 5]));
        mx3 = fmaxf(fmaxf(mx3, __uint_as_float(r[j + 6])), __uint_as_float(r[j + 7]));
      }
      const float m_cand = fmaxf(m_run, fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)));
      // lazy rescale: keep the stale max while the block max stays within 2^8 of it (p <= 256: exact enough in 16 bit,
      // sums in fp32); first block: -inf -> always
      const bool need = (m_cand - m_run) * sl2 > 8.f;
      float alpha = 1.f;
      if (need) {
        alpha = ex2_approx((m_run - m_cand) * sl2);
        m_run = m_cand;
        l_run *= alpha;
      }
      const float mb = m_run * sl2;
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
      uint32_t pk[64];
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        const float p0 = ex2_approx(fmaf(__uint_as_float(r[q * 8 + 0]), sl2, -mb));
        const float p1 = ex2_approx(fmaf(__uint_as_float(r[q * 8 + 1]), sl2, -mb));
        const float p2 = ex2_approx(fmaf(__uint_as_float(r[q * 8 + 2]), sl2, -mb));
        const float p3 = ex2_approx(fmaf(__uint_as_float(r[q * 8 + 3]), sl2, -mb));
        const float p4 = ex2_approx(fmaf(__uint_as_float(r[q * 8 + 4]), sl2, -mb));
        const float p5 = ex2_approx(fmaf(__uint_as_float(r[q * 8 + 5]), sl2, -mb));
        const float p6 = ex2_approx(fmaf(__uint_as_float(r[q * 8 + 6]), sl2, -mb));
        const float p7 = ex2_approx(fmaf(__uint_as_float(r[q * 8 + 7]), sl2, -mb));
        s0 += p0 + p1; s1 += p2 + p3; s2 += p4 + p5; s3 += p6 + p7;
        pk[q * 4 + 0] = T::pack(p0, p1);
        pk[q * 4 + 1] = T::pack(p2, p3);
        pk[q * 4 + 2] = T::pack(p4, p5);
        pk[q * 4 + 3] = T::pack(p6, p7);
      }
      l_run += (s0 + s1) + (s2 + s3);
      if (i >= 1) {
        mbar_wait(pv_done(t), (uint32_t)((i - 1) & 1));  // O_t holds blocks < i, P_t is free
        tc_fence_after();
        if (__any_sync(0xffffffffu, need)) {  // some row of this warp moved its max: rescale the warp's O rows (factor 1 elsewhere)
          for (int c = 0; c < o_chunks; ++c) {
            uint32_t o[32];
            tmem_ld32(t_o + c * 32, o);
            tc_wait_ld();
#pragma unroll
            for (int j = 0; j < 32; ++j) o[j] = __float_as_uint(__uint_as_float(o[j]) * alpha);
            tmem_st32(t_o + c * 32, o);
          }
          tc_wait_st();
        }
      }
#pragma unroll
      for (int u = 0; u < 16; ++u) {  // 16-byte unit u of the row: slab u >> 3, unit (u & 7) ^ (row & 7) (128B swizzle)
        const uint32_t addr = p_row + (uint32_t)(u >> 3) * SLAB4 + (uint32_t)(((u & 7) ^ (row & 7)) * 16);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pk[u * 4 + 0]), "r"(pk[u * 4 + 1]),
                     "r"(pk[u * 4 + 2]), "r"(pk[u * 4 + 3])
                     : "memory");
      }
      tc_fence_before();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_ready(t));
    }
    // ---- epilogue
    mbar_wait(pv_done(t), (uint32_t)((nblk - 1) & 1));
    tc_fence_after();
    const int q = q0 + t * 128 + row;
    const float inv_l = 1.f / l_run;
    TT* orow = reinterpret_cast<TT*>(a.out) + ((size_t)hb_b * a.Nq + q) * a.ldo + a.out_col0 + hb_h * a.dv;
    for (int c = 0; c < o_chunks; ++c) {
      uint32_t o[32];
      tmem_ld32(t_o + c * 32, o);
      tc_wait_ld();
      if (q < a.Nq) {
#pragma unroll
        for (int g = 0; g < 32; g += 8) {
          if (c * 32 + g + 8 <= a.dv) {
            uint4 u;
            u.x = T::pack(__uint_as_float(o[g + 0]) * inv_l, __uint_as_float(o[g + 1]) * inv_l);
            u.y = T::pack(__uint_as_float(o[g + 2]) * inv_l, __uint_as_float(o[g + 3]) * inv_l);
            u.z = T::pack(__uint_as_float(o[g + 4]) * inv_l, __uint_as_float(o[g + 5]) * inv_l);
            u.w = T::pack(__uint_as_float(o[g + 6]) * inv_l, __uint_as_float(o[g + 7]) * inv_l);
            *reinterpret_cast<uint4*>(orow + c * 32 + g) = u;
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

int attention4_init() {
  static bool done = false;
  if (!done) {
    SDXE_CUDA_CHECK(cudaFuncSetAttribute(attention4_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    SDXE_CUDA_CHECK(cudaFuncSetAttribute(attention4_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    done = true;
  }
  return 0;
}

bool attention4_eligible(const AttnArgs& a) {
  return a.dqk_slabs <= 2 && a.dv_slabs <= 2 && a.Nq >= 256 && a.dv <= 128 && a.dv % 8 == 0;
}

int attention4_launch(const AttnArgs& a_in, bool bf16, cudaStream_t stream) {
  AttnArgs a = a_in;
  const int budget = (224 * 1024 - 2048) / SLAB4;  // 13 slabs
  a.q_resident = 1;
  a.num_slots = std::min(10, budget - 4 - 2 * a.dqk_slabs);
  if (a.num_slots < a.dqk_slabs + a.dv_slabs + 1) { set_last_error(__FILE__, __LINE__, "attention4: smem"); return -1; }
  const size_t smem = (size_t)(2 * a.dqk_slabs + a.num_slots + 4) * SLAB4 + 8 * (2 * a.num_slots + 9) + 16 + 1024;
  if (attention4_init() != 0) return -1;
  auto kern = bf16 ? attention4_kernel<true> : attention4_kernel<false>;
  dim3 grid((a.Nq + 255) / 256, a.B * a.H);
  SDXE_CUDA_CHECK(launch_k(kern, grid, dim3(ATT4_THREADS), smem, stream, a));
  return 0;
}

}  // namespace sdxe
