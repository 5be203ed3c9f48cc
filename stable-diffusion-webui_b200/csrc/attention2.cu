// Two-query-tile flash attention for sm_100a (head dims up to 128): the variant used for the UNet's large
// self-attention layers. Same contract as attention.cu (AttnArgs), but one CTA owns 256 query rows (tiles A and B):
//
//  * every K / V slab fetched from L2 is used by two S = Q K^T and two O += P V tile products, which halves the
//    L2 -> shared-memory traffic per FLOP. With small head dims (SD1.5: d = 40) a 128-row tile is L2-bound at
//    ~80 FLOP/byte; this kernel doubles that.
//  * the two tiles ping-pong on the tensor pipe: while tile A's rows are in their softmax, the MMA thread issues tile
//    B's S and PV, and vice versa (the FA-4 schedule). 16 softmax warps (tile x column-half x lane quarter): every
//    scheduler interleaves four of them.
//  * TMEM: S_A | S_B | O_A | O_B at columns 0 / 128 / 256 / 384 (fp32, 128 columns each).
//
// Barrier protocol (all single-phase-per-iteration, parity = i & 1):
//   s_full[T]   MMA -> softmax_T : S_T(i) complete in TMEM
//   p_ready[T]  softmax_T -> MMA : P_T(i) in smem, S_T(i) consumed (so S_T(i+1) may overwrite it)
//   pv_done[T]  MMA -> softmax_T : O_T holds blocks <= i, P_T buffer free
#include "attention.cuh"
#include <algorithm>

namespace sdxe {

static constexpr int SLAB2 = 16384;
static constexpr int ATT2_THREADS = 576;  // warp 0 TMA, warp 1 MMA, 16 softmax warps

template <bool BF16>
__global__ void __launch_bounds__(ATT2_THREADS, 1) attention2_kernel(const __grid_constant__ AttnArgs a) {
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
  const uint32_t sRing = sQ + 2 * QS * SLAB2;
  const uint32_t sP = sRing + NS * SLAB2;             // [2 tiles][2 slabs]
  const uint32_t bar_base = sP + 4 * SLAB2;
  auto slot_full = [&](int s) { return bar_base + 8u * s; };
  auto slot_empty = [&](int s) { return bar_base + 8u * (NS + s); };
  const uint32_t q_full = bar_base + 8u * (2 * NS);
  auto s_full = [&](int t) { return bar_base + 8u * (2 * NS + 1 + t); };
  auto p_ready = [&](int t) { return bar_base + 8u * (2 * NS + 3 + t); };
  auto pv_done = [&](int t) { return bar_base + 8u * (2 * NS + 5 + t); };
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + (bar_base - sbase) + 8 * (2 * NS + 7));
  float* xch = reinterpret_cast<float*>(smem + (bar_base - sbase) + 8 * (2 * NS + 7) + 16);  // row max / sum exchange

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // timeline trace (debug): slot layout [role 0..3][iteration 0..47][event 0..7]
  const bool tracing = a.trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0;
  auto TR = [&](int role, int it, int ev) {
    if (tracing && lane == 0 && it < 48) {
      unsigned long long c;
      asm volatile("mov.u64 %0, %%clock64;" : "=l"(c));
      a.trace[(role * 48 + it) * 8 + ev] = c;
    }
  };
  const int q0 = blockIdx.x * 256;
  const int bh = blockIdx.y;
  const int hb_b = bh / a.H, hb_h = bh - hb_b * a.H;  // (batch, head) coordinates of the 4D per-head tensor maps
  const int nblk = (a.Nk + 127) / 128;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NS; ++s) { mbar_init(slot_full(s), 1); mbar_init(slot_empty(s), 1); }
    mbar_init(q_full, 1);
    for (int t = 0; t < 2; ++t) { mbar_init(s_full(t), 1); mbar_init(p_ready(t), 8); mbar_init(pv_done(t), 1); }
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
    // ---------------------------------------------------------------- producer
    // converged warp, elected issue (see gemm.cu: issuing from a divergent single-lane region costs an ELECT/BRA.U.ANY
    // loop per uniform-datapath instruction)
    {
      if (elect_one()) {
        mbar_expect_tx(q_full, 2 * QS * SLAB2);
        for (int t = 0; t < 2; ++t)
          for (int c = 0; c < QS; ++c) tma_load_4d(sQ + (t * QS + c) * SLAB2, &a.tmQ, q_full, c * 64, q0 + t * 128, hb_h, hb_b);
      }
      __syncwarp();
      int slot = 0;
      uint32_t phase = 0;
      auto push = [&](const CUtensorMap* tm, int c0, int r0) {
        mbar_wait(slot_empty(slot), phase ^ 1u);
        if (elect_one()) {
          mbar_expect_tx(slot_full(slot), SLAB2);
          tma_load_4d(sRing + slot * SLAB2, tm, slot_full(slot), c0, r0, hb_h, hb_b);
        }
        __syncwarp();
        if (++slot == NS) { slot = 0; phase ^= 1u; }
      };
      // ring order == consumption order: K_0, (K_1, V_0), (K_2, V_1), ..., V_{n-1}
      for (int i = 0; i <= nblk; ++i) {
        if (i < nblk)
          for (int c = 0; c < QS; ++c) push(&a.tmK, c * 64, i * 128);
        TR(3, i, 0);
        if (i >= 1)
          for (int vs = 0; vs < VS; ++vs) push(&a.tmV, vs * 64, (i - 1) * 128);
        TR(3, i, 1);
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer (converged warp, elected issue)
    {
      const uint32_t idesc_s = umma_idesc(BF16 ? 1 : 0, 128, 128, 0, 0);
      const uint32_t idesc_pv = umma_idesc(BF16 ? 1 : 0, 128, 64, 0, 1);
      const int ksteps_last = (a.dqk - (QS - 1) * 64 + 15) / 16;
      const int n_last = (a.dv - (VS - 1) * 64 + 15) / 16 * 16;
      const uint32_t idesc_pv_last = umma_idesc(BF16 ? 1 : 0, 128, n_last, 0, 1);
      int slot = 0;
      uint32_t phase = 0;
      // The issuing thread is a single lane: every instruction it spends on descriptor arithmetic delays S / PV for
      // 256 softmax threads. All descriptors are therefore built once (Q, P: constant per CTA) or once per slab
      // (K, V: at pop time); inside the MMA loops only a 64-bit add remains. Loops have constant bounds (<= 2 slabs)
      // so nothing is indexed dynamically.
      auto pop = [&](int& slot_id) -> uint32_t {  // wait for the next slab in ring order; caller releases it later
        mbar_wait(slot_full(slot), phase);
        slot_id = slot;
        const uint32_t addr = sRing + slot * SLAB2;
        if (++slot == NS) { slot = 0; phase ^= 1u; }
        return addr;
      };
      uint64_t qd[2][2], pd[2][2];
#pragma unroll
      for (int t = 0; t < 2; ++t) {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          qd[t][c] = umma_desc_sw128(sQ + (t * QS + c) * SLAB2, 16, 1024);
          pd[t][c] = umma_desc_sw128(sP + (uint32_t)(t * 2 + c) * SLAB2, 16, 1024);
        }
      }
      uint64_t kd[2] = {0, 0}, vd[2] = {0, 0};
      int k_slot[2] = {0, 0}, v_slot[2] = {0, 0};
      auto issue_s = [&](int t) {
        const uint32_t d_s = tmem_base + (uint32_t)(t * 128);
        if (elect_one()) {
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            if (c < QS && !(a.dbg & 128)) {
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
            if (vs < VS && !(a.dbg & 64)) {
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
        TR(2, i, 0);
        if (more) {
#pragma unroll
          for (int c = 0; c < 2; ++c)
            if (c < QS) kd[c] = umma_desc_sw128(pop(k_slot[c]), 16, 1024);
        }
#pragma unroll
        for (int vs = 0; vs < 2; ++vs)
          if (vs < VS) vd[vs] = umma_desc_sw128(pop(v_slot[vs]), SLAB2, 1024);
        TR(2, i, 1);
        // the two tiles are served in ARRIVAL order (a tile that finishes its softmax first must not wait for the
        // other one's p_ready: that is what lets the tiles drift half a period apart and alternate on the MUFU pipe)
        {
          bool served0 = false, served1 = false;
          for (uint32_t spin = 0; !(served0 && served1); ++spin) {
            if (!served0 && __any_sync(0xffffffffu, mbar_test(p_ready(0), (uint32_t)(i & 1)))) {
              TR(2, i, 2);
              tc_fence_after();
              if (more) issue_s(0);
              issue_pv(0, i);
              TR(2, i, 3);
              served0 = true;
            }
            if (!served1 && __any_sync(0xffffffffu, mbar_test(p_ready(1), (uint32_t)(i & 1)))) {
              TR(2, i, 4);
              tc_fence_after();
              if (more) issue_s(1);
              issue_pv(1, i);
              TR(2, i, 5);
              served1 = true;
            }
            if (spin > (1u << 28)) {
              printf("sdxe: attention2 MMA warp watchdog block(%d,%d) i %d\n", blockIdx.x, blockIdx.y, i);
              __trap();
            }
          }
        }
        if (more) release(k_slot, QS);
        release(v_slot, VS);
      }
    }
  } else {
    // ---------------------------------------------------------------- softmax / epilogue
    // 16 warps: (tile, column half) x 4 lane quarters. A query row of a tile is shared by two threads (warps w and
    // w + 8 own the same TMEM lane quarter of the same tile); one takes key columns 0-63 of the block, the other
    // 64-127. Four softmax warps per scheduler keep the MUFU / FMA pipes busy through each other's latencies; the
    // softmax is latency-bound with fewer (measured: 1 or 2 warps per scheduler -> ~35 % issue utilisation).
    const int sw = warp - 2;
    const int quarter = warp & 3;
    const int t = (sw >> 2) & 1;   // tile
    const int half = sw >> 3;      // column half
    const int row = quarter * 32 + lane;
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    const uint32_t t_s = tmem_base + (uint32_t)(t * 128 + half * 64) + lane_base;
    const uint32_t t_o = tmem_base + 256u + (uint32_t)(t * 128) + lane_base;
    const uint32_t p_row = sP + (uint32_t)(t * 2 + half) * SLAB2 + row * 128;  // K-major swizzle atom `half` of P_t
    const uint32_t pair_bar = 1u + (uint32_t)(t * 4 + quarter);               // the two warps sharing these 32 rows
    float* xq = xch + t * 256;                                                 // [2 buf][2 tiles][2 halves][128]
    const float sl2 = a.scale_log2;
    const int ko = a.dbg;
    float m_run = -INFINITY, l_run = 0.f;
    const bool trw = half == 0 && quarter == 0;
    for (int i = 0; i < nblk; ++i) {
      if (trw) TR(t, i, 0);
      mbar_wait(s_full(t), (uint32_t)(i & 1));
      tc_fence_after();
      if (trw) TR(t, i, 1);
      const int kv0 = i * 128 + half * 64;
      const bool tail = kv0 + 64 > a.Nk;  // only the last block has invalid key columns
      if (i >= 1) {
        mbar_wait(pv_done(t), (uint32_t)((i - 1) & 1));  // O_T holds blocks < i, P_T buffer free
        tc_fence_after();
      }
      // TMEM reads run at ~64 B/clk/SM: one pass over the 128 x 128 fp32 score tile costs as much as its 16 K exp2 on
      // the MUFU pipe, so the scores must be read ONCE. The running max is kept stale on purpose (lazy rescale), which
      // lets the common case exponentiate against it in the same pass that finds the block max; only when the block
      // max beats the stale one by more than 2^8 (first block, then rarely) is the pass repeated with the new max.
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
      float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
      auto pass = [&](float mb, bool with_max) {
        s0 = s1 = s2 = s3 = 0.f;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t r[32];
          if (!(ko & 8)) {
            tmem_ld32(t_s + c * 32, r);
            tc_wait_ld();
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(-1.f - 0.01f * j);
          }
          if (tail) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (kv0 + c * 32 + j >= a.Nk) r[j] = 0xff800000u;  // -inf
          }
          if (with_max && !(ko & 16)) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              mx0 = fmaxf(mx0, __uint_as_float(r[j]));
              mx1 = fmaxf(mx1, __uint_as_float(r[j + 1]));
              mx2 = fmaxf(mx2, __uint_as_float(r[j + 2]));
              mx3 = fmaxf(mx3, __uint_as_float(r[j + 3]));
            }
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float p0, p1, p2, p3, p4, p5, p6, p7;
            if (!(ko & 1)) {
              p0 = ex2_approx(fmaf(__uint_as_float(r[q * 8 + 0]), sl2, -mb));
              p1 = ex2_approx(fmaf(__uint_as_float(r[q * 8 + 1]), sl2, -mb));
              p2 = ex2_approx(fmaf(__uint_as_float(r[q * 8 + 2]), sl2, -mb));
              p3 = ex2_approx(fmaf(__uint_as_float(r[q * 8 + 3]), sl2, -mb));
              p4 = ex2_approx(fmaf(__uint_as_float(r[q * 8 + 4]), sl2, -mb));
              p5 = ex2_approx(fmaf(__uint_as_float(r[q * 8 + 5]), sl2, -mb));
              p6 = ex2_approx(fmaf(__uint_as_float(r[q * 8 + 6]), sl2, -mb));
              p7 = ex2_approx(fmaf(__uint_as_float(r[q * 8 + 7]), sl2, -mb));
            } else {
              p0 = fmaf(__uint_as_float(r[q * 8 + 0]), sl2, -mb);
              p1 = fmaf(__uint_as_float(r[q * 8 + 1]), sl2, -mb);
              p2 = fmaf(__uint_as_float(r[q * 8 + 2]), sl2, -mb);
              p3 = fmaf(__uint_as_float(r[q * 8 + 3]), sl2, -mb);
              p4 = fmaf(__uint_as_float(r[q * 8 + 4]), sl2, -mb);
              p5 = fmaf(__uint_as_float(r[q * 8 + 5]), sl2, -mb);
              p6 = fmaf(__uint_as_float(r[q * 8 + 6]), sl2, -mb);
              p7 = fmaf(__uint_as_float(r[q * 8 + 7]), sl2, -mb);
            }
            s0 += p0 + p1; s1 += p2 + p3; s2 += p4 + p5; s3 += p6 + p7;
            const uint32_t chunk = (uint32_t)(c * 4 + q) ^ (uint32_t)(row & 7);  // 16-byte chunk of this row, 128B swizzle
            if (!(ko & 2)) asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(p_row + chunk * 16),
                         "r"(T::pack(p0, p1)), "r"(T::pack(p2, p3)), "r"(T::pack(p4, p5)), "r"(T::pack(p6, p7))
                         : "memory");
          }
        }
      };
      // speculative pass against the stale max (first block: m_run = -inf -> mb = -inf -> p = inf/NaN garbage that the
      // mandatory redo below overwrites; the sums are recomputed by the redo as well)
      if (trw) TR(t, i, 2);
      pass(i == 0 ? 0.f : m_run * sl2, true);
      if (trw) TR(t, i, 3);
      float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
      // agree on the row max with the thread that owns the other 64 columns
      float* xb = xq + (i & 1) * 512;
      if (!(ko & 4)) {
        xb[half * 128 + row] = mx;
        asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
        mx = fmaxf(mx, xb[(half ^ 1) * 128 + row]);
      }
      const float m_cand = fmaxf(m_run, mx);
      const bool need = (ko & 32) ? (i == 0) : ((m_cand - m_run) * sl2 > 8.f);  // first block: +inf > 8
      if (__any_sync(0xffffffffu, need)) {  // same rows, same decision in both warps of the pair
        const float alpha = ex2_approx((m_run - m_cand) * sl2);
        if (i >= 1) {
          for (int c = half * a.dv_slabs; c < (half + 1) * a.dv_slabs; ++c) {  // each half rescales its O columns
            uint32_t r[32];
            tmem_ld32(t_o + c * 32, r);
            tc_wait_ld();
#pragma unroll
            for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) * alpha);
            tmem_st32(t_o + c * 32, r);
          }
          tc_wait_st();
        }
        l_run *= alpha;
        m_run = m_cand;
        pass(m_run * sl2, false);  // redo this block against the new max
      }
      l_run += (s0 + s1) + (s2 + s3);
      if (trw) TR(t, i, 4);
      tc_fence_before();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_ready(t));
      if (trw) TR(t, i, 5);
    }
    // ---- epilogue: the row sum is the sum of the two halves' partial sums
    {
      float* xb = xq + (nblk & 1) * 512;
      xb[half * 128 + row] = l_run;
      asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
      l_run += xb[(half ^ 1) * 128 + row];
    }
    mbar_wait(pv_done(t), (uint32_t)((nblk - 1) & 1));
    tc_fence_after();
    const int q = q0 + t * 128 + row;
    const float inv_l = 1.f / l_run;
    const int b = bh / a.H, h = bh - b * a.H;
    TT* orow = reinterpret_cast<TT*>(a.out) + ((size_t)b * a.Nq + q) * a.ldo + a.out_col0 + h * a.dv;
    for (int c = half * a.dv_slabs; c < (half + 1) * a.dv_slabs; ++c) {
      if (c * 32 >= a.dv) break;
      uint32_t r[32];
      tmem_ld32(t_o + c * 32, r);
      tc_wait_ld();
      if (q < a.Nq) {
#pragma unroll
        for (int g = 0; g < 32; g += 8) {
          if (c * 32 + g + 8 <= a.dv) {
            uint4 u;
            u.x = T::pack(__uint_as_float(r[g + 0]) * inv_l, __uint_as_float(r[g + 1]) * inv_l);
            u.y = T::pack(__uint_as_float(r[g + 2]) * inv_l, __uint_as_float(r[g + 3]) * inv_l);
            u.z = T::pack(__uint_as_float(r[g + 4]) * inv_l, __uint_as_float(r[g + 5]) * inv_l);
            u.w = T::pack(__uint_as_float(r[g + 6]) * inv_l, __uint_as_float(r[g + 7]) * inv_l);
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

int attention2_init() {
  static bool done = false;
  if (!done) {
    SDXE_CUDA_CHECK(cudaFuncSetAttribute(attention2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    SDXE_CUDA_CHECK(cudaFuncSetAttribute(attention2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    done = true;
  }
  return 0;
}

bool attention2_eligible(const AttnArgs& a) {
  return a.dqk_slabs <= 2 && a.dv_slabs <= 2 && a.Nq >= 256 && a.dv <= 128;
}

int attention2_launch(const AttnArgs& a_in, bool bf16, cudaStream_t stream) {
  AttnArgs a = a_in;
  const int budget = (224 * 1024 - 2048) / SLAB2;  // 13 slabs
  a.q_resident = 1;
  {
    static int ko = -1;
    if (ko < 0) { const char* e = getenv("SDXE_ATT_KO"); ko = e ? atoi(e) : 0; }
    a.dbg = ko;
  }
  static unsigned long long* trace_buf = nullptr;
  static int trace_mode = -1;
  if (trace_mode < 0) { const char* e = getenv("SDXE_ATT_TRACE"); trace_mode = e ? atoi(e) : 0; }
  a.trace = nullptr;
  if (trace_mode == 1) {
    if (!trace_buf) { cudaMalloc(&trace_buf, 4 * 48 * 8 * 8); }
    cudaMemsetAsync(trace_buf, 0, 4 * 48 * 8 * 8, stream);
    a.trace = trace_buf;
  }
  a.num_slots = std::min(10, budget - 4 - 2 * a.dqk_slabs);
  if (a.num_slots < a.dqk_slabs + a.dv_slabs + 1) { set_last_error(__FILE__, __LINE__, "attention2: smem"); return -1; }
  const size_t smem = (size_t)(2 * a.dqk_slabs + a.num_slots + 4) * SLAB2 + 8 * (2 * a.num_slots + 7) + 16 + 4096 + 1024;
  if (attention2_init() != 0) return -1;
  auto kern = bf16 ? attention2_kernel<true> : attention2_kernel<false>;
  dim3 grid((a.Nq + 255) / 256, a.B * a.H);
  SDXE_CUDA_CHECK(launch_k(kern, grid, dim3(ATT2_THREADS), smem, stream, a));
  if (trace_mode == 1) {
    trace_mode = 2;  // once
    cudaStreamSynchronize(stream);
    static unsigned long long h[4 * 48 * 8];
    cudaMemcpy(h, trace_buf, sizeof(h), cudaMemcpyDeviceToHost);
    FILE* f = fopen("gpurun_out/attn_trace.txt", "w");
    if (f) {
      unsigned long long t0 = ~0ull;
      for (auto v : h) if (v && v < t0) t0 = v;
      for (int role = 0; role < 4; ++role)
        for (int it = 0; it < 48; ++it) {
          fprintf(f, "role %d it %2d:", role, it);
          for (int ev = 0; ev < 8; ++ev) fprintf(f, " %8lld", h[(role * 48 + it) * 8 + ev] ? (long long)(h[(role * 48 + it) * 8 + ev] - t0) : -1ll);
          fprintf(f, "\n");
        }
      fclose(f);
    }
  }
  return 0;
}

}  // namespace sdxe
