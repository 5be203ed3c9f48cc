// Two-query-tile flash attention for sm_100a (head dims up to 128): the variant used for the UNet's large
// self-attention layers. Same contract as attention.cu (AttnArgs), but one CTA owns 256 query rows (tiles A and B):
//
//  * every K / V slab fetched from L2 is used by two S = Q K^T and two O += P V tile products, which halves the
//    L2 -> shared-memory traffic per FLOP. With small head dims (SD1.5: d = 40) a 128-row tile is L2-bound at
//    ~80 FLOP/byte; this kernel doubles that.
//  * ONE MMA-ISSUING WARP PER TILE. A timeline trace of the single-issuer version (profiles/r2_attention_notes.md) showed
//    the issuing warp itself was the critical path: sharing a scheduler with four busy softmax warps, its serial chain of
//    polls, pops, 11 MMAs and 6 commits took ~3000 of each 4200-clk iteration while the 16 softmax warps idled ~1500 clk
//    waiting for S and PV. With an issuer per tile each chain is half as long and the issuers block in try_wait
//    (hardware sleep) instead of polling.
//  * S IS HANDED BACK AT ONCE. A softmax thread copies its row's 128 scores to registers and releases the TMEM buffer
//    (s_free) before doing any arithmetic, so S(i+1) = Q K(i+1)^T is computed DURING softmax(i) and the softmax warps
//    never wait for the tensor pipe: the round trip P(i) -> S(i+1) (>= 500 clk: wake-up, 3 MMAs, commit, wake-up) is off
//    the critical path. The only thing PV(i) gates is the P buffer, which softmax(i+1) needs ~1000 clk later.
//    With the scores in registers the row max is taken first and the exponentials once (no speculative pass, no
//    redo); the packed 16-bit P row overwrites the score registers in place. O is rescaled lazily (only when the row
//    max moves by more than 2^8).
//  * ONE SOFTMAX THREAD PER QUERY ROW (8 softmax warps: tile x TMEM lane quarter; 168 registers each hold the row's 128
//    scores): no cross-thread exchange of row maxima / sums at all. Each scheduler runs one warp of tile A and one of
//    tile B, and their exponential loops ALTERNATE (mufu_turn token per scheduler): while A's warp owns the MUFU pipe
//    (128 ex2: >= 1024 clk at 16 ex2/clk/SM), B's warp loads its next scores, finds the row max, stores P and hands
//    over — everything that is not MUFU work of one tile overlaps the MUFU work of the other. Measured (timeline
//    trace): a lone warp sustains one MUFU.EX2 per ~12 clk, not 8, so the exponential loop takes ~1550 clk and the
//    period of a tile pair is 2 x 1550 + hand-over; without the token the tiles fall into lock-step and it is 10 % worse.
//    Evaluating part of the exponentials on the FMA pipe (ex2_poly3, ATT2_POLY_MASK) makes this organisation slower
//    (the loop is issue- / latency-bound in one warp, not MUFU-throughput-bound): mask 0 is the default.
//  * P IN TENSOR MEMORY for head dims <= 64 (PT = true: SD1.5 level 0, all of SDXL): the softmax threads write the packed
//    16-bit P row with one tcgen05.st and PV runs as a TS-form MMA (A from TMEM). That removes 64 KB of st.shared + 64 KB
//    of MMA operand reads per iteration from the shared-memory port (which the K / V / Q operand reads and the TMA
//    writes need), the fence.proxy.async, and frees 64 KB of smem for a deeper K / V ring.
//    TMEM (PT): S_A | S_B | O_A | O_B | P_A | P_B at columns 0 / 128 / 256 / 320 / 384 / 448;
//    otherwise (d <= 128, P in smem): S_A | S_B | O_A | O_B at 0 / 128 / 256 / 384.
//
// Barrier protocol (all single-phase-per-iteration, parity = i & 1):
//   s_full[T]   MMA_T -> softmax_T : S_T(i) complete in TMEM
//   s_free[T]   softmax_T -> MMA_T : S_T(i) copied to registers (so S_T(i+1) may overwrite it)
//   mufu_turn[T][q]: softmax warp of the OTHER tile on scheduler q -> this tile's: its exponentials are done
//   p_ready[T]  softmax_T -> MMA_T : P_T(i) written
//   pv_done[T]  MMA_T -> softmax_T : O_T holds blocks <= i, P_T buffer free
//   slot_full / slot_empty: K / V ring (one TMA producer warp; a slab is released by BOTH issuers: count 2)
#include "attention.cuh"
#include <algorithm>

#ifndef SDXE_ATT_TRACE
#define SDXE_ATT_TRACE 0
#endif
// which of every 8 consecutive exponentials are evaluated on the FMA pipe instead of the MUFU pipe (bit j = element j)
#ifndef ATT2_POLY_MASK
#define ATT2_POLY_MASK 0x00
#endif
#ifndef ATT2_LAG
#define ATT2_LAG 1
#endif
// The MUFU turn is handed to the other tile's warp after this many of the 16 groups of 8 exponentials (16 = at the end).
// Half way (8) measured best: a lone warp sustains ~2/3 of the MUFU rate, so letting the second half of one row's
// exponentials overlap the first half of the other tile's keeps the pipe fuller than strict alternation (L0 846 -> 790 us,
// SDXL L1 550 -> 507 us; 4 / 6 / 7 / 9 / 10 / 12 are all worse: tools/gpu_scripts/attn_handover*.sh).
#ifndef ATT2_HANDOVER
#define ATT2_HANDOVER 8
#endif

namespace sdxe {

static constexpr int SLAB2 = 16384;
static constexpr int ATT2_THREADS = 384;   // warps 0-1 MMA issuers of tile A / B, warp 2 TMA producer, warp 3 idle, warps 4-11 softmax

template <bool BF16, bool PT>
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
  const uint32_t sP = sRing + NS * SLAB2;             // [2 tiles][2 slabs] (not allocated when P lives in TMEM)
  const uint32_t bar_base = sP + (PT ? 0 : 4 * SLAB2);
  auto slot_full = [&](int s) { return bar_base + 8u * s; };
  auto slot_empty = [&](int s) { return bar_base + 8u * (NS + s); };
  const uint32_t q_full = bar_base + 8u * (2 * NS);
  auto s_full = [&](int t) { return bar_base + 8u * (2 * NS + 1 + t); };
  auto p_ready = [&](int t) { return bar_base + 8u * (2 * NS + 3 + t); };
  auto pv_done = [&](int t) { return bar_base + 8u * (2 * NS + 5 + t); };
  auto s_free = [&](int t) { return bar_base + 8u * (2 * NS + 7 + t); };
  auto mufu_turn = [&](int t, int q) { return bar_base + 8u * (2 * NS + 9 + t * 4 + q); };
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + (bar_base - sbase) + 8 * (2 * NS + 17));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#if SDXE_ATT_TRACE
  // timeline of CTA (0,0): [role 0..3][iteration 0..47][event 0..7] clock64 stamps
  const bool tracing = a.trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0;
  auto TR = [&](int role, int it, int ev) {
    if (tracing && lane == 0 && it < 48) {
      unsigned long long c;
      asm volatile("mov.u64 %0, %%clock64;" : "=l"(c));
      a.trace[(role * 48 + it) * 8 + ev] = c;
    }
  };
#else
#define TR(role, it, ev) ((void)0)
#endif
  const int q0 = blockIdx.x * 256;
  const int bh = blockIdx.y;
  const int hb_b = bh / a.H, hb_h = bh - hb_b * a.H;  // (batch, head) coordinates of the 4D per-head tensor maps
  const int nblk = (a.Nk + 127) / 128;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NS; ++s) { mbar_init(slot_full(s), 1); mbar_init(slot_empty(s), 2); }
    mbar_init(q_full, 1);
    for (int t = 0; t < 2; ++t) {
      mbar_init(s_full(t), 1); mbar_init(s_free(t), 4); mbar_init(p_ready(t), 4); mbar_init(pv_done(t), 1);
      for (int q = 0; q < 4; ++q) mbar_init(mufu_turn(t, q), 1);
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

  if (warp <= 1) {
    // ---------------------------------------------------------------- MMA issuer of tile t (converged warp, elected issue;
    // see gemm.cu: issuing from a divergent single-lane region costs an ELECT/BRA.U.ANY loop per uniform-datapath
    // instruction).
    const int t = warp;
    const uint32_t idesc_s = umma_idesc(BF16 ? 1 : 0, 128, 128, 0, 0);
    const uint32_t idesc_pv = umma_idesc(BF16 ? 1 : 0, 128, 64, 0, 1);
    const int ksteps_last = (a.dqk - (QS - 1) * 64 + 15) / 16;
    const int n_last = (a.dv - (VS - 1) * 64 + 15) / 16 * 16;
    const uint32_t idesc_pv_last = umma_idesc(BF16 ? 1 : 0, 128, n_last, 0, 1);
    int slot = 0;
    uint32_t phase = 0;
    // Both issuers walk the same ring in the same order (waits on slot_full do not consume anything); a slab is handed
    // back to the producer once BOTH have committed on its slot_empty barrier. All descriptors are built once (Q, P:
    // constant per CTA) or once per slab (K, V: at pop time); inside the MMA loops only a 64-bit add remains.
    auto pop = [&](int& slot_id) -> uint32_t {  // wait for the next slab in ring order; caller releases it later
      mbar_wait(slot_full(slot), phase);
      slot_id = slot;
      const uint32_t addr = sRing + slot * SLAB2;
      if (++slot == NS) { slot = 0; phase ^= 1u; }
      return addr;
    };
    uint64_t qd[2], pd[2];
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      qd[c] = umma_desc_sw128(sQ + (t * QS + c) * SLAB2, 16, 1024);
      pd[c] = umma_desc_sw128(sP + (uint32_t)(t * 2 + c) * SLAB2, 16, 1024);
    }
    uint64_t kd[2] = {0, 0}, vd[2] = {0, 0};
    int k_slot[2] = {0, 0}, v_slot[2] = {0, 0};
    const uint32_t d_s = tmem_base + (uint32_t)(t * 128);
    auto issue_s = [&]() {
      if (elect_one()) {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          if (c < QS) {
            const int ks = (c == QS - 1) ? ksteps_last : 4;
#pragma unroll
            for (int k = 0; k < 4; ++k)
              if (k < ks) tc_mma_f16(d_s, qd[c] + 2 * k, kd[c] + 2 * k, idesc_s, (c | k) != 0 ? 1u : 0u);
          }
        }
        tc_commit(s_full(t));
      }
      __syncwarp();
    };
    auto issue_pv = [&](int j) {
      if (elect_one()) {
#pragma unroll
        for (int vs = 0; vs < 2; ++vs) {
          if (vs < VS) {
            const uint32_t d_o = tmem_base + 256u + (uint32_t)(t * (PT ? 64 : 128) + vs * 64);
            const uint32_t id = (vs == VS - 1) ? idesc_pv_last : idesc_pv;
#pragma unroll
            for (int k = 0; k < 8; ++k) {  // 16 key rows per step: +2048 B in V (= +128 in the addr>>4 field), +32 B in P
              if (PT) tc_mma_f16_ts(d_o, tmem_base + 384u + (uint32_t)(t * 64 + k * 8), vd[vs] + 128 * k, id, (j | k) != 0 ? 1u : 0u);
              else tc_mma_f16(d_o, pd[k >> 2] + 2 * (k & 3), vd[vs] + 128 * k, id, (j | k) != 0 ? 1u : 0u);
            }
          }
        }
        tc_commit(pv_done(t));
      }
      __syncwarp();
    };
    auto release = [&](const int* slots, int n) {  // this issuer's MMAs so far have finished reading these slabs
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
    issue_s();
    release(k_slot, QS);
    for (int i = 0; i < nblk; ++i) {
      const bool more = i + 1 < nblk;
      TR(2 + t, i, 0);
      if (more) {  // S(i+1) as soon as the softmax threads have copied S(i) out of tensor memory
#pragma unroll
        for (int c = 0; c < 2; ++c)
          if (c < QS) kd[c] = umma_desc_sw128(pop(k_slot[c]), 16, 1024);
        mbar_wait(s_free(t), (uint32_t)(i & 1));
        TR(2 + t, i, 1);
        tc_fence_after();
        issue_s();
      }
      TR(2 + t, i, 2);
#pragma unroll
      for (int vs = 0; vs < 2; ++vs)
        if (vs < VS) vd[vs] = umma_desc_sw128(pop(v_slot[vs]), SLAB2, 1024);
      mbar_wait(p_ready(t), (uint32_t)(i & 1));
      TR(2 + t, i, 3);
      tc_fence_after();
      issue_pv(i);
      TR(2 + t, i, 4);
      if (more) release(k_slot, QS);
      release(v_slot, VS);
      TR(2 + t, i, 5);
    }
  } else if (warp == 2) {
    // ---------------------------------------------------------------- TMA producer (converged warp, elected issue)
    if (elect_one()) {
      mbar_expect_tx(q_full, 2 * QS * SLAB2);
      for (int tt = 0; tt < 2; ++tt)
        for (int c = 0; c < QS; ++c) tma_load_4d(sQ + (tt * QS + c) * SLAB2, &a.tmQ, q_full, c * 64, q0 + tt * 128, hb_h, hb_b);
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
      if (i >= 1)
        for (int vs = 0; vs < VS; ++vs) push(&a.tmV, vs * 64, (i - 1) * 128);
    }
  } else if (warp >= 4) {
    // ---------------------------------------------------------------- softmax / epilogue: thread = query row
    const int quarter = warp & 3;
    const int t = (warp - 4) >> 2;  // tile
    const int row = quarter * 32 + lane;
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    const uint32_t t_s = tmem_base + (uint32_t)(t * 128) + lane_base;
    const uint32_t t_o = tmem_base + 256u + (uint32_t)(t * (PT ? 64 : 128)) + lane_base;
    const uint32_t t_p = tmem_base + 384u + (uint32_t)(t * 64) + lane_base;   // PT: the row's 64 packed P columns
    const uint32_t p_row = sP + (uint32_t)(t * 2) * SLAB2 + row * 128;        // !PT: two K-major 128B-swizzle atoms
    const float sl2 = a.scale_log2;
    float m_run = -INFINITY, l_run = 0.f;
#if SDXE_ATT_TRACE
    const bool trw = quarter == 0;
#endif
    for (int i = 0; i < nblk; ++i) {
#if SDXE_ATT_TRACE
      if (trw) TR(t, i, 0);
#endif
      mbar_wait(s_full(t), (uint32_t)(i & 1));
      tc_fence_after();
#if SDXE_ATT_TRACE
      if (trw) TR(t, i, 1);
#endif
      // the row's 128 scores -> registers, then the TMEM buffer goes straight back to the MMA issuer
      uint32_t r[128];
      tmem_ld32(t_s, r);
      tmem_ld32(t_s + 32, r + 32);
      tmem_ld32(t_s + 64, r + 64);
      tmem_ld32(t_s + 96, r + 96);
      tc_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_free(t));
      const int kv0 = i * 128;
      if (kv0 + 128 > a.Nk) {  // only the last block has invalid key columns
#pragma unroll
        for (int j = 0; j < 128; ++j)
          if (kv0 + j >= a.Nk) r[j] = 0xff800000u;  // -inf
      }
      float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
      for (int j = 0; j < 128; j += 4) {
        mx0 = fmaxf(mx0, __uint_as_float(r[j]));
        mx1 = fmaxf(mx1, __uint_as_float(r[j + 1]));
        mx2 = fmaxf(mx2, __uint_as_float(r[j + 2]));
        mx3 = fmaxf(mx3, __uint_as_float(r[j + 3]));
      }
      const float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
#if SDXE_ATT_TRACE
      if (trw) TR(t, i, 2);
#endif
      // lazy rescale: the running max only moves when the block max beats it by more than 2^8 (always on block 0);
      // until then P = 2^(s - m_run) <= 2^8 stays well inside the 16-bit range and O needs no correction
      const float m_cand = fmaxf(m_run, mx);
      const bool need = (m_cand - m_run) * sl2 > 8.f;  // first block: +inf > 8
      bool pv_waited = false;
      if (__any_sync(0xffffffffu, need)) {
        if (i >= 1) {
          mbar_wait(pv_done(t), (uint32_t)((i - 1) & 1));  // O_T holds blocks < i
          tc_fence_after();
          pv_waited = true;
          const float alpha = need ? ex2_approx((m_run - m_cand) * sl2) : 1.f;
          for (int c = 0; c < 2 * a.dv_slabs; ++c) {
            uint32_t o[32];
            tmem_ld32(t_o + c * 32, o);
            tc_wait_ld();
#pragma unroll
            for (int j = 0; j < 32; ++j) o[j] = __float_as_uint(__uint_as_float(o[j]) * alpha);
            tmem_st32(t_o + c * 32, o);
          }
          tc_wait_st();
          l_run *= alpha;
        }
        if (need) m_run = m_cand;
      }
      const float mb = m_run * sl2;
      // my tile's turn on this scheduler's MUFU pipe: the other tile's exponentials (B: block i, A: block i - 1) are done
      if (!(a.q_resident & 2) && (t == 1 || i >= 1)) mbar_wait(mufu_turn(t, quarter), (uint32_t)((t == 1 ? i : i - 1) & 1));
#if SDXE_ATT_TRACE
      if (trw) TR(t, i, 5);
#endif
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
      // 8 scores -> 4 packed registers, in place (r[4q..4q+3] <- r[8q..8q+7]), software-pipelined:
      // ATT2_LAG groups of 8 exponentials stay in flight: group q is issued (FFMA + MUFU.EX2), then group q - ATT2_LAG is
      // consumed (row sums, 16-bit packing into r[4q'..4q'+3]).
      float e[ATT2_LAG][8];
#pragma unroll
      for (int q = 0; q < 16 + ATT2_LAG; ++q) {
        if (ATT2_HANDOVER < 16 && q == ATT2_HANDOVER) {
          __syncwarp();
          if (lane == 0) mbar_arrive(mufu_turn(t ^ 1, quarter));
        }
        if (q >= ATT2_LAG) {
          const int c = q - ATT2_LAG;
          float* v = e[c % ATT2_LAG];
          s0 += v[0] + v[1]; s1 += v[2] + v[3]; s2 += v[4] + v[5]; s3 += v[6] + v[7];
          const uint32_t k0 = T::pack(v[0], v[1]), k1 = T::pack(v[2], v[3]), k2 = T::pack(v[4], v[5]), k3 = T::pack(v[6], v[7]);
          if (q < 16) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float x = fmaf(__uint_as_float(r[q * 8 + j]), sl2, -mb);
              v[j] = ((ATT2_POLY_MASK >> j) & 1) ? ex2_poly3(x) : ex2_approx(x);
            }
          }
          r[c * 4 + 0] = k0; r[c * 4 + 1] = k1; r[c * 4 + 2] = k2; r[c * 4 + 3] = k3;
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float x = fmaf(__uint_as_float(r[q * 8 + j]), sl2, -mb);
            e[q % ATT2_LAG][j] = ((ATT2_POLY_MASK >> j) & 1) ? ex2_poly3(x) : ex2_approx(x);
          }
        }
      }
      if (ATT2_HANDOVER >= 16) {
        __syncwarp();
        if (lane == 0) mbar_arrive(mufu_turn(t ^ 1, quarter));
      }
      l_run += (s0 + s1) + (s2 + s3);
#if SDXE_ATT_TRACE
      if (trw) TR(t, i, 3);
#endif
      if (i >= 1 && !pv_waited) {
        mbar_wait(pv_done(t), (uint32_t)((i - 1) & 1));  // P_T buffer free (PV(i-1) was issued ~one softmax ago)
        tc_fence_after();
      }
      if (PT) {
        tmem_st32(t_p, r);
        tmem_st32(t_p + 32, r + 32);
        tc_wait_st();
        tc_fence_before();
      } else {
#pragma unroll
        for (int ch = 0; ch < 16; ++ch) {  // 16-byte chunk ch & 7 of atom ch >> 3, 128B swizzle
          const uint32_t chunk = (uint32_t)(ch & 7) ^ (uint32_t)(row & 7);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(p_row + (uint32_t)(ch >> 3) * SLAB2 + chunk * 16),
                       "r"(r[ch * 4 + 0]), "r"(r[ch * 4 + 1]), "r"(r[ch * 4 + 2]), "r"(r[ch * 4 + 3])
                       : "memory");
        }
        tc_fence_before();
        fence_proxy_async_smem();
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(p_ready(t));
#if SDXE_ATT_TRACE
      if (trw) TR(t, i, 4);
#endif
    }
    // ---- epilogue
    mbar_wait(pv_done(t), (uint32_t)((nblk - 1) & 1));
    tc_fence_after();
    const int q = q0 + t * 128 + row;
    const float inv_l = 1.f / l_run;
    const int b = bh / a.H, h = bh - b * a.H;
    TT* orow = reinterpret_cast<TT*>(a.out) + ((size_t)b * a.Nq + q) * a.ldo + a.out_col0 + h * a.dv;
    for (int c = 0; c < 2 * a.dv_slabs; ++c) {
      if (c * 32 >= a.dv) break;
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

int attention2_init() {
  static bool done = false;
  if (!done) {
    SDXE_CUDA_CHECK(cudaFuncSetAttribute(attention2_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    SDXE_CUDA_CHECK(cudaFuncSetAttribute(attention2_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    SDXE_CUDA_CHECK(cudaFuncSetAttribute(attention2_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    SDXE_CUDA_CHECK(cudaFuncSetAttribute(attention2_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
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
  { static int nt = -1; if (nt < 0) { const char* e = getenv("SDXE_ATT_NOTOKEN"); nt = e ? atoi(e) : 0; } if (nt) a.q_resident |= 2; }
  a.trace = nullptr;
#if SDXE_ATT_TRACE
  static unsigned long long* trace_buf = nullptr;
  static int trace_mode = -1;
  if (trace_mode < 0) { const char* e = getenv("SDXE_ATT_TRACE_DUMP"); trace_mode = e ? atoi(e) : 0; }
  if (trace_mode == 1) {
    if (!trace_buf) cudaMalloc(&trace_buf, 4 * 48 * 8 * 8);
    cudaMemsetAsync(trace_buf, 0, 4 * 48 * 8 * 8, stream);
    a.trace = trace_buf;
  }
#endif
  const bool pt = a.dqk_slabs == 1 && a.dv_slabs == 1;  // head dim <= 64: P lives in tensor memory
  const int p_slabs = pt ? 0 : 4;
  a.num_slots = std::min(10, budget - p_slabs - 2 * a.dqk_slabs);
  if (a.num_slots < a.dqk_slabs + a.dv_slabs + 1) { set_last_error(__FILE__, __LINE__, "attention2: smem"); return -1; }
  const size_t smem = (size_t)(2 * a.dqk_slabs + a.num_slots + p_slabs) * SLAB2 + 8 * (2 * a.num_slots + 17) + 16 + 1024;
  if (attention2_init() != 0) return -1;
  auto kern = pt ? (bf16 ? attention2_kernel<true, true> : attention2_kernel<false, true>)
                 : (bf16 ? attention2_kernel<true, false> : attention2_kernel<false, false>);
  dim3 grid((a.Nq + 255) / 256, a.B * a.H);
  SDXE_CUDA_CHECK(launch_k(kern, grid, dim3(ATT2_THREADS), smem, stream, a));
#if SDXE_ATT_TRACE
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
#endif
  return 0;
}

}  // namespace sdxe
