// Short-KV (cross-) attention for sm_100a: all keys of a head fit one block (Nk <= 128, e.g. the 77 CLIP tokens), head
// dim <= 64. Same contract as attention.cu (AttnArgs).
//
// With a single key block there is no online-softmax loop to hide latencies behind: one (head, 128-query-row) item is
// a strictly serial chain TMA -> S = Q K^T -> softmax -> P -> O = P V -> store (~8 us end to end), and a one-item CTA
// per SM (the general kernels' shape) runs the layer at that latency: 118 us for 4096 items on 148 SMs where the
// exponentials need ~15 us. This kernel is PERSISTENT and software-pipelined across items instead:
//
//   * a CTA owns a CONTIGUOUS range of items, so K and V are fetched once per head (two K|V buffers, one per "epoch")
//     and only Q streams per item, five stages deep; S and O accumulators (TMEM) and P (smem) are double buffered;
//   * the MMA warp issues S(k+1) before it waits for P(k), so the tensor pipe and TMA work one item ahead of the
//     softmax warps; two softmax groups of 8 warps take even / odd items (each owns one S / P / O buffer), so four
//     warps per scheduler cover each other's MUFU, TMEM and barrier latencies; a group stores O(k-2) just before it
//     starts on item k;
//   * only the key columns that exist are touched: S is computed N = ceil(Nk/16)*16 wide (80 for 77 keys), a softmax
//     thread owns half of those columns of one query row (40: no wasted exponentials), O = P V runs N/16 k-steps.
//
// Barriers per buffer b = k & 1 (k = CTA-local item counter; "full" parity (k >> 1) & 1, "free" parity flipped);
// the Q stages cycle as k % 5 with parity (k / 5) & 1; K|V buffers by epoch e = head(k) - head(0), e & 1, (e >> 1) & 1:
//   stage_full  TMA -> MMA          Q of item k landed
//   stage_free  MMA commit -> TMA   S(k) has consumed the Q slab
//   kv_full     TMA -> MMA          K, V of the epoch landed
//   kv_free     MMA commit -> TMA   the epoch's last PV has been issued and completed
//   s_full      MMA commit -> softmax
//   s_free      softmax -> MMA      scores copied to registers (8 warps)
//   p_ready     softmax -> MMA      P(k) in smem (8 warps)
//   o_full      MMA commit -> softmax
//   o_free      softmax -> MMA      O(k) stored (8 warps)
#include "attention.cuh"
#include <algorithm>

namespace sdxe {

static constexpr int XSLAB = 16384;
static constexpr int ATTX_THREADS = 576;  // warp 0 TMA, warp 1 MMA, 2 softmax groups x 8 warps (column half x TMEM lane quarter)
static constexpr uint32_t XTM_O = 256;    // TMEM: S0 | S1 at 0 / 128, O0 | O1 at 256 / 320
static constexpr int XSTG = 5;            // Q stages: the strided 80..128-byte row gathers have a long TMA round trip

template <bool BF16, int CPH>  // CPH: key columns per softmax thread (40 covers Nk <= 80, 64 covers Nk <= 128)
__global__ void __launch_bounds__(ATTX_THREADS, 1) attentionx_kernel(const __grid_constant__ AttnArgs a) {
  using T = T16<BF16>;
  using TT = typename T::type;
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t pad = ((raw + 1023u) & ~1023u) - raw;
  uint8_t* smem = smem_raw + pad;
  const uint32_t sbase = raw + pad;

  const uint32_t sStage = sbase;                  // [XSTG] Q slabs
  const uint32_t sKV = sStage + XSTG * XSLAB;     // [2 epochs][K | V] slabs
  const uint32_t sP = sKV + 4 * XSLAB;            // [2][2 slabs] K-major, 64 keys per slab
  const uint32_t bar_base = sP + 4 * XSLAB;
  auto stage_full = [&](int s) { return bar_base + 8u * (0 + s); };
  auto stage_free = [&](int s) { return bar_base + 8u * (8 + s); };
  auto s_full = [&](int b) { return bar_base + 8u * (16 + b); };
  auto s_free = [&](int b) { return bar_base + 8u * (18 + b); };
  auto p_ready = [&](int b) { return bar_base + 8u * (20 + b); };
  auto o_full = [&](int b) { return bar_base + 8u * (22 + b); };
  auto o_free = [&](int b) { return bar_base + 8u * (24 + b); };
  auto kv_full = [&](int e) { return bar_base + 8u * (26 + e); };
  auto kv_free = [&](int e) { return bar_base + 8u * (28 + e); };
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + (bar_base - sbase) + 8 * 30);
  float* xch = reinterpret_cast<float*>(smem + (bar_base - sbase) + 8 * 30 + 16);  // [max | sum][2 buffers][2 halves][128]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qtiles = (a.Nq + 127) / 128;
  const int total = a.B * a.H * qtiles;
  const int G = gridDim.x;
  // contiguous item range of this CTA: consecutive items are consecutive query tiles of the same head
  const int item0 = (int)(((long long)blockIdx.x * total) / G);
  const int n_items = (int)(((long long)(blockIdx.x + 1) * total) / G) - item0;
  const int bh0 = item0 / qtiles;
  const int ncol = (a.Nk + 15) / 16 * 16;                      // key columns computed (multiple of 16, <= 128)
  const int cph = ncol >> 1;                                   // per softmax thread (multiple of 8, <= CPH)

  if (threadIdx.x == 0) {
    for (int st = 0; st < XSTG; ++st) { mbar_init(stage_full(st), 1); mbar_init(stage_free(st), 1); }
    for (int b = 0; b < 2; ++b) {
      mbar_init(s_full(b), 1);
      mbar_init(s_free(b), 8);
      mbar_init(p_ready(b), 8);
      mbar_init(o_full(b), 1);
      mbar_init(o_free(b), 8);
      mbar_init(kv_full(b), 1);
      mbar_init(kv_free(b), 1);
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
  pdl_wait();

  if (warp == 0) {
    // ---------------------------------------------------------------- producer (converged warp, elected issue)
    int sg = 0;
    uint32_t sg_phase = 0;
    int prev_bh = -1;
    for (int k = 0; k < n_items; ++k) {
      const int item = item0 + k;
      const int bh = item / qtiles, qt = item - bh * qtiles;
      const int hb_b = bh / a.H, hb_h = bh - hb_b * a.H;
      if (bh != prev_bh) {  // new head: its K and V go to the other K|V buffer
        prev_bh = bh;
        const int e = bh - bh0;
        mbar_wait(kv_free(e & 1), (uint32_t)(((e >> 1) & 1) ^ 1));
        if (elect_one()) {
          const uint32_t kv = sKV + (uint32_t)(e & 1) * 2 * XSLAB;
          mbar_expect_tx(kv_full(e & 1), 2 * XSLAB);
          tma_load_4d(kv, &a.tmK, kv_full(e & 1), 0, 0, hb_h, hb_b);
          tma_load_4d(kv + XSLAB, &a.tmV, kv_full(e & 1), 0, 0, hb_h, hb_b);
        }
        __syncwarp();
      }
      mbar_wait(stage_free(sg), sg_phase ^ 1u);
      if (elect_one()) {
        mbar_expect_tx(stage_full(sg), XSLAB);
        tma_load_4d(sStage + (uint32_t)sg * XSLAB, &a.tmQ, stage_full(sg), 0, qt * 128, hb_h, hb_b);
      }
      __syncwarp();
      if (++sg == XSTG) { sg = 0; sg_phase ^= 1u; }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer (converged warp, elected issue)
    const uint32_t idesc_s = umma_idesc(BF16 ? 1 : 0, 128, ncol, 0, 0);
    const int ksteps_s = (a.dqk + 15) / 16;
    const int n_o = (a.dv + 15) / 16 * 16;
    const uint32_t idesc_pv = umma_idesc(BF16 ? 1 : 0, 128, n_o, 0, 1);
    const int ksteps_pv = ncol / 16;
    auto epoch_of = [&](int k) { return (item0 + k) / qtiles - bh0; };
    auto issue_s = [&](int k) {
      const int b = k & 1, sg = k % XSTG, e = epoch_of(k);
      mbar_wait(kv_full(e & 1), (uint32_t)((e >> 1) & 1));
      mbar_wait(stage_full(sg), (uint32_t)((k / XSTG) & 1));
      mbar_wait(s_free(b), (uint32_t)(((k >> 1) & 1) ^ 1));
      tc_fence_after();
      if (elect_one()) {
        const uint64_t qd = umma_desc_sw128(sStage + (uint32_t)sg * XSLAB, 16, 1024);
        const uint64_t kd = umma_desc_sw128(sKV + (uint32_t)(e & 1) * 2 * XSLAB, 16, 1024);
        const uint32_t d_s = tmem_base + (uint32_t)(b * 128);
#pragma unroll
        for (int s = 0; s < 4; ++s)
          if (s < ksteps_s) tc_mma_f16(d_s, qd + 2 * s, kd + 2 * s, idesc_s, s != 0 ? 1u : 0u);
        tc_commit(s_full(b));
        tc_commit(stage_free(sg));  // the Q slab is free as soon as S(k) has been computed
      }
      __syncwarp();
    };
    if (n_items > 0) issue_s(0);
    for (int k = 0; k < n_items; ++k) {
      const int b = k & 1;
      if (k + 1 < n_items) issue_s(k + 1);  // one item ahead of the softmax
      mbar_wait(p_ready(b), (uint32_t)((k >> 1) & 1));
      mbar_wait(o_free(b), (uint32_t)(((k >> 1) & 1) ^ 1));
      tc_fence_after();
      const int e = epoch_of(k);
      const bool last_of_epoch = (k + 1 == n_items) || epoch_of(k + 1) != e;
      if (elect_one()) {
        const uint64_t vd = umma_desc_sw128(sKV + (uint32_t)(e & 1) * 2 * XSLAB + XSLAB, XSLAB, 1024);  // MN-major: 16 key rows = +2048 B
        const uint64_t pd0 = umma_desc_sw128(sP + (uint32_t)(b * 2) * XSLAB, 16, 1024);
        const uint64_t pd1 = umma_desc_sw128(sP + (uint32_t)(b * 2 + 1) * XSLAB, 16, 1024);
        const uint32_t d_o = tmem_base + XTM_O + (uint32_t)(b * 64);
#pragma unroll
        for (int s = 0; s < 8; ++s)
          if (s < ksteps_pv) tc_mma_f16(d_o, (s < 4 ? pd0 : pd1) + 2 * (s & 3), vd + 128 * s, idesc_pv, s != 0 ? 1u : 0u);
        tc_commit(o_full(b));
        if (last_of_epoch) tc_commit(kv_free(e & 1));  // every MMA that reads this head's K | V has been issued
      }
      __syncwarp();
    }
  } else {
    // ---------------------------------------------------------------- softmax / epilogue
    const int grp = (warp - 2) >> 3;        // group g owns buffer g: items k = g, g + 2, ...
    const int quarter = warp & 3;           // TMEM lane quarter (hardware: warp id mod 4)
    const int half = ((warp - 2) & 7) >> 2; // key columns [half * cph, half * cph + cph)
    const int row = quarter * 32 + lane;
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    const uint32_t pair_bar = 1u + (uint32_t)(grp * 4 + quarter);  // the two warps that share these 32 rows of this group
    const float sl2 = a.scale_log2;
    const int c_lo = half * cph;
    const int b = grp;
    float* xm = xch + (b * 2) * 128;        // [2 halves][128] block max
    float* xl = xch + 512 + (b * 2) * 128;  // [2 halves][128] partial row sums
    const uint32_t t_s = tmem_base + (uint32_t)(b * 128 + c_lo) + lane_base;
    const uint32_t t_o = tmem_base + XTM_O + (uint32_t)(b * 64 + half * 32) + lane_base;
    const uint32_t p_row = sP + (uint32_t)(b * 2) * XSLAB + (uint32_t)row * 128u;
    const int nvalid = min(cph, a.Nk - c_lo);  // key columns of this half that exist
    uint32_t p_addr[CPH / 8];                  // this thread's 16-byte P units (slab, swizzled unit): constant across items
#pragma unroll
    for (int q = 0; q < CPH / 8; ++q) {
      const int cc = (c_lo >> 3) + q;  // unit index along the key axis: slab cc >> 3, unit cc & 7
      p_addr[q] = p_row + (uint32_t)(cc >> 3) * XSLAB + (uint32_t)(((cc & 7) ^ (row & 7)) * 16);
    }
    auto epilogue = [&](int k) {  // store O(k) / l(k); runs one group-item late so that PV(k) has had time to finish
      const int item = item0 + k;
      const int bh = item / qtiles, qt = item - bh * qtiles;
      const int hb_b = bh / a.H, hb_h = bh - hb_b * a.H;
      asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");  // the partner's partial sum of item k is visible
      mbar_wait(o_full(b), (uint32_t)((k >> 1) & 1));
      tc_fence_after();
      const float inv_l = 1.f / (xl[row] + xl[128 + row]);
      uint32_t o[32];
      tmem_ld32(t_o, o);
      tc_wait_ld();
      tc_fence_before();
      const int q = qt * 128 + row;
      if (q < a.Nq) {
        TT* orow = reinterpret_cast<TT*>(a.out) + ((size_t)hb_b * a.Nq + q) * a.ldo + a.out_col0 + hb_h * a.dv + half * 32;
#pragma unroll
        for (int g = 0; g < 32; g += 8) {
          if (half * 32 + g + 8 <= a.dv) {
            uint4 u;
            u.x = T::pack(__uint_as_float(o[g + 0]) * inv_l, __uint_as_float(o[g + 1]) * inv_l);
            u.y = T::pack(__uint_as_float(o[g + 2]) * inv_l, __uint_as_float(o[g + 3]) * inv_l);
            u.z = T::pack(__uint_as_float(o[g + 4]) * inv_l, __uint_as_float(o[g + 5]) * inv_l);
            u.w = T::pack(__uint_as_float(o[g + 6]) * inv_l, __uint_as_float(o[g + 7]) * inv_l);
            *reinterpret_cast<uint4*>(orow + g) = u;
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(o_free(b));
    };
    int k_last = -1;
    for (int k = grp; k < n_items; k += 2) {
      if (k >= 2) epilogue(k - 2);
      k_last = k;
      mbar_wait(s_full(b), (uint32_t)((k >> 1) & 1));
      tc_fence_after();
      uint32_t r[CPH <= 48 ? 48 : 64];
      tmem_ld32(t_s, r);
      if (CPH <= 48) tmem_ld16(t_s + 32, r + 32);
      else tmem_ld32(t_s + 32, r + 32);
      tc_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_free(b));  // scores live in registers: the MMA warp may overwrite S_b
      if (nvalid < CPH) {  // warp-uniform: only the half that holds the ragged end of the keys masks anything
#pragma unroll
        for (int j = 0; j < CPH; ++j)
          if (j >= nvalid) r[j] = 0xff800000u;  // -inf: columns of the other half / beyond the keys
      }
      float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
      for (int j = 0; j < CPH; j += 8) {
        mx0 = fmaxf(fmaxf(mx0, __uint_as_float(r[j + 0])), __uint_as_float(r[j + 1]));
        mx1 = fmaxf(fmaxf(mx1, __uint_as_float(r[j + 2])), __uint_as_float(r[j + 3]));
        mx2 = fmaxf(fmaxf(mx2, __uint_as_float(r[j + 4])), __uint_as_float(r[j + 5]));
        mx3 = fmaxf(fmaxf(mx3, __uint_as_float(r[j + 6])), __uint_as_float(r[j + 7]));
      }
      float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
      xm[half * 128 + row] = mx;
      asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
      mx = fmaxf(mx, xm[(half ^ 1) * 128 + row]);
      const float mb = mx * sl2;
      float s0 = 0.f, s1 = 0.f;
#pragma unroll
      for (int q = 0; q < CPH / 8; ++q) {
        float p[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) p[j] = ex2_approx(fmaf(__uint_as_float(r[q * 8 + j]), sl2, -mb));
        s0 += (p[0] + p[1]) + (p[2] + p[3]);
        s1 += (p[4] + p[5]) + (p[6] + p[7]);
        if (q * 8 < cph) {
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(p_addr[q]), "r"(T::pack(p[0], p[1])),
                       "r"(T::pack(p[2], p[3])), "r"(T::pack(p[4], p[5])), "r"(T::pack(p[6], p[7]))
                       : "memory");
        }
      }
      xl[half * 128 + row] = s0 + s1;
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_ready(b));
    }
    if (k_last >= 0) epilogue(k_last);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

int attentionx_init() {
  static bool done = false;
  if (!done) {
    SDXE_CUDA_CHECK(cudaFuncSetAttribute(attentionx_kernel<true, 40>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    SDXE_CUDA_CHECK(cudaFuncSetAttribute(attentionx_kernel<false, 40>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    SDXE_CUDA_CHECK(cudaFuncSetAttribute(attentionx_kernel<true, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    SDXE_CUDA_CHECK(cudaFuncSetAttribute(attentionx_kernel<false, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    done = true;
  }
  return 0;
}

bool attentionx_eligible(const AttnArgs& a) {
  return a.dqk_slabs == 1 && a.dv_slabs == 1 && a.Nk >= 1 && a.Nk <= 128 && a.dv % 8 == 0 && a.dv <= 64 && a.out_col0 == 0;
}

int attentionx_launch(const AttnArgs& a, bool bf16, cudaStream_t stream) {
  const size_t smem = (size_t)(XSTG + 8) * XSLAB + 8 * 30 + 16 + 2 * 512 * 4 + 1024;
  if (attentionx_init() != 0) return -1;
  const int total = a.B * a.H * ((a.Nq + 127) / 128);
  if (total <= 0) return 0;
  const int grid = std::min(total, num_sms());
  const int cph = ((a.Nk + 15) / 16 * 16) / 2;
  if (cph <= 40) {
    auto kern = bf16 ? attentionx_kernel<true, 40> : attentionx_kernel<false, 40>;
    SDXE_CUDA_CHECK(launch_k(kern, dim3(grid), dim3(ATTX_THREADS), smem, stream, a));
  } else {
    auto kern = bf16 ? attentionx_kernel<true, 64> : attentionx_kernel<false, 64>;
    SDXE_CUDA_CHECK(launch_k(kern, dim3(grid), dim3(ATTX_THREADS), smem, stream, a));
  }
  return 0;
}

}  // namespace sdxe
