// tcgen05 GEMM / implicit-GEMM 3x3 convolution for sm_100a.
//
//   out[M, N] = A[M, K] * W[N, K]^T  (+ bias + per-sample vector + residual | GEGLU)
//
// Replaces, on the UNet / VAE hot path of the reference (all PyTorch library dispatch there):
//   * ResBlock / Upsample / VAE conv3x3      (ldm openaimodel.py ResBlock) -> conv mode
//   * conv1x1 proj_in/proj_out/skip, Linear q/k/v/out/FF                          -> plain mode
//   * GEGLU (ldm attention.py GEGLU: x * gelu(gate))                                         -> EPI_GEGLU
//   * nn.LayerNorm in front of q/k/v, cross-attention q and the feed-forward (BasicTransformerBlock norm1-3)
//     -> folded: the producing GEMM emits per-row (sum, sum of squares) (STAT), the consuming GEMM normalises in its
//        epilogue (LNF) with gamma / beta pre-multiplied into its packed weight / bias (gemm.cuh)
//
// Structure (one CTA per SM, persistent over output tiles of 128 x BN):
//   warp 0   : TMA producer. A tile = 128 rows x 64 K (16 KB, 128B swizzle). In conv mode the A tile is a
//              4D NHWC box (64 ch x W x bh x bn) fetched at the tap's (dx-1, dy-1) offset: TMA's out-of-bounds
//              zero fill IS the convolution padding, so no im2col buffer ever exists.
//   warp 1   : tcgen05.mma issuer (single thread), fp32 accumulators in TMEM, double buffered (2 x BN columns)
//   warps 2-9: epilogue (two warps per TMEM lane quarter, alternating 32-column chunks). tcgen05.ld the accumulator
//              (thread = row), fuse bias / timestep-embedding vector / residual / GEGLU, convert to 16 bit, write the
//              32 x 32 chunk into a swizzled staging buffer and TMA-store it; the residual chunk is TMA-loaded into
//              the same buffer one chunk ahead and updated in place. Bias staged in smem per tile.
#include "gemm.cuh"
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#ifndef SDXE_GEMM_TRACE
#define SDXE_GEMM_TRACE 0
#endif

namespace sdxe {

static constexpr int BLOCK_M = 128;
static constexpr int BLOCK_K = 64;
static constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;
static constexpr int GEMM_THREADS = 320;  // warp 0 TMA, warp 1 MMA, warps 2-9 epilogue (two per TMEM lane quarter)
// GEGLU variants: 16 epilogue warps (four per lane quarter). Their epilogue is instruction-bound — ~24 issue slots per
// output element (two accumulators, bias / LayerNorm fold on both, erf-GELU with two MUFU ops) — and with eight warps
// only two were resident per scheduler: issue slots 53 % used, 6.2 k cycles per 128 x 256 tile against 2.5 k of tensor
// work at K = 320 (profiles/r2_ncu_geglu.txt). Blocks above 512 threads are held to 96 registers, so these warps work
// in 16-column chunks (32 live accumulator registers instead of 64).
static constexpr int GEGLU_EPI_WARPS = 16;
static constexpr int GEGLU_THREADS = 64 + 32 * GEGLU_EPI_WARPS;
static constexpr int GEGLU_CHUNK_BYTES = 32 * 16 * 2;  // 32 rows x 16 columns, 16 bit (same staging bytes per CTA as 8 x 2 KB)
static constexpr int TMEM_COLS = 512;
static constexpr int SBIAS_BYTES = 4 * 256 * 4;  // per-tile bias and LayerNorm-fold c1 slices in smem, double buffered
static constexpr int CHUNK_BYTES = 32 * 32 * 2;   // one epilogue chunk: 32 rows x 32 columns, 16 bit
static constexpr int NUM_EPI_WARPS = 8;
static constexpr int NUM_BARS_FIXED = 4 + 3 * GEGLU_EPI_WARPS;  // tfull[2], tempty[2], residual-landed[warp][<= 3 buffers]

// EPI / RES / ROWVEC are compile-time so that the epilogue's inner loop carries no mode branches (it was spending
// two thirds of its instructions on flag tests and parameter reloads, and the epilogue bounds every small-K GEMM).
template <bool BF16, int EPI, bool RES, bool ROWVEC, bool CLUSTER, bool LNF, bool STAT>
__global__ void __launch_bounds__(EPI == EPI_GEGLU ? GEGLU_THREADS : GEMM_THREADS, 1) gemm_kernel(const __grid_constant__ GemmArgs a) {
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t pad = ((raw + 1023u) & ~1023u) - raw;
  uint8_t* smem = smem_raw + pad;
  const uint32_t smem_base = raw + pad;

  const int S = a.num_stages;
  const int BN = a.BN;
  // paired (cta_group::2): each CTA keeps its own A tile and HALF of the B tile
  const uint32_t stage_bytes = A_STAGE_BYTES + (uint32_t)(CLUSTER ? (BN >> 1) : BN) * 128u;
  const uint32_t bar_base = smem_base + S * stage_bytes;  // 8-byte aligned (stage_bytes % 1024 == 0)
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (S + s); };
  auto tfull_bar = [&](int i) { return bar_base + 8u * (2 * S + i); };
  auto tempty_bar = [&](int i) { return bar_base + 8u * (2 * S + 2 + i); };
  auto res_bar = [&](int w, int b) { return bar_base + 8u * (2 * S + 4 + 4 * w + b); };  // <= 4 residual buffers per warp
  const uint32_t misc_off = S * stage_bytes + 8 * (2 * S + NUM_BARS_FIXED);
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + misc_off);
  float* sbias = reinterpret_cast<float*>(smem + misc_off + 16);  // [2][256]
  // staging chunks per epilogue warp: one more than strictly needed (plain 2, residual 3) lets chunk c be written while
  // the TMA store of chunk c-1 is still reading its buffer (the store's smem read latency otherwise serialises the
  // chunks: measured ~5.5 k cycles per 128 x 160 tile of pure epilogue); falls back to 1 / 2 when smem is short
  const int NBUF = a.epi_bufs;
  const bool deep = NBUF > (RES ? 2 : 1);
  const uint32_t stage_buf_base = (smem_base + misc_off + 16 + SBIAS_BYTES + 1023u) & ~1023u;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
#if SDXE_GEMM_TRACE
  // timeline of CTA 0: [role 0 producer | 1 MMA | 2 first epilogue warp | 3 last epilogue warp |
  // 4, 5 first epilogue warp inside its first chunk: ld issued, ld done, arithmetic done, buffer free | stored, fenced, TMA issued][tile < 40][event < 4]
  const bool tracing = a.trace != nullptr && blockIdx.x == 0;
  auto TR = [&](int role, int it, int ev) {
    if (tracing && lane == 0 && it < 40) {
      unsigned long long c;
      asm volatile("mov.u64 %0, %%clock64;" : "=l"(c));
      a.trace[(role * 40 + it) * 4 + ev] = c;
    }
  };
#else
#define TR(role, it, ev) ((void)0)
#endif

  const int num_m = (a.M + BLOCK_M - 1) / BLOCK_M;
  const int num_n = (a.N + BN - 1) / BN;
  const int num_kb = (a.K + BLOCK_K - 1) / BLOCK_K;
  // Work enumeration. Unpaired: tile t -> (m_blk, n_blk) = (t / num_n, t % num_n), CTA stride gridDim.x.
  // Paired (CLUSTER, cta_group::2): the CTA pair walks 256-row tile pairs p -> m_blk = 2 * (p / num_n) + rank,
  // n_blk = p % num_n. Both CTAs load (own A rows, own half of B) into their own smem and signal the LEADER's full
  // barrier; the leader's MMA lane issues M = 256 MMAs that read both CTAs' smem and write both CTAs' TMEM; its
  // commits are multicast to both CTAs (smem slot free, accumulator ready); both epilogues drain their own TMEM and
  // arrive on the leader's accumulator-free barrier. Shared-memory traffic per CTA drops by the half B tile, which
  // is what bounds the single-CTA 128 x BN tile (profiles/r1_notes.md, finding 5).
  constexpr bool clustered = CLUSTER;  // compile-time: the default (unpaired) kernel carries none of this
  const uint32_t crank = clustered ? cluster_ctarank() : 0u;
  const int work_first = clustered ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int work_step = clustered ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int num_tiles = clustered ? (num_m >> 1) * num_n : num_m * num_n;  // work items per CTA walk

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(full_bar(s), clustered ? 2 : 1);  // paired: leader's expect_tx arrive + peer's remote arrive
      mbar_init(empty_bar(s), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(tfull_bar(i), 1);
      // every epilogue warp frees the accumulator (paired: both CTAs' warps free the leader's)
      mbar_init(tempty_bar(i), (EPI == EPI_GEGLU ? GEGLU_EPI_WARPS : NUM_EPI_WARPS) * (clustered ? 2 : 1));
    }
    if (RES)
      for (int w = 0; w < NUM_EPI_WARPS; ++w)
        for (int b = 0; b < 4; ++b) mbar_init(res_bar(w, b), 1);
    fence_mbar_init();
    tma_prefetch_desc(&a.tmA);
    tma_prefetch_desc(&a.tmB);
  }
  if (warp == 1) {
    if (clustered) tmem_alloc2(smem_u32(tmem_ptr_smem), TMEM_COLS);
    else tmem_alloc(smem_u32(tmem_ptr_smem), TMEM_COLS);
  }
  tc_fence_before();
  __syncthreads();
  if (clustered) cluster_sync_all();  // peer barriers initialised before any multicast / remote arrive can reach them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_wait();  // the producing kernel's activations (A, residual) are complete; our output buffer is no longer read

  // Producer and MMA warps run their loops CONVERGED (all 32 lanes wait on the barriers) and only the asynchronous
  // issue itself is predicated on one elected lane. Issuing from inside a divergent `if (lane == 0)` region makes the
  // compiler wrap every uniform-datapath instruction (UTMALDG / UTCHMMA / UTCBAR) in an ELECT + BRA.U.ANY loop with
  // R2UR conversions: ~70 dependent instructions (~700 cycles) per k-block on the single lane that feeds the tensor
  // pipe — measured as the limiter of every tile narrower than 256 (tensor pipe 50 % busy at BN = 160).
  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = work_first; tile < num_tiles; tile += work_step) {
      const int n_blk = tile % num_n;
      const int m_blk = clustered ? 2 * (tile / num_n) + (int)crank : tile / num_n;
      const int m0 = m_blk * BLOCK_M;
      int img0 = 0, h0 = 0, w0 = 0;
      if (a.conv) {
        const int hw = a.H * a.W;
        img0 = m0 / hw;
        const int rem = m0 - img0 * hw;
        h0 = rem / a.W;
        w0 = rem - h0 * a.W;
      }
      int tap = 0, cb = 0;  // conv: k-block -> (tap, channel block) without divisions
      TR(0, (tile - work_first) / work_step, 0);
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(empty_bar(stage), phase ^ 1u);
        if (kb == 0) TR(0, (tile - work_first) / work_step, 1);
        if (elect_one()) {
          const uint32_t sA = smem_base + stage * stage_bytes;
          const uint32_t sB = sA + A_STAGE_BYTES;
          const uint32_t fb = full_bar(stage);
          if (clustered) {
            // every byte of both CTAs is accounted on the leader's barrier
            if (crank == 0) mbar_expect_tx(fb, 2u * stage_bytes);
            else mbar_arrive_remote(fb, 0u);
            if (a.conv) {
              const int dy = tap / 3, dx = tap - dy * 3;
              tma2_load_4d(sA, &a.tmA, fb, cb * BLOCK_K, w0 + dx - 1, h0 + dy - 1, img0);
            } else {
              const int k0 = kb * BLOCK_K;
              if (k0 < a.K1) tma2_load_2d(sA, &a.tmA, fb, k0, m0);
              else tma2_load_2d(sA, &a.tmA2, fb, k0 - a.K1, m0);
            }
            tma2_load_2d(sB, &a.tmB, fb, kb * BLOCK_K, n_blk * BN + (int)crank * (BN >> 1));  // my half of B
          } else {
            mbar_expect_tx(fb, stage_bytes);
            if (a.conv == 2) {  // stride 2: (pixel pair, parity) addressing of the 5-D view, see make_tmap_nhwc_s2
              const int dy = tap / 3, dx = tap - dy * 3;
              const int ty = dy - a.pad_lo, tx = dx - a.pad_lo;
              tma_load_5d(sA, &a.tmA, fb, (tx & 1) * (a.cblocks * BLOCK_K) + cb * BLOCK_K, w0 + (tx >> 1), ty & 1, h0 + (ty >> 1), img0);
            } else if (a.conv) {
              const int dy = tap / 3, dx = tap - dy * 3;
              tma_load_4d(sA, &a.tmA, fb, cb * BLOCK_K, w0 + dx - 1, h0 + dy - 1, img0);
            } else {
              const int k0 = kb * BLOCK_K;
              if (k0 < a.K1) tma_load_2d(sA, &a.tmA, fb, k0, m0);
              else tma_load_2d(sA, &a.tmA2, fb, k0 - a.K1, m0);
            }
            tma_load_2d(sB, &a.tmB, fb, kb * BLOCK_K, n_blk * BN);
          }
        }
        __syncwarp();
        if (++cb == a.cblocks) { cb = 0; ++tap; }
        if (++stage == S) { stage = 0; phase ^= 1u; }
      }
      TR(0, (tile - work_first) / work_step, 2);
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (!clustered || crank == 0) {  // paired: only the leader CTA issues
      const uint32_t idesc = umma_idesc(BF16 ? 1 : 0, clustered ? 2 * BLOCK_M : BLOCK_M, BN, 0, 0);
      // descriptors of stage 0; stage s adds s * stage_bytes to the (address >> 4) field
      const uint64_t adesc0 = umma_desc_sw128(smem_base, 16, 1024);
      const uint64_t bdesc0 = umma_desc_sw128(smem_base + A_STAGE_BYTES, 16, 1024);
      const uint32_t stage_inc = stage_bytes >> 4;
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = work_first; tile < num_tiles; tile += work_step) {
        TR(1, (tile - work_first) / work_step, 0);
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        tc_fence_after();
        TR(1, (tile - work_first) / work_step, 1);
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          if (kb == 0) TR(1, (tile - work_first) / work_step, 2);
          if (elect_one()) {
            const uint64_t adesc = adesc0 + (uint64_t)(stage * stage_inc);
            const uint64_t bdesc = bdesc0 + (uint64_t)(stage * stage_inc);
            // +32 bytes along K inside the 128B swizzle atom = +2 in the (addr >> 4) field
            if (clustered) {
              tc2_mma_f16(d_tmem, adesc, bdesc, idesc, kb != 0 ? 1u : 0u);
              tc2_mma_f16(d_tmem, adesc + 2, bdesc + 2, idesc, 1u);
              tc2_mma_f16(d_tmem, adesc + 4, bdesc + 4, idesc, 1u);
              tc2_mma_f16(d_tmem, adesc + 6, bdesc + 6, idesc, 1u);
              tc2_commit_mc(empty_bar(stage), (uint16_t)3);  // both CTAs' slot `stage` is free
            } else {
              tc_mma_f16(d_tmem, adesc, bdesc, idesc, kb != 0 ? 1u : 0u);
              tc_mma_f16(d_tmem, adesc + 2, bdesc + 2, idesc, 1u);
              tc_mma_f16(d_tmem, adesc + 4, bdesc + 4, idesc, 1u);
              tc_mma_f16(d_tmem, adesc + 6, bdesc + 6, idesc, 1u);
              tc_commit(empty_bar(stage));
            }
          }
          __syncwarp();
          if (++stage == S) { stage = 0; phase ^= 1u; }
        }
        if (elect_one()) {
          if (clustered) tc2_commit_mc(tfull_bar(acc), (uint16_t)3);  // both CTAs' accumulator halves are ready
          else tc_commit(tfull_bar(acc));
        }
        __syncwarp();
        TR(1, (tile - work_first) / work_step, 3);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
  } else if constexpr (EPI == EPI_GEGLU) {
    // ------------------------------------------------------------------ epilogue warps, GEGLU: out = value * gelu(gate)
    // Sixteen warps: warp w may touch TMEM lanes 32 (w % 4) .. +31, the four warps of a lane quarter take every fourth
    // 16-column chunk. Value and gate accumulators of a chunk are read together (columns c and BN / 2 + c), the product
    // is staged in a 32B-swizzled 1 KB buffer and leaves through a TMA store (same path as the plain epilogue below).
    const int quarter = warp & 3;
    const int ew = warp - 2;             // 0..15
    const int esub = ew >> 2;            // 0..3
    const int row = quarter * 32 + lane;
    const int et = threadIdx.x - 64;
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t tile_ctr = 0;
    const int half = BN >> 1;            // output columns per tile; the gate accumulators start at column `half`
    const float* const biasp = a.bias;
    const int M = a.M;
    const uint32_t my_buf = stage_buf_base + (uint32_t)(ew * NBUF) * GEGLU_CHUNK_BYTES;
    uint32_t chunk_ctr = 0;
    for (int tile = work_first; tile < num_tiles; tile += work_step, ++tile_ctr) {
      const int n_blk = tile % num_n;
      const int m_blk = clustered ? 2 * (tile / num_n) + (int)crank : tile / num_n;
      const int m = m_blk * BLOCK_M + row;
      const int row0 = m_blk * BLOCK_M + quarter * 32;
      const int n_out0 = n_blk * half;
      float* sb = sbias + (tile_ctr & 1u) * 256;
      float* sc = sbias + 512 + (tile_ctr & 1u) * 256;
      for (int j = et; j < BN; j += 32 * GEGLU_EPI_WARPS) {
        // PACK_GEGLU weights / bias / c1: tile n_blk's rows are [half value rows | the matching half gate rows]; N % BN == 0
        sb[j] = biasp != nullptr ? __ldg(biasp + n_blk * BN + j) : 0.f;
        if (LNF) sc[j] = __ldg(a.c1 + n_blk * BN + j);
      }
      float ln_mu = 0.f, ln_rs = 1.f;
      if (LNF) {
        float ps = 0.f, pq = 0.f;
        if (m < M) {
          for (int pp = 0; pp < a.ln_parts; ++pp) {
            const float2 t = __ldg(a.ln_part + (size_t)pp * M + m);
            ps += t.x;
            pq += t.y;
          }
        }
        ln_mu = ps * a.ln_inv_c;
        ln_rs = rsqrtf(fmaxf(pq * a.ln_inv_c - ln_mu * ln_mu, 0.f) + a.ln_eps);
      }
      asm volatile("bar.sync 1, %0;" ::"n"(32 * GEGLU_EPI_WARPS) : "memory");  // bias of this tile visible to all epilogue warps
      if (ew == 0 || ew == GEGLU_EPI_WARPS - 1) TR(ew == 0 ? 2 : 3, (int)tile_ctr, 0);
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      if (ew == 0 || ew == GEGLU_EPI_WARPS - 1) TR(ew == 0 ? 2 : 3, (int)tile_ctr, 1);
      const uint32_t t_row = tmem_base + (uint32_t)(acc * BN) + ((uint32_t)(quarter * 32) << 16);
      for (int c0 = esub * 16; c0 < half; c0 += 64, ++chunk_ctr) {
        const uint32_t buf = my_buf + (chunk_ctr % (uint32_t)NBUF) * GEGLU_CHUNK_BYTES;
        uint32_t r[16], rg[16];
        tmem_ld16(t_row + c0, r);
        tmem_ld16(t_row + half + c0, rg);
        tc_wait_ld();
        uint4 packed[2];
#pragma unroll
        for (int g = 0; g < 16; g += 8) {
          float v[8], gt[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) { v[j] = __uint_as_float(r[g + j]); gt[j] = __uint_as_float(rg[g + j]); }
          const float4 b0 = *reinterpret_cast<const float4*>(sb + c0 + g), b1 = *reinterpret_cast<const float4*>(sb + c0 + g + 4);
          const float4 d0 = *reinterpret_cast<const float4*>(sb + half + c0 + g), d1 = *reinterpret_cast<const float4*>(sb + half + c0 + g + 4);
          if (LNF) {
            const float4 k0 = *reinterpret_cast<const float4*>(sc + c0 + g), k1 = *reinterpret_cast<const float4*>(sc + c0 + g + 4);
            const float4 q0 = *reinterpret_cast<const float4*>(sc + half + c0 + g), q1 = *reinterpret_cast<const float4*>(sc + half + c0 + g + 4);
            v[0] = fmaf(ln_rs, v[0] - ln_mu * k0.x, b0.x); v[1] = fmaf(ln_rs, v[1] - ln_mu * k0.y, b0.y);
            v[2] = fmaf(ln_rs, v[2] - ln_mu * k0.z, b0.z); v[3] = fmaf(ln_rs, v[3] - ln_mu * k0.w, b0.w);
            v[4] = fmaf(ln_rs, v[4] - ln_mu * k1.x, b1.x); v[5] = fmaf(ln_rs, v[5] - ln_mu * k1.y, b1.y);
            v[6] = fmaf(ln_rs, v[6] - ln_mu * k1.z, b1.z); v[7] = fmaf(ln_rs, v[7] - ln_mu * k1.w, b1.w);
            gt[0] = fmaf(ln_rs, gt[0] - ln_mu * q0.x, d0.x); gt[1] = fmaf(ln_rs, gt[1] - ln_mu * q0.y, d0.y);
            gt[2] = fmaf(ln_rs, gt[2] - ln_mu * q0.z, d0.z); gt[3] = fmaf(ln_rs, gt[3] - ln_mu * q0.w, d0.w);
            gt[4] = fmaf(ln_rs, gt[4] - ln_mu * q1.x, d1.x); gt[5] = fmaf(ln_rs, gt[5] - ln_mu * q1.y, d1.y);
            gt[6] = fmaf(ln_rs, gt[6] - ln_mu * q1.z, d1.z); gt[7] = fmaf(ln_rs, gt[7] - ln_mu * q1.w, d1.w);
          } else {
            v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
            v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
            gt[0] += d0.x; gt[1] += d0.y; gt[2] += d0.z; gt[3] += d0.w;
            gt[4] += d1.x; gt[5] += d1.y; gt[6] += d1.z; gt[7] += d1.w;
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] *= gelu_fast_f(gt[j]);
          packed[g >> 3].x = T16<BF16>::pack(v[0], v[1]);
          packed[g >> 3].y = T16<BF16>::pack(v[2], v[3]);
          packed[g >> 3].z = T16<BF16>::pack(v[4], v[5]);
          packed[g >> 3].w = T16<BF16>::pack(v[6], v[7]);
        }
        // the buffer was last read by the store of chunk c - NBUF: that read must be over before it is overwritten
        if (elect_one()) { if (deep) bulk_wait_read_1(); else bulk_wait_read_all(); }
        __syncwarp();
        const uint32_t lin0 = (uint32_t)lane * 32u;  // 32-byte rows, 32B swizzle: 16-byte unit index ^= address bit 7
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          const uint32_t lin = lin0 + (uint32_t)(g * 16);
          const uint4 u = packed[g];
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};"
                       ::"r"(buf + (lin ^ (((lin >> 7) & 1u) << 4))), "r"(u.x), "r"(u.y), "r"(u.z), "r"(u.w)
                       : "memory");
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (elect_one()) {
          tma_store_2d(&a.tmO16, buf, n_out0 + c0, row0);
          bulk_commit_group();
        }
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (ew == 0 || ew == GEGLU_EPI_WARPS - 1) TR(ew == 0 ? 2 : 3, (int)tile_ctr, 2);
      if (lane == 0) {
        if (clustered && crank != 0) mbar_arrive_remote(tempty_bar(acc), 0u);
        else mbar_arrive(tempty_bar(acc));
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
    if (elect_one()) bulk_wait_all();
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue warps
    // Row-per-thread (the only way tcgen05.ld hands out data): a thread owns 64 contiguous bytes of one output row per
    // chunk, so direct global accesses would touch 32 different 128-byte lines per warp instruction (measured: the
    // store path alone held every small-K GEMM at ~1.5 TB/s). The chunk is therefore staged in shared memory (TMA
    // 64B / 32B swizzle = conflict-free 16-byte row pieces) and moved by TMA in both directions; out-of-range rows and
    // columns are clipped (stores) or zero-filled (residual) by the tensor maps, so the loop carries no guards.
    const int quarter = warp & 3;        // TMEM lane quarter this warp may access (warps w and w+4 share it)
    const int ew = warp - 2;             // 0..7
    const int ehalf = ew >> 2;           // the two warps of a quarter take alternate 32-column chunks
    const int row = quarter * 32 + lane;
    const int et = threadIdx.x - 64;     // 0..255 among the epilogue threads
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t tile_ctr = 0;
    uint32_t chunk_ctr = 0;              // chunks this warp has handled: staging buffer / barrier parity
    const int out_cols = BN;             // output columns produced per tile
    const float* const biasp = a.bias;
    const int M = a.M, N = a.N;
    const uint32_t my_buf = stage_buf_base + (uint32_t)(ew * NBUF) * CHUNK_BYTES;
    // RES: this warp's residual chunks form one stream across its tiles; a prefetch cursor runs a.res_dist chunks ahead
    // of the chunk being processed (also across tile boundaries), each chunk TMA-loaded into the staging buffer it will
    // be updated in and stored from. (Distance 1 left ~0.6 k cycles of residual latency in every chunk's chain and the
    // whole latency at each tile start: profiles/r2_gemm_epilogue_trace.txt.)
    int p_tile = work_first, p_c0 = ehalf * 32;
    uint32_t p_ctr = 0;
    auto res_prefetch = [&]() {  // elected lane only
      if (p_tile >= num_tiles || p_c0 >= out_cols) return;
      const int pn = p_tile % num_n;
      const int pm = clustered ? 2 * (p_tile / num_n) + (int)crank : p_tile / num_n;
      const uint32_t pb = p_ctr % (uint32_t)NBUF;
      const int nc = min(32, out_cols - p_c0);
      mbar_expect_tx(res_bar(ew, pb), (uint32_t)nc * 64u);
      tma_load_2d(my_buf + pb * CHUNK_BYTES, nc == 32 ? &a.tmR : &a.tmR16, res_bar(ew, pb), pn * out_cols + p_c0, pm * BLOCK_M + quarter * 32);
      ++p_ctr;
      p_c0 += 64;
      if (p_c0 >= out_cols) { p_c0 = ehalf * 32; p_tile += work_step; }
    };
    if (RES) {
      if (elect_one())
        for (int i = 0; i < a.res_dist; ++i) res_prefetch();  // buffers are fresh: nothing to wait for
      __syncwarp();
    }
    for (int tile = work_first; tile < num_tiles; tile += work_step, ++tile_ctr) {
      const int n_blk = tile % num_n;
      const int m_blk = clustered ? 2 * (tile / num_n) + (int)crank : tile / num_n;
      const int m = m_blk * BLOCK_M + row;
      const int row0 = m_blk * BLOCK_M + quarter * 32;  // first row of this warp's band
      const bool row_ok = m < M;
      const int n_out0 = n_blk * out_cols;  // first output column of this tile
      float* sb = sbias + (tile_ctr & 1u) * 256;
      float* sc = sbias + 512 + (tile_ctr & 1u) * 256;  // LayerNorm fold: c1 of this tile's columns
      for (int j = et; j < BN; j += 256) {
        const bool in = n_blk * BN + j < N;
        sb[j] = (biasp != nullptr && in) ? __ldg(biasp + n_blk * BN + j) : 0.f;
        if (LNF) sc[j] = in ? __ldg(a.c1 + n_blk * BN + j) : 0.f;
      }
      float ln_mu = 0.f, ln_rs = 1.f;
      if (LNF) {  // this row's mean / rstd from the producer's partial sums (fixed order: deterministic)
        float ps = 0.f, pq = 0.f;
        if (row_ok) {
          for (int pp = 0; pp < a.ln_parts; ++pp) {
            const float2 t = __ldg(a.ln_part + (size_t)pp * M + m);
            ps += t.x;
            pq += t.y;
          }
        }
        ln_mu = ps * a.ln_inv_c;
        ln_rs = rsqrtf(fmaxf(pq * a.ln_inv_c - ln_mu * ln_mu, 0.f) + a.ln_eps);
      }
      float st_s = 0.f, st_q = 0.f;  // STAT: this thread's share of its row's (sum, sum of squares)
      asm volatile("bar.sync 1, 256;" ::: "memory");  // bias of this tile visible to all epilogue warps
      if (ew == 0 || ew == NUM_EPI_WARPS - 1) TR(ew == 0 ? 2 : 3, (int)tile_ctr, 0);
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      if (ew == 0 || ew == NUM_EPI_WARPS - 1) TR(ew == 0 ? 2 : 3, (int)tile_ctr, 1);
      const uint32_t t_row = tmem_base + (uint32_t)(acc * BN) + ((uint32_t)(quarter * 32) << 16);
      const float* rv = nullptr;
      if (ROWVEC && row_ok) rv = a.rowvec + (size_t)(m / a.rows_per_sample) * a.ldrv + n_out0;
      const bool cols_full = n_out0 + out_cols <= N;

      for (int c0 = ehalf * 32; c0 < out_cols; c0 += 64, ++chunk_ctr) {
        const int nc = min(32, out_cols - c0);
        const uint32_t b = chunk_ctr % (uint32_t)NBUF;
        const uint32_t buf = my_buf + b * CHUNK_BYTES;
        uint32_t r[32];
        const bool trc = ew == 0 && c0 == 0;  // (trace builds) stamps inside the first warp's first chunk
        if (trc) TR(4, (int)tile_ctr, 0);
        if (nc == 32) tmem_ld32(t_row + c0, r);
        else tmem_ld16(t_row + c0, r);
        if (RES) {
          // keep the residual stream res_dist chunks ahead: the buffer of chunk c + res_dist was last read by the store of
          // chunk c + res_dist - NBUF, which may have one younger store still in flight (deep) or none
          if (elect_one()) {
            if (deep) bulk_wait_read_1(); else bulk_wait_read_all();
            res_prefetch();
          }
          __syncwarp();
          mbar_wait(res_bar(ew, b), (chunk_ctr / (uint32_t)NBUF) & 1u);
        }
        tc_wait_ld();
        if (trc) TR(4, (int)tile_ctr, 1);
        // this thread's row piece inside the chunk: 16-byte unit g at (lane * rowbytes + 16 g) ^ swizzle
        const uint32_t lin0 = (uint32_t)lane * (uint32_t)(nc * 2);
        const uint32_t swz_mask = nc == 32 ? 3u : 1u;  // 64B / 32B swizzle: unit index ^= address bits [7, 8] / [7]
        uint4 packed[4];
#pragma unroll
        for (int g = 0; g < 32; g += 8) {
          if (g < nc) {
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[g + j]);
            {
              const float4 b0 = *reinterpret_cast<const float4*>(sb + c0 + g), b1 = *reinterpret_cast<const float4*>(sb + c0 + g + 4);
              if (LNF) {
                const float4 k0 = *reinterpret_cast<const float4*>(sc + c0 + g), k1 = *reinterpret_cast<const float4*>(sc + c0 + g + 4);
                v[0] = fmaf(ln_rs, v[0] - ln_mu * k0.x, b0.x); v[1] = fmaf(ln_rs, v[1] - ln_mu * k0.y, b0.y);
                v[2] = fmaf(ln_rs, v[2] - ln_mu * k0.z, b0.z); v[3] = fmaf(ln_rs, v[3] - ln_mu * k0.w, b0.w);
                v[4] = fmaf(ln_rs, v[4] - ln_mu * k1.x, b1.x); v[5] = fmaf(ln_rs, v[5] - ln_mu * k1.y, b1.y);
                v[6] = fmaf(ln_rs, v[6] - ln_mu * k1.z, b1.z); v[7] = fmaf(ln_rs, v[7] - ln_mu * k1.w, b1.w);
              } else {
                v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
                v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
              }
            }
            if (ROWVEC) {
              if (rv) {
                if (cols_full || n_out0 + c0 + g + 8 <= N) {
                  const float4 e0 = __ldg(reinterpret_cast<const float4*>(rv + c0 + g)), e1 = __ldg(reinterpret_cast<const float4*>(rv + c0 + g + 4));
                  v[0] += e0.x; v[1] += e0.y; v[2] += e0.z; v[3] += e0.w;
                  v[4] += e1.x; v[5] += e1.y; v[6] += e1.z; v[7] += e1.w;
                } else {
#pragma unroll
                  for (int j = 0; j < 8; ++j)
                    if (n_out0 + c0 + g + j < N) v[j] += __ldg(rv + c0 + g + j);
                }
              }
            }
            if (RES) {
              const uint32_t lin = lin0 + (uint32_t)(g * 2);
              uint4 u;
              asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                           : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w)
                           : "r"(buf + (lin ^ (((lin >> 7) & swz_mask) << 4))));
              float2 f;
              f = T16<BF16>::unpack(u.x); v[0] += f.x; v[1] += f.y;
              f = T16<BF16>::unpack(u.y); v[2] += f.x; v[3] += f.y;
              f = T16<BF16>::unpack(u.z); v[4] += f.x; v[5] += f.y;
              f = T16<BF16>::unpack(u.w); v[6] += f.x; v[7] += f.y;
            }
            packed[g >> 3].x = T16<BF16>::pack(v[0], v[1]);
            packed[g >> 3].y = T16<BF16>::pack(v[2], v[3]);
            packed[g >> 3].z = T16<BF16>::pack(v[4], v[5]);
            packed[g >> 3].w = T16<BF16>::pack(v[6], v[7]);
            if (STAT) {  // statistics of the values as stored (rounded to 16 bit), like a LayerNorm reading them back
              const uint4 pk = packed[g >> 3];
              float2 f;
              f = T16<BF16>::unpack(pk.x); st_s += f.x + f.y; st_q = fmaf(f.x, f.x, fmaf(f.y, f.y, st_q));
              f = T16<BF16>::unpack(pk.y); st_s += f.x + f.y; st_q = fmaf(f.x, f.x, fmaf(f.y, f.y, st_q));
              f = T16<BF16>::unpack(pk.z); st_s += f.x + f.y; st_q = fmaf(f.x, f.x, fmaf(f.y, f.y, st_q));
              f = T16<BF16>::unpack(pk.w); st_s += f.x + f.y; st_q = fmaf(f.x, f.x, fmaf(f.y, f.y, st_q));
            }
          }
        }
        if (trc) TR(4, (int)tile_ctr, 2);
        if (!RES) {
          // this buffer was last read by the store of chunk c - NBUF: it must be done before the buffer is overwritten
          if (elect_one()) { if (deep) bulk_wait_read_1(); else bulk_wait_read_all(); }
          __syncwarp();
        }
        if (trc) TR(4, (int)tile_ctr, 3);
#pragma unroll
        for (int g = 0; g < 32; g += 8) {
          if (g < nc) {
            const uint32_t lin = lin0 + (uint32_t)(g * 2);
            const uint4 u = packed[g >> 3];
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};"
                         ::"r"(buf + (lin ^ (((lin >> 7) & swz_mask) << 4))), "r"(u.x), "r"(u.y), "r"(u.z), "r"(u.w)
                         : "memory");
          }
        }
        if (trc) TR(5, (int)tile_ctr, 0);
        fence_proxy_async_smem();
        __syncwarp();
        if (trc) TR(5, (int)tile_ctr, 1);
        if (elect_one()) {
          tma_store_2d(nc == 32 ? &a.tmO : &a.tmO16, buf, n_out0 + c0, row0);
          bulk_commit_group();
        }
        __syncwarp();
        if (trc) TR(5, (int)tile_ctr, 2);
      }
      if (STAT && row_ok) a.stat_out[(size_t)(n_blk * 2 + ehalf) * M + m] = make_float2(st_s, st_q);
      tc_fence_before();
      __syncwarp();
      if (ew == 0 || ew == NUM_EPI_WARPS - 1) TR(ew == 0 ? 2 : 3, (int)tile_ctr, 2);
      if (lane == 0) {
        if (clustered && crank != 0) mbar_arrive_remote(tempty_bar(acc), 0u);  // the leader's MMA lane owns the accumulators
        else mbar_arrive(tempty_bar(acc));
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
    if (elect_one()) bulk_wait_all();  // all of this warp's stores have reached global memory
    __syncwarp();
  }

  tc_fence_before();
  __syncthreads();
  if (clustered) cluster_sync_all();  // no CTA may exit while its peer can still multicast into it
  if (warp == 1) {
    tc_fence_after();
    if (clustered) tmem_dealloc2(tmem_base, TMEM_COLS);
    else tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

bool conv_tile_shape(int H, int W, int* bw, int* bh, int* bn) {
  if (W >= 128) {
    if (W % 128) return false;
    *bw = 128; *bh = 1; *bn = 1;
    return true;
  }
  if (128 % W) return false;
  const int rows = 128 / W;
  if (H >= rows) {
    if (H % rows) return false;
    *bw = W; *bh = rows; *bn = 1;
    return true;
  }
  if (rows % H) return false;
  *bw = W; *bh = H; *bn = rows / H;
  return true;
}

static size_t gemm_smem_fixed(int num_stages, int epi_bufs) {  // everything but the operand stages
  return 1024 /*base alignment*/ + 8 * (2 * num_stages + NUM_BARS_FIXED) + 16 + SBIAS_BYTES + 1024 /*staging alignment*/ +
         (size_t)NUM_EPI_WARPS * epi_bufs * CHUNK_BYTES;
}
int gemm_pick_stages(int BN, int epi_bufs) {
  const int stage_bytes = A_STAGE_BYTES + BN * 128;
  int s = (int)((227 * 1024 - gemm_smem_fixed(8, epi_bufs)) / stage_bytes);
  return std::max(2, std::min(s, 8));
}
int gemm_finish_args(GemmArgs& a, const void* W, int64_t w_rows, int64_t w_ld) {
  a.cluster = (a.c1 || a.stat_out || a.conv == 2) ? 1 : gemm_pick_cluster(a.M, a.BN);  // fold / statistics / stride-2 variants are un-paired only
  const int bn_cta = a.cluster == 2 ? a.BN / 2 : a.BN;
  // one staging buffer more than the minimum unless that costs an operand stage below four
  int min_bufs = a.residual ? 2 : 1;
  a.res_dist = 1;
  if (a.residual) {
    // residual prefetch distance 2 needs four staging buffers per epilogue warp (64 KB): taken when the operand ring keeps
    // at least min(k-blocks, 3) stages and the GEMM is short (K <= 640: epilogue-paced) or loses no stage to it
    static int dist2 = -1;
    if (dist2 < 0) { const char* e = getenv("SDXE_RES_DIST2"); dist2 = e ? atoi(e) : 1; }
    const int num_kb = (a.K + BLOCK_K - 1) / BLOCK_K;
    const int s4 = gemm_pick_stages(bn_cta, 4);
    if (dist2 && !a.conv && s4 >= std::min(num_kb, 3) && (num_kb <= 10 || s4 == gemm_pick_stages(bn_cta, 2))) { min_bufs = 4; a.res_dist = 2; }
  }
  const int s_min = gemm_pick_stages(bn_cta, min_bufs), s_deep = gemm_pick_stages(bn_cta, min_bufs + 1);
  static int deep_ok = -1;
  // default off: measured no gain (same-box A/B, SD1.5 UNet 18.74 / 18.76 ms with vs 18.65 / 18.98 ms without)
  if (deep_ok < 0) { const char* e = getenv("SDXE_EPI_DEEP"); deep_ok = e ? atoi(e) : 0; }
  a.epi_bufs = (min_bufs < 4 && deep_ok && (s_deep == s_min || s_deep >= 4)) ? min_bufs + 1 : min_bufs;
  a.num_stages = a.epi_bufs > min_bufs ? s_deep : s_min;
  if (make_tmap_2d(&a.tmB, W, w_rows, a.K, w_ld, bn_cta)) return -1;
  const int64_t out_cols = a.epi == EPI_GEGLU ? a.N / 2 : (a.N + 7) / 8 * 8;
  if (a.ldo % 8 || (a.residual && a.ldr % 8)) { set_last_error(__FILE__, __LINE__, "gemm: ldo / ldr must be multiples of 8"); return -1; }
  if (make_tmap_2d_box(&a.tmO, a.out, a.M, std::min<int64_t>(out_cols, a.ldo), a.ldo, 32, 32)) return -1;
  if (make_tmap_2d_box(&a.tmO16, a.out, a.M, std::min<int64_t>(out_cols, a.ldo), a.ldo, 16, 32)) return -1;
  if (a.residual) {
    if (make_tmap_2d_box(&a.tmR, a.residual, a.M, std::min<int64_t>(out_cols, a.ldr), a.ldr, 32, 32)) return -1;
    if (make_tmap_2d_box(&a.tmR16, a.residual, a.M, std::min<int64_t>(out_cols, a.ldr), a.ldr, 16, 32)) return -1;
  } else {
    a.tmR = a.tmO;
    a.tmR16 = a.tmO16;
  }
  return 0;
}

int gemm_pick_bn(int M, int N, int K, int epi) {
  (void)K;
  const int sms = num_sms();
  const int num_m = (M + BLOCK_M - 1) / BLOCK_M;
  double best = 1e30;
  int best_bn = 16;
  for (int bn = 16; bn <= 256; bn += 16) {
    if (epi == EPI_GEGLU && (bn % 32 != 0 || N % bn != 0)) continue;
    const int num_n = (N + bn - 1) / bn;
    const long tiles = (long)num_m * num_n;
    const long waves = (tiles + sms - 1) / sms;
    // cycles per 64-deep k-block: tensor pipe (2*BN) vs. operand fetch through L2 (~48 B/clk/SM) + fixed overhead
    const double kb = std::max(2.0 * bn, (A_STAGE_BYTES + 128.0 * bn) / 48.0) + 24.0;
    const double cost = waves * kb;
    if (cost < best - 1e-9) { best = cost; best_bn = bn; }
  }
  return best_bn;
}

typedef void (*GemmKernel)(const GemmArgs);
// variant index = cluster * 10 + bf16 * 5 + {0: plain, 1: plain+res, 2: plain+rowvec, 3: plain+res+rowvec, 4: geglu};
// un-paired only: 20 + bf16 * 4 + {0: plain+lnfold, 1: geglu+lnfold, 2: plain+stat, 3: plain+res+stat}
template <bool BF16, bool CL>
static GemmKernel gemm_variant_e(int e) {
  switch (e) {
    case 0: return gemm_kernel<BF16, EPI_PLAIN, false, false, CL, false, false>;
    case 1: return gemm_kernel<BF16, EPI_PLAIN, true, false, CL, false, false>;
    case 2: return gemm_kernel<BF16, EPI_PLAIN, false, true, CL, false, false>;
    case 3: return gemm_kernel<BF16, EPI_PLAIN, true, true, CL, false, false>;
    default: return gemm_kernel<BF16, EPI_GEGLU, false, false, CL, false, false>;
  }
}
template <bool BF16>
static GemmKernel gemm_variant_x(int e) {
  switch (e) {
    case 0: return gemm_kernel<BF16, EPI_PLAIN, false, false, false, true, false>;
    case 1: return gemm_kernel<BF16, EPI_GEGLU, false, false, false, true, false>;
    case 2: return gemm_kernel<BF16, EPI_PLAIN, false, false, false, false, true>;
    default: return gemm_kernel<BF16, EPI_PLAIN, true, false, false, false, true>;
  }
}
static constexpr int GEMM_NUM_VARIANTS = 28;
static GemmKernel gemm_variant(int i) {
  if (i >= 20) return (i - 20) / 4 ? gemm_variant_x<true>((i - 20) % 4) : gemm_variant_x<false>((i - 20) % 4);
  const int cl = i / 10, bf = (i % 10) / 5, e = i % 5;
  if (cl) return bf ? gemm_variant_e<true, true>(e) : gemm_variant_e<false, true>(e);
  return bf ? gemm_variant_e<true, false>(e) : gemm_variant_e<false, false>(e);
}

int gemm_init() {
  static bool done = false;
  if (!done) {
    for (int i = 0; i < GEMM_NUM_VARIANTS; ++i) SDXE_CUDA_CHECK(cudaFuncSetAttribute(gemm_variant(i), cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    done = true;
  }
  return 0;
}

#if SDXE_GEMM_TRACE
// SDXE_GEMM_TRACE_DUMP=<n>: the n-th gemm_launch of the process (1-based) runs with the timeline buffer and writes
// gpurun_out/gemm_trace.txt (tools/analyze_gemm_trace.py). Debug builds only (make GEMM_TRACE=1).
static int gemm_launch_inner(const GemmArgs& a, bool bf16, cudaStream_t stream);
int gemm_launch(const GemmArgs& a0, bool bf16, cudaStream_t stream) {
  static int want = -1, count = 0;
  static unsigned long long* buf = nullptr;
  if (want < 0) { const char* e = getenv("SDXE_GEMM_TRACE_DUMP"); want = e ? atoi(e) : 0; }
  if (!want || ++count != want) return gemm_launch_inner(a0, bf16, stream);
  GemmArgs a = a0;
  const size_t bytes = 6 * 40 * 4 * 8;
  if (!buf) cudaMalloc(&buf, bytes);
  cudaMemsetAsync(buf, 0, bytes, stream);
  a.trace = buf;
  const int rc = gemm_launch_inner(a, bf16, stream);
  cudaStreamSynchronize(stream);
  static unsigned long long h[6 * 40 * 4];
  cudaMemcpy(h, buf, bytes, cudaMemcpyDeviceToHost);
  FILE* f = fopen("gpurun_out/gemm_trace.txt", "w");
  if (f) {
    fprintf(f, "# M=%d N=%d K=%d BN=%d stages=%d epi=%d res=%d conv=%d\n", a.M, a.N, a.K, a.BN, a.num_stages, a.epi, a.residual ? 1 : 0, a.conv);
    for (int r = 0; r < 6; ++r)
      for (int t = 0; t < 40; ++t)
        fprintf(f, "%d %d %llu %llu %llu %llu\n", r, t, h[(r * 40 + t) * 4], h[(r * 40 + t) * 4 + 1], h[(r * 40 + t) * 4 + 2], h[(r * 40 + t) * 4 + 3]);
    fclose(f);
  }
  return rc;
}
static int gemm_launch_inner(const GemmArgs& a, bool bf16, cudaStream_t stream) {
#else
int gemm_launch(const GemmArgs& a, bool bf16, cudaStream_t stream) {
#endif
  if (a.BN % 16 != 0 || a.BN < 16 || a.BN > 256) { set_last_error(__FILE__, __LINE__, "gemm: bad BN"); return -1; }
  if (a.K1 != a.K && (a.K1 % BLOCK_K) != 0) { set_last_error(__FILE__, __LINE__, "gemm: K1 % 64"); return -1; }
  if (a.epi == EPI_GEGLU && (a.BN % 32 != 0 || a.N % a.BN != 0)) { set_last_error(__FILE__, __LINE__, "gemm: geglu tile"); return -1; }
  const int stage_bytes = A_STAGE_BYTES + (a.cluster == 2 ? a.BN / 2 : a.BN) * 128;
  if (a.epi_bufs < (a.residual ? 2 : 1) || a.epi_bufs > 4 || (a.residual && (a.res_dist < 1 || a.res_dist > a.epi_bufs - 1))) { set_last_error(__FILE__, __LINE__, "gemm: epi_bufs (call gemm_finish_args)"); return -1; }
  const size_t smem = (size_t)a.num_stages * stage_bytes + gemm_smem_fixed(a.num_stages, a.epi_bufs);
  if (smem > 227 * 1024) { set_last_error(__FILE__, __LINE__, "gemm: shared memory budget"); return -1; }
  const int num_m = (a.M + BLOCK_M - 1) / BLOCK_M, num_n = (a.N + a.BN - 1) / a.BN;
  const int tiles = num_m * num_n;
  if (tiles <= 0) return 0;
  int grid = std::min(tiles, num_sms());
  if (a.cluster == 2) {
    if ((num_m & 1) || (a.BN % 16)) { set_last_error(__FILE__, __LINE__, "gemm: cluster needs an even m-tile count"); return -1; }
    grid = std::min(tiles, num_sms() & ~1);
  }
  const int vi = a.epi == EPI_GEGLU ? 4 : ((a.residual ? 1 : 0) | (a.rowvec ? 2 : 0));
  if ((a.epi != EPI_PLAIN) && (a.residual || a.rowvec)) { set_last_error(__FILE__, __LINE__, "gemm: residual / rowvec need EPI_PLAIN"); return -1; }
  GemmKernel kern = gemm_variant(vi + (bf16 ? 5 : 0) + (a.cluster == 2 ? 10 : 0));
  if (a.c1 || a.stat_out) {
    if (a.cluster == 2 || a.rowvec || (a.c1 && (a.residual || a.stat_out)) || (a.stat_out && a.epi != EPI_PLAIN)) {
      set_last_error(__FILE__, __LINE__, "gemm: unsupported LayerNorm-fold / statistics combination");
      return -1;
    }
    const int xi = a.c1 ? (a.epi == EPI_GEGLU ? 1 : 0) : (a.residual ? 3 : 2);
    kern = gemm_variant(20 + (bf16 ? 4 : 0) + xi);
  }
  if (gemm_init() != 0) return -1;
  const int threads = a.epi == EPI_GEGLU ? GEGLU_THREADS : GEMM_THREADS;
  if (a.cluster == 2) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    SDXE_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, a));
  } else {
    SDXE_CUDA_CHECK(launch_k(kern, dim3(grid), dim3(threads), smem, stream, a));
  }
  SDXE_CUDA_CHECK(cudaGetLastError());
  return 0;
}

// Pair CTAs (cta_group::2, 256-row tile pairs) when the geometry allows it. History (profiles/r1_notes.md): plain TMA
// multicast of the B tile across the pair gave 0 % — the single-CTA tile is bound by shared-memory bandwidth, and
// multicast still lands the full B tile in both CTAs; splitting B across the pair's smem is what cta_group::2 buys.
int gemm_pick_cluster(int M, int BN) {
  static int mode = -1;
  if (mode < 0) { const char* e = getenv("SDXE_CLUSTER"); mode = e ? atoi(e) : 0; }
  if (!mode) return 1;
  const int num_m = (M + BLOCK_M - 1) / BLOCK_M;
  return (num_m % 2 == 0 && num_m >= 2 && BN % 16 == 0 && BN >= 32) ? 2 : 1;
}

}  // namespace sdxe
