// sm_100a device primitives shared by every kernel of the denoising engine:
// mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (MMA / TMEM alloc / ld / st / commit),
// UMMA shared-memory + instruction descriptors, and 16-bit pack helpers.
// Everything here is inline PTX; nothing is borrowed from a library at run time.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <cstdio>
#include <cstring>

namespace sdxe {

// ---------------------------------------------------------------------------------------------
// dtype tags (match include/sdxe.h)
// ---------------------------------------------------------------------------------------------
enum : int { DT_F16 = 0, DT_BF16 = 1, DT_F32 = 2 };

#define SDXE_DEVINL __device__ __forceinline__

SDXE_DEVINL uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
SDXE_DEVINL uint32_t lane_id() { return threadIdx.x & 31; }

SDXE_DEVINL bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
      "elect.sync rx|px, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, px;\n\t}\n"
      : "=r"(pred));
  return pred != 0;
}

// ---------------------------------------------------------------------------------------------
// mbarrier
// ---------------------------------------------------------------------------------------------
SDXE_DEVINL void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
SDXE_DEVINL void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
SDXE_DEVINL void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
SDXE_DEVINL void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
SDXE_DEVINL uint64_t global_timer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// Wait for the phase with the given parity to complete. A watchdog turns a protocol bug
// (wrong phase, missing arrive, bad TMA descriptor) into a trap instead of a hung GPU.
SDXE_DEVINL void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  uint64_t t0 = 0;
  for (uint32_t it = 0;; ++it) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) return;
    if ((it & 0x3ff) == 0x3ff) {
      uint64_t now = global_timer_ns();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000ull) {  // 4 s
        printf("sdxe: mbarrier watchdog block(%d,%d,%d) thread %d bar 0x%x parity %u\n", blockIdx.x,
               blockIdx.y, blockIdx.z, threadIdx.x, bar, parity);
        __trap();
      }
    }
  }
}

// Programmatic dependent launch (PDL). A kernel launched through launch_k() may start while its predecessor in the
// stream is still draining: everything before pdl_wait() (smem carve-up, mbarrier init, TMEM allocation, tensor-map
// prefetch) overlaps the predecessor's tail; pdl_wait() returns once the predecessor grid has completed and its writes
// are visible, so NO global memory may be read or written before it. pdl_launch_dependents() lets the successor
// begin launching once every CTA of this grid has started. Both are no-ops for a normally launched kernel.
SDXE_DEVINL void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
SDXE_DEVINL void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// Non-blocking probe of a phase (for a consumer that serves several producers in arrival order).
SDXE_DEVINL bool mbar_test(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}\n"
      : "=r"(done)
      : "r"(bar), "r"(parity)
      : "memory");
  return done != 0;
}

// generic-proxy smem writes -> visible to async proxy (TMA / tcgen05.mma operand reads)
SDXE_DEVINL void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------
// TMA loads (tile mode), completion on an mbarrier.
// ---------------------------------------------------------------------------------------------
SDXE_DEVINL void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
SDXE_DEVINL void tma_load_2d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
SDXE_DEVINL void tma_load_3d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
SDXE_DEVINL void tma_load_4d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
SDXE_DEVINL void tma_load_5d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}

// TMA store smem -> global (bulk async-group completion). The issuing thread commits and later waits on its own groups.
SDXE_DEVINL void tma_store_2d(const CUtensorMap* m, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(src), "r"(c0), "r"(c1)
               : "memory");
}
SDXE_DEVINL void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
SDXE_DEVINL void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }  // smem reusable
SDXE_DEVINL void bulk_wait_read_1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }    // all but the newest group
SDXE_DEVINL void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }            // writes done

// multicast variant: the box lands at the same smem offset in every CTA of `mask`, each CTA's mbarrier (same offset)
// receives the complete_tx for the bytes written into it
SDXE_DEVINL void tma_load_2d_mc(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%4, %5}], [%2], %3;"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "h"(mask), "r"(c0), "r"(c1)
      : "memory");
}
// ---- CTA-pair (cta_group::2) forms. The mbarrier operand of a pair TMA load is a shared::cluster address whose
// "peer bit" (bit 24) selects the CTA of the pair; clearing it makes both CTAs' loads signal the LEADER's barrier
// (cute/arch/copy_sm100_tma.hpp: Sm100MmaPeerBitMask).
static constexpr uint32_t PEER_BIT_MASK = 0xFEFFFFFFu;
SDXE_DEVINL void tma2_load_2d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar & PEER_BIT_MASK), "r"(c0), "r"(c1)
      : "memory");
}
SDXE_DEVINL void tma2_load_4d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar & PEER_BIT_MASK), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
// arrive on the barrier at the same smem offset in CTA `cta` of the cluster
SDXE_DEVINL void mbar_arrive_remote(uint32_t bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}\n"
      ::"r"(bar), "r"(cta)
      : "memory");
}
SDXE_DEVINL uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
SDXE_DEVINL void cluster_sync_all() {  // every thread of every CTA of the cluster
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// ---------------------------------------------------------------------------------------------
// tcgen05: TMEM allocation, fences, commit, MMA, ld/st
// ---------------------------------------------------------------------------------------------
SDXE_DEVINL void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {  // whole warp, ncols pow2 >= 32
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
SDXE_DEVINL void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // whole warp (the allocating one)
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
SDXE_DEVINL void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
SDXE_DEVINL void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// All previously issued tcgen05.mma of this thread arrive on `bar` when complete.
SDXE_DEVINL void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// same, arriving on the barrier at this offset in every CTA of `mask` (smem slot shared through TMA multicast)
SDXE_DEVINL void tc_commit_mc(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask) : "memory");
}
// ---- cta_group::2: one MMA spans the CTA pair (M = 256: 128 accumulator rows in each CTA's TMEM; A from each CTA's
// own smem, B rows split half / half across the two CTAs' smem). Issued by the leader CTA only.
SDXE_DEVINL void tmem_alloc2(uint32_t smem_dst, uint32_t ncols) {  // same warp id in both CTAs
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
SDXE_DEVINL void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
SDXE_DEVINL void tc2_commit_mc(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask) : "memory");
}
SDXE_DEVINL void tc2_mma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], 16-bit inputs, fp32 accumulate.
SDXE_DEVINL void tc_mma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// same with the A operand read from tensor memory (TS form): A = [128 rows (lanes) x 16 k] 16-bit, two elements per 32-bit
// column (k even in the low half), 8 columns per K = 16 step
SDXE_DEVINL void tc_mma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
      ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
SDXE_DEVINL void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
SDXE_DEVINL void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// 32 lanes x 32 columns of fp32: thread i of the warp receives row (lane base + i), 32 consecutive columns.
SDXE_DEVINL void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
SDXE_DEVINL void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
SDXE_DEVINL void tmem_st32(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
      "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
        "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
        "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}

// ---------------------------------------------------------------------------------------------
// UMMA descriptors (bit layouts: cute/arch/mma_sm100_desc.hpp, restated).
//
// Shared-memory matrix descriptor (64 bit):
//   [0,14)  start address >> 4          [16,30) leading-dim byte offset >> 4
//   [32,46) stride-dim byte offset >> 4 [46,48) version = 1 (Blackwell)
//   [49,52) base offset = 0             [61,64) layout: 0 none, 2 = 128B swizzle, 4 = 64B, 6 = 32B
//
// K-major, 128B swizzle, 16-bit elements (our A/B tiles: rows of 64 elements = 128 bytes, as TMA
// writes them): 8-row groups are 1024 B apart -> SBO = 1024; LBO unused (1).
// MN-major, 128B swizzle (our V slabs: [kv rows][64 dv] with 128-byte rows): along K 8-row groups
// are 1024 B apart -> SBO = 1024; LBO = distance between 64-element MN chunks (slab size).
// ---------------------------------------------------------------------------------------------
SDXE_DEVINL uint64_t umma_desc_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;  // version
  d |= (uint64_t)2 << 61;  // SWIZZLE_128B
  return d;
}
// Instruction descriptor for kind::f16 (fp16 or bf16 inputs, fp32 accumulate).
//   [4,6) c fmt (1 = f32)  [7,10) a fmt  [10,13) b fmt (0 = f16, 1 = bf16)
//   [15] a major (0 = K)   [16] b major (0 = K, 1 = MN)   [17,23) N >> 3   [24,29) M >> 4
__host__ __device__ inline uint32_t umma_idesc(int bf16, int M, int N, int a_mn_major, int b_mn_major) {
  uint32_t d = 0;
  d |= 1u << 4;
  d |= (uint32_t)(bf16 ? 1 : 0) << 7;
  d |= (uint32_t)(bf16 ? 1 : 0) << 10;
  d |= (uint32_t)(a_mn_major ? 1 : 0) << 15;
  d |= (uint32_t)(b_mn_major ? 1 : 0) << 16;
  d |= (uint32_t)(N >> 3) << 17;
  d |= (uint32_t)(M >> 4) << 24;
  return d;
}

// ---------------------------------------------------------------------------------------------
// 16-bit conversions, templated on the storage type
// ---------------------------------------------------------------------------------------------
template <bool BF16> struct T16;
template <> struct T16<false> {
  using type = __half;
  using type2 = __half2;
  static SDXE_DEVINL uint32_t pack(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
  static SDXE_DEVINL float2 unpack(uint32_t u) { return __half22float2(*reinterpret_cast<__half2*>(&u)); }
  static SDXE_DEVINL float to_f(type v) { return __half2float(v); }
  static SDXE_DEVINL type from_f(float v) { return __float2half_rn(v); }
};
template <> struct T16<true> {
  using type = __nv_bfloat16;
  using type2 = __nv_bfloat162;
  static SDXE_DEVINL uint32_t pack(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
  static SDXE_DEVINL float2 unpack(uint32_t u) { return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&u)); }
  static SDXE_DEVINL float to_f(type v) { return __bfloat162float(v); }
  static SDXE_DEVINL type from_f(float v) { return __float2bfloat16_rn(v); }
};

SDXE_DEVINL float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// 2^x on the FMA pipe (Cody-Waite: n = round(x) through the 1.5 * 2^23 magic add, r = x - n in [-0.5, 0.5], degree-3
// minimax polynomial for 2^r, 2^n through the exponent field). Max relative error 7.5e-5 — below half an ulp of fp16
// (2.4e-4) and far below bf16's (2e-3): used for a fraction of the softmax exponentials so that the MUFU pipe
// (16 ex2 / clk / SM) is not the only unit doing them (the FA-4 trick). x <= ~100; x -> -inf gives 2^-125 (~0).
SDXE_DEVINL float ex2_poly3(float x) {
  x = fmaxf(x, -125.f);
  const float magic = 12582912.f;
  const float t = x + magic;
  const float r = x - (t - magic);
  float p = fmaf(0.0551716685295105f, r, 0.2426111400127411f);
  p = fmaf(p, r, 0.6932609677314758f);
  p = fmaf(p, r, 0.9999280571937561f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}
SDXE_DEVINL float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
SDXE_DEVINL float silu_f(float x) { return x * rcp_approx(1.f + ex2_approx(-1.4426950408889634f * x)); }
SDXE_DEVINL float gelu_erf_f(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752f)); }
// erf-GELU with erf from Abramowitz-Stegun 7.1.26 (|abs err| < 1.5e-7, far below 16-bit output resolution):
// one MUFU.EX2, one MUFU.RCP and a handful of FMAs instead of libdevice erff's branchy ~30 instructions.
// With erf(|z|) = 1 - P(t) e^{-z^2}, t = 1 / (1 + p |z|), z = x / sqrt 2:
//   gelu(x) = x/2 (1 + erf z) = max(x, 0) - |x|/2 P(t) e^{-z^2}        (both signs of x)
// and with u = x sqrt(log2(e) / 2) (so that e^{-z^2} = 2^{-u^2}) the constants |z| / |u| and |x| / (2 |u|) fold into p and
// into P's coefficients: 11 FP32-pipe instructions + 2 MUFU per element (the textbook arrangement below took 15 + 2, and
// the GEGLU epilogue is instruction-bound: profiles/r1_notes.md finding 3).
#ifndef SDXE_GELU_V1
SDXE_DEVINL float gelu_fast_f(float x) {
  constexpr float C = 0.84932180028801904f;        // sqrt(log2(e) / 2)
  constexpr float S = 0.83255461115769776f;        // |z| / |u| = 1 / sqrt(log2(e))
  constexpr float H = 0.58870501125773735f;        // |x| / (2 |u|) = 1 / (2 C)
  const float u = x * C;
  const float a = fabsf(u);
  const float t = rcp_approx(fmaf(0.3275911f * S, a, 1.f));
  float p = fmaf(1.061405429f * H, t, -1.453152027f * H);
  p = fmaf(p, t, 1.421413741f * H);
  p = fmaf(p, t, -0.284496736f * H);
  p = fmaf(p, t, 0.254829592f * H);
  const float e = ex2_approx(-u * u);
  return fmaf(-(a * e), p * t, fmaxf(x, 0.f));
}
#else
SDXE_DEVINL float gelu_fast_f(float x) {
  const float z = fabsf(x) * 0.70710678118654752f;
  const float t = rcp_approx(fmaf(0.3275911f, z, 1.f));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  const float e = ex2_approx(-z * z * 1.4426950408889634f);
  const float erf_abs = fmaf(-p * t, e, 1.f);          // erf(|x|/sqrt2)
  const float erfv = copysignf(erf_abs, x);
  return 0.5f * x * (1.f + erfv);
}
#endif

// ---------------------------------------------------------------------------------------------
// host: error handling + tensor-map encoding through the driver entry point (no -lcuda link)
// ---------------------------------------------------------------------------------------------
#define SDXE_CUDA_CHECK(expr)                                                         \
  do {                                                                                \
    cudaError_t _e = (expr);                                                          \
    if (_e != cudaSuccess) {                                                          \
      ::sdxe::set_last_error(__FILE__, __LINE__, cudaGetErrorString(_e));             \
      return -1;                                                                      \
    }                                                                                 \
  } while (0)

void set_last_error(const char* file, int line, const char* msg);
const char* last_error();

// 16-bit row-major 2D [rows, cols] with row pitch ld (elements): box = 64 cols x box_rows, 128B swizzle.
int make_tmap_2d(CUtensorMap* out, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows);
// same with an explicit box; swizzle = the box's row bytes (32 / 64 / 128 B) — the GEMM epilogue's store / residual boxes
int make_tmap_2d_box(CUtensorMap* out, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_cols, int box_rows);
// per-head view of a row-major activation: element (j, tok, h, b) at base + b*batch_stride + tok*tok_stride + h*head_stride
// + j (strides in elements, multiples of 8). Inner extent d (NOT padded): a 64-wide box past d is zero-filled by TMA.
int make_tmap_heads(CUtensorMap* out, const void* base, int64_t d, int64_t tokens, int64_t heads, int64_t batch,
                    int64_t tok_stride, int64_t head_stride, int64_t batch_stride, int box_rows);
// 16-bit 3D [d2, d1 rows, d0 cols] with pitches: box = 64 x box_rows x 1.
int make_tmap_3d(CUtensorMap* out, const void* base, int64_t d0, int64_t d1, int64_t d2, int64_t pitch1, int64_t pitch2,
                 int box_rows);
// 16-bit NHWC activations [N, H, W, C]: box = 64 ch x bw x bh x bn, 128B swizzle, OOB -> zero (= conv padding).
int make_tmap_nhwc(CUtensorMap* out, const void* base, int N, int H, int W, int C, int bw, int bh, int bn);
int make_tmap_nhwc_s2(CUtensorMap* out, const void* base, int N, int H, int W, int C, int bw, int bh, int bn);

int num_sms();
bool pdl_enabled();  // SDXE_PDL=1 turns programmatic dependent launch on (default off: measured 1-2 % slower)

// Launch with the programmatic-stream-serialization attribute (see pdl_wait above). Only for kernels that call
// pdl_wait() before touching global memory.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

}  // namespace sdxe
