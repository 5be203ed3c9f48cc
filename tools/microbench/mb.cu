// Microbenchmarks of the per-SM resources that bound the flash-attention softmax on sm_100a (B200):
//   1. tcgen05.ld throughput (fp32 score tile out of TMEM), 4 / 8 / 16 warps, 1 or 4 loads in flight
//   2. tcgen05.st throughput
//   3. MUFU.EX2 throughput
//   4. exp2 on the FMA pipe (Cody-Waite range reduction + polynomial), accuracy and throughput
//   5. MUFU / FMA-pipe mixes (fraction f of the elements through the polynomial)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o mb mb.cu ; run on a B200: ./mb
#include "../../stable-diffusion-webui_b200/csrc/common.cuh"
#include <cmath>
#include <vector>

using namespace sdxe;

namespace sdxe {  // symbols common.cuh declares but this standalone binary does not link
void set_last_error(const char*, int, const char*) {}
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ long long clk() { long long c; asm volatile("mov.u64 %0, %%clock64;" : "=l"(c)); return c; }

// ------------------------------------------------------------------------------------------------ TMEM ld / st
template <int INFLIGHT, bool STORE>
__global__ void __launch_bounds__(512, 1) tmem_kernel(int iters, long long* cycles, uint32_t* sink) {
  __shared__ uint32_t tptr;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc(smem_u32(&tptr), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = tptr + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  uint32_t r[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) r[j] = threadIdx.x + j;
  __syncthreads();
  const long long t0 = clk();
  for (int it = 0; it < iters; ++it) {
    if (STORE) {
#pragma unroll
      for (int k = 0; k < INFLIGHT; ++k) tmem_st32(base + (uint32_t)(((warp >> 2) * INFLIGHT + k) * 32 & 511), r);
      tc_wait_st();
    } else {
      if (INFLIGHT == 1) {
        tmem_ld32(base + (uint32_t)((it & 15) * 32), r);
        tc_wait_ld();
#pragma unroll
        for (int j = 0; j < 32; ++j) acc ^= r[j];
      } else {
        uint32_t q[INFLIGHT][32];
#pragma unroll
        for (int k = 0; k < INFLIGHT; ++k) tmem_ld32(base + (uint32_t)(((it + k) & 15) * 32), q[k]);
        tc_wait_ld();
#pragma unroll
        for (int k = 0; k < INFLIGHT; ++k)
#pragma unroll
          for (int j = 0; j < 32; ++j) acc ^= q[k][j];
      }
    }
  }
  __syncthreads();
  const long long t1 = clk();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  if (acc == 0x12345678u) sink[0] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tptr, 512); }
}

// ------------------------------------------------------------------------------------------------ exp2 variants
// exp2 on the FMA pipe: n = round(x), r = x - n in [-0.5, 0.5], 2^r by a degree-DEG polynomial, 2^n through the exponent field.
template <int DEG>
__device__ __forceinline__ float exp2_poly(float x) {
  x = fmaxf(x, -125.f);
  const float magic = 12582912.f;  // 1.5 * 2^23: adding it rounds x to the nearest integer in the low mantissa bits
  const float t = x + magic;
  const float r = x - (t - magic);
  float p;
  if (DEG == 3) {  // minimax (relative) on [-0.5, 0.5]: max rel err 7.5e-5
    p = fmaf(0.0551716685295105f, r, 0.2426111400127411f);
    p = fmaf(p, r, 0.6932609677314758f);
    p = fmaf(p, r, 0.9999280571937561f);
  } else {         // degree 4: max rel err 2.7e-6
    p = fmaf(0.009570101276040077f, r, 0.05591786280274391f);
    p = fmaf(p, r, 0.240247443318367f);
    p = fmaf(p, r, 0.6931217908859253f);
    p = fmaf(p, r, 0.9999992847442627f);
  }
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}

template <int MODE>  // 0: all MUFU; 1: all poly3; 2: all poly4; 3: 1 of 4 poly3; 4: 2 of 4 poly3; 5: 1 of 4 poly4; 6: 2 of 4 poly4; 7: 1 of 8 poly3
__global__ void __launch_bounds__(512, 1) exp_kernel(int iters, long long* cycles, float* sink, float x0) {
  float v[8], s[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { v[j] = x0 - 0.01f * (threadIdx.x & 31) - j; s[j] = 0.f; }
  __syncthreads();
  const long long t0 = clk();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float x = fmaf(v[j], 0.999f, -0.001f * it);  // the scale/subtract FFMA the softmax does anyway
      float e;
      bool poly = false;
      int deg = 3;
      if (MODE == 1) poly = true;
      if (MODE == 2) { poly = true; deg = 4; }
      if (MODE == 3) poly = (j & 3) == 3;
      if (MODE == 4) poly = (j & 1) == 1;
      if (MODE == 5) { poly = (j & 3) == 3; deg = 4; }
      if (MODE == 6) { poly = (j & 1) == 1; deg = 4; }
      if (MODE == 7) poly = (j & 7) == 7;
      if (poly) e = deg == 3 ? exp2_poly<3>(x) : exp2_poly<4>(x);
      else e = ex2_approx(x);
      s[j] += e;
    }
  }
  __syncthreads();
  const long long t1 = clk();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  float tot = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) tot += s[j];
  if (tot == 1234.5f) sink[0] = tot;
}

__global__ void acc_kernel(const float* x, float* y3, float* y4, float* ym, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { y3[i] = exp2_poly<3>(x[i]); y4[i] = exp2_poly<4>(x[i]); ym[i] = ex2_approx(x[i]); }
}


// ------------------------------------------------------------------------------------------------ sync primitive costs
// one warp, clock64 around N repetitions of each primitive (issue + completion latency as seen by the issuing warp)
__global__ void __launch_bounds__(128, 1) sync_kernel(long long* out) {
  __shared__ __align__(8) uint64_t bars[8];
  __shared__ uint32_t tptr;
  __shared__ __align__(16) uint32_t buf[256];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) mbar_init(smem_u32(&bars[i]), 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(smem_u32(&tptr), 32);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp != 0) return;
  const uint32_t b0 = smem_u32(&bars[0]), b1 = smem_u32(&bars[1]);
  const int N = 64;
  long long t0, t1;
  // complete phase 0 of bar0 so that waits on parity 0 succeed at once
  if (lane == 0) mbar_arrive(b0);
  __syncwarp();
  // (0) try_wait on a completed phase, all 32 lanes
  t0 = clk();
  for (int i = 0; i < N; ++i) mbar_wait(b0, 0);
  t1 = clk();
  if (lane == 0) out[0] = (t1 - t0) / N;
  // (1) same, lane 0 only + __syncwarp
  t0 = clk();
  for (int i = 0; i < N; ++i) { if (lane == 0) mbar_wait(b0, 0); __syncwarp(); }
  t1 = clk();
  if (lane == 0) out[1] = (t1 - t0) / N;
  // (2) test_wait, all lanes + any_sync
  t0 = clk();
  int acc = 0;
  for (int i = 0; i < N; ++i) acc += __any_sync(0xffffffffu, mbar_test(b0, 0));
  t1 = clk();
  if (lane == 0) out[2] = (t1 - t0) / N + (acc == 12345);
  // (3) tcgen05.commit (no MMA outstanding), elected lane, then wait for the arrival (round trip)
  t0 = clk();
  for (int i = 0; i < N; ++i) {
    if (elect_one()) tc_commit(b1);
    __syncwarp();
    mbar_wait(b1, (uint32_t)(i & 1));
  }
  t1 = clk();
  if (lane == 0) out[3] = (t1 - t0) / N;
  // (4) tcgen05.commit issue only (arrivals drain in the background; phases flip freely)
  t0 = clk();
  for (int i = 0; i < N; ++i) {
    if (elect_one()) tc_commit(b1);
    __syncwarp();
  }
  t1 = clk();
  if (lane == 0) out[4] = (t1 - t0) / N;
  // (5) mbarrier.arrive by lane 0
  t0 = clk();
  for (int i = 0; i < N; ++i) { if (lane == 0) mbar_arrive(smem_u32(&bars[2])); __syncwarp(); }
  t1 = clk();
  if (lane == 0) out[5] = (t1 - t0) / N;
  // (6) fence.proxy.async after a shared store
  t0 = clk();
  for (int i = 0; i < N; ++i) { buf[lane] = i; fence_proxy_async_smem(); }
  t1 = clk();
  if (lane == 0) out[6] = (t1 - t0) / N;
  // (7) tcgen05 fence before + after
  t0 = clk();
  for (int i = 0; i < N; ++i) { tc_fence_before(); tc_fence_after(); }
  t1 = clk();
  if (lane == 0) out[7] = (t1 - t0) / N;
  // (8) elect_one + syncwarp alone
  t0 = clk();
  for (int i = 0; i < N; ++i) { if (elect_one()) buf[0] = i; __syncwarp(); }
  t1 = clk();
  if (lane == 0) out[8] = (t1 - t0) / N;
  // (9) bar.sync of 64 threads is measured elsewhere; here: st.shared.v4 x8 per lane (the P row) 
  t0 = clk();
  for (int i = 0; i < N; ++i) {
#pragma unroll
    for (int c = 0; c < 2; ++c)
      asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(smem_u32(&buf[0]) + (uint32_t)(((lane + c) & 15) * 16)), "r"(i) : "memory");
  }
  t1 = clk();
  if (lane == 0) out[9] = (t1 - t0) / N;
  // (10) clock64 read + global store (the trace stamp itself)
  t0 = clk();
  for (int i = 0; i < N; ++i) { if (lane == 0) out[16 + (i & 7)] = clk(); }
  t1 = clk();
  if (lane == 0) out[10] = (t1 - t0) / N;
  __syncwarp();
  tc_fence_before();
  tmem_dealloc(tptr, 32);
}

template <typename K, typename... A>
static double run(K kern, int threads, int iters, A... args) {
  long long* d;
  cudaMalloc(&d, 148 * sizeof(long long));
  kern<<<148, threads>>>(iters, d, args...);
  cudaDeviceSynchronize();
  kern<<<148, threads>>>(iters, d, args...);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); return -1; }
  std::vector<long long> h(148);
  cudaMemcpy(h.data(), d, 148 * sizeof(long long), cudaMemcpyDeviceToHost);
  cudaFree(d);
  double s = 0;
  for (auto v : h) s += (double)v;
  return s / 148.0;
}

int main() {
  uint32_t* sink;
  CK(cudaMalloc(&sink, 64));
  {
    long long* d;
    CK(cudaMalloc(&d, 64 * sizeof(long long)));
    sync_kernel<<<1, 128>>>(d);
    CK(cudaDeviceSynchronize());
    sync_kernel<<<1, 128>>>(d);
    CK(cudaDeviceSynchronize());
    long long h[16];
    CK(cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost));
    const char* nm[11] = {"mbarrier.try_wait (done phase), 32 lanes", "mbarrier.try_wait (done phase), lane 0 + syncwarp", "mbarrier.test_wait + any_sync, 32 lanes",
                          "tcgen05.commit -> arrival seen (round trip)", "tcgen05.commit issue only", "mbarrier.arrive lane 0 + syncwarp",
                          "st.shared + fence.proxy.async", "tcgen05.fence before+after", "elect_one + st.shared + syncwarp", "2 x st.shared.v4", "clock64 + st.global (trace stamp)"};
    printf("== sync primitive costs (clk per operation, one warp)\n");
    for (int i = 0; i < 11; ++i) printf("  %-52s %lld\n", nm[i], h[i]);
  }
  const int iters = 2000;
  printf("== tcgen05.ld 32x32b.x32 (4096 B per warp-instruction), grid 148 x 1 CTA/SM\n");
  for (int warps : {4, 8, 16}) {
    double c1 = run(tmem_kernel<1, false>, warps * 32, iters, sink);
    double c2 = run(tmem_kernel<2, false>, warps * 32, iters, sink);
    double c3 = run(tmem_kernel<3, false>, warps * 32, iters, sink);
    printf("  %2d warps: 1 in flight %.1f B/clk/SM   2 in flight %.1f B/clk/SM   3 in flight %.1f B/clk/SM\n", warps,
           warps * 4096.0 * iters / c1, warps * 4096.0 * 2 * iters / c2, warps * 4096.0 * 3 * iters / c3);
  }
  printf("== tcgen05.st 32x32b.x32\n");
  for (int warps : {4, 8, 16}) {
    double c1 = run(tmem_kernel<1, true>, warps * 32, iters, sink);
    double c2 = run(tmem_kernel<2, true>, warps * 32, iters, sink);
    printf("  %2d warps: 1 in flight %.1f B/clk/SM   2 in flight %.1f B/clk/SM\n", warps, warps * 4096.0 * iters / c1,
           warps * 4096.0 * 2 * iters / c2);
  }
  printf("== exp2 throughput (elements / clk / SM; each element also pays 1 FFMA + 1 FADD)\n");
  float* fs = reinterpret_cast<float*>(sink);
  const char* names[8] = {"all MUFU", "all poly3", "all poly4", "1/4 poly3", "1/2 poly3", "1/4 poly4", "1/2 poly4", "1/8 poly3"};
  for (int warps : {4, 8, 16}) {
    double c[8];
    c[0] = run(exp_kernel<0>, warps * 32, iters, fs, -1.f);
    c[1] = run(exp_kernel<1>, warps * 32, iters, fs, -1.f);
    c[2] = run(exp_kernel<2>, warps * 32, iters, fs, -1.f);
    c[3] = run(exp_kernel<3>, warps * 32, iters, fs, -1.f);
    c[4] = run(exp_kernel<4>, warps * 32, iters, fs, -1.f);
    c[5] = run(exp_kernel<5>, warps * 32, iters, fs, -1.f);
    c[6] = run(exp_kernel<6>, warps * 32, iters, fs, -1.f);
    c[7] = run(exp_kernel<7>, warps * 32, iters, fs, -1.f);
    printf("  %2d warps:", warps);
    for (int m = 0; m < 8; ++m) printf("  %s %.1f", names[m], warps * 32.0 * 8 * iters / c[m]);
    printf("\n");
  }
  // accuracy
  {
    const int n = 1 << 20;
    std::vector<float> hx(n), h3(n), h4(n), hm(n);
    for (int i = 0; i < n; ++i) hx[i] = -30.f + 38.f * (float)i / n;  // [-30, 8)
    float *dx, *d3, *d4, *dm;
    CK(cudaMalloc(&dx, n * 4)); CK(cudaMalloc(&d3, n * 4)); CK(cudaMalloc(&d4, n * 4)); CK(cudaMalloc(&dm, n * 4));
    CK(cudaMemcpy(dx, hx.data(), n * 4, cudaMemcpyHostToDevice));
    acc_kernel<<<n / 256, 256>>>(dx, d3, d4, dm, n);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h3.data(), d3, n * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(h4.data(), d4, n * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hm.data(), dm, n * 4, cudaMemcpyDeviceToHost));
    double e3 = 0, e4 = 0, em = 0;
    for (int i = 0; i < n; ++i) {
      const double ref = std::exp2((double)hx[i]);
      e3 = std::max(e3, std::fabs(h3[i] - ref) / ref);
      e4 = std::max(e4, std::fabs(h4[i] - ref) / ref);
      em = std::max(em, std::fabs(hm[i] - ref) / ref);
    }
    printf("== exp2 max relative error on [-30, 8): poly3 %.3e  poly4 %.3e  MUFU ex2.approx %.3e\n", e3, e4, em);
  }
  return 0;
}
