// Micro-probes that decide the conv kernel design (run on the GPU box, results kept under profiles/):
//   1. shifted-window A operands from ONE halo tile in shared memory, SWIZZLE_NONE K-major layout
//      ([channel-group][y][x][8 ch]): does tcgen05.mma accept 16-byte-granular start addresses?
//   2. the same with SWIZZLE_128B rows (one pixel = one 128-byte row) and the descriptor's base_offset field
//   3. tcgen05.mma issue rate vs N for shared-memory operands (is N=32/64 shared-memory-read bound?)
//   4. L2 -> shared memory bandwidth per SM with cp.async (16 B per thread) and with bulk copies
//   next round (written, compiled, not yet run): `ws` — weights as the A operand (smem or TMEM) against pixel views as B;
//   `coll` — A-operand collector reuse across MMAs that share an input view; `wst` — correctness of the swapped roles
//   (weights M = 128 / 64 as A, 128 shifted pixels as B) incl. which TMEM lane holds which output channel
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I reliability-challenges-uncertainty_b200/csrc \
//        tools/ubench/umma_probe.cu -o tools/ubench/umma_probe
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "conv_tc.cuh"

using namespace rcu;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred;
}

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout, uint32_t base_off) {
  return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) |
         (uint64_t(1) << 46) | ((uint64_t)base_off << 49) | ((uint64_t)layout << 61);
}

// ------------------------------------------------------------------------------------------------ 1 + 2: correctness
// mode 0: SWIZZLE_NONE planes, C = 32, halo 18 x 10, N = 32
// mode 1: SWIZZLE_128B rows, C = 64, halo 18 x 16 (pitch 16), N = 32, base_offset = (start >> 7) & 7
// mode 2: as 1 with base_offset = 0
__global__ void __launch_bounds__(128, 1) shift_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ w,
                                                       float* __restrict__ out, int mode) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sp = smem_raw + (base - smem_u32(smem_raw));
  __shared__ uint64_t bar;
  __shared__ uint32_t s_tmem;
  // mode 3: SWIZZLE_128B rows, C = 64, DENSE halo 18 x 10 (pitch 10, what one TMA box writes), base_offset 0
  // mode 4: SWIZZLE_64B rows, C = 32, dense halo 18 x 10, base_offset 0
  const int C = (mode == 0 || mode == 4) ? 32 : 64;
  const int P = (mode == 1 || mode == 2) ? 16 : 10;
  const int N = 32;
  const int RB = C * 2;  // row bytes
  const uint32_t a_bytes = mode == 0 ? 4 * 18 * 10 * 16 : 18 * 16 * 128;
  const uint32_t sA = base, sB = (base + a_bytes + 1023u) & ~1023u;
  uint8_t* pA = sp;
  uint8_t* pB = sp + (sB - base);
  // ---- fill A
  for (int i = threadIdx.x; i < 18 * 10 * (C / 8); i += blockDim.x) {
    const int g = i % (C / 8), px = (i / (C / 8)) % 10, py = i / (C / 8) / 10;
    const uint4 v = *reinterpret_cast<const uint4*>(x + ((size_t)(py * 10 + px) * C + g * 8));
    uint32_t off;
    if (mode == 0) off = (uint32_t)g * (18 * 10 * 16) + (uint32_t)(py * P + px) * 16;
    else {
      const uint32_t row = (uint32_t)(py * P + px) * RB;             // byte address of the row (tile base is 1024-aligned)
      const uint32_t x7 = mode == 4 ? ((row >> 7) & 3u) : ((row >> 7) & 7u);  // SW64: bits[4:5] ^= bits[7:8]; SW128: bits[4:6] ^= bits[7:9]
      off = row + (uint32_t)((g ^ x7) * 16);
    }
    *reinterpret_cast<uint4*>(pA + off) = v;
  }
  // ---- fill B: w[tap][n][ci]
  for (int i = threadIdx.x; i < 9 * N * (C / 8); i += blockDim.x) {
    const int g = i % (C / 8), n = (i / (C / 8)) % N, tap = i / (C / 8) / N;
    const uint4 v = *reinterpret_cast<const uint4*>(w + ((size_t)(tap * N + n) * C + g * 8));
    uint32_t off;
    if (mode == 0) off = (uint32_t)(tap * (C / 8) + g) * (N * 16) + (uint32_t)n * 16;
    else {
      const uint32_t row = (uint32_t)tap * (N * RB) + (uint32_t)n * RB;
      const uint32_t x7 = mode == 4 ? ((row >> 7) & 3u) : ((row >> 7) & 7u);
      off = row + (uint32_t)((g ^ x7) * 16);
    }
    *reinterpret_cast<uint4*>(pB + off) = v;
  }
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  if (threadIdx.x < 32) { tmem_alloc(smem_u32(&s_tmem), 32); tmem_relinquish(); }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;
  if (threadIdx.x == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(N >> 3) << 17) | (uint32_t(128 >> 4) << 24);
    int first = 1;
    for (int tap = 0; tap < 9; ++tap) {
      const int ty = tap / 3, tx = tap % 3;
      for (int ks = 0; ks < C / 16; ++ks) {
        uint64_t da, db;
        if (mode == 0) {
          da = make_desc(sA + (uint32_t)(2 * ks) * (18 * 10 * 16) + (uint32_t)(ty * P + tx) * 16, 18 * 10 * 16, P * 16, 0, 0);
          db = make_desc(sB + (uint32_t)(tap * (C / 8) + 2 * ks) * (N * 16), N * 16, 128, 0, 0);
        } else {
          const uint32_t start = sA + (uint32_t)(ty * P + tx) * RB + (uint32_t)ks * 32;
          const uint32_t lay = mode == 4 ? 4u : 2u;
          da = make_desc(start, 16, P * RB, lay, mode == 1 ? ((start >> 7) & 7u) : 0u);
          db = make_desc(sB + (uint32_t)tap * (N * RB) + (uint32_t)ks * 32, 16, 8 * RB, lay, 0);
        }
        umma_bf16(tmem, da, db, idesc, first ? 0u : 1u);
        first = 0;
      }
    }
    umma_commit(smem_u32(&bar));
  }
  mbar_wait(smem_u32(&bar), 0);
  tc_fence_after();
  uint32_t v[32];
  const int warp = threadIdx.x >> 5;
  tmem_ld_32x32b_x32(tmem + ((uint32_t)(warp * 32) << 16), v);
  tmem_ld_wait();
  for (int c = 0; c < 32; ++c) out[threadIdx.x * 32 + c] = __uint_as_float(v[c]);
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 32);
}

static void run_shift(int mode) {
  const int C = (mode == 0 || mode == 4) ? 32 : 64, N = 32;
  std::vector<__nv_bfloat16> hx(18 * 10 * C), hw(9 * N * C);
  std::vector<float> fx(hx.size()), fw(hw.size());
  srand(1234 + mode);
  for (size_t i = 0; i < hx.size(); ++i) { fx[i] = (float)(rand() % 5 - 2); hx[i] = __float2bfloat16(fx[i]); }
  for (size_t i = 0; i < hw.size(); ++i) { fw[i] = (float)(rand() % 5 - 2); hw[i] = __float2bfloat16(fw[i]); }
  __nv_bfloat16 *dx, *dw;
  float* dout;
  CK(cudaMalloc(&dx, hx.size() * 2)); CK(cudaMalloc(&dw, hw.size() * 2)); CK(cudaMalloc(&dout, 128 * 32 * 4));
  CK(cudaMemcpy(dx, hx.data(), hx.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dw, hw.data(), hw.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(dout, 0, 128 * 32 * 4));
  const int smem = 110 * 1024;
  CK(cudaFuncSetAttribute(shift_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  shift_kernel<<<1, 128, smem>>>(dx, dw, dout, mode);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("SHIFT mode %d: kernel failed: %s\n", mode, cudaGetErrorString(e)); exit(2); }
  std::vector<float> got(128 * 32);
  CK(cudaMemcpy(got.data(), dout, got.size() * 4, cudaMemcpyDeviceToHost));
  int bad = 0;
  double maxd = 0;
  for (int r = 0; r < 128; ++r) {
    const int y = r / 8, xq = r % 8;
    for (int n = 0; n < N; ++n) {
      float acc = 0;
      for (int tap = 0; tap < 9; ++tap) {
        const int ty = tap / 3, tx = tap % 3;
        for (int ci = 0; ci < C; ++ci) acc += fx[((y + ty) * 10 + xq + tx) * C + ci] * fw[(tap * N + n) * C + ci];
      }
      const double d = fabs((double)acc - got[r * 32 + n]);
      if (d > maxd) maxd = d;
      if (d > 1e-3) ++bad;
    }
  }
  printf("SHIFT mode %d (%s): mismatches %d / %d, max |diff| %.3f -> %s\n", mode,
         mode == 0 ? "SWIZZLE_NONE planes, 16B-granular start" : (mode == 1 ? "SWIZZLE_128B rows pitch 16, base_offset=(addr>>7)&7" : (mode == 2 ? "SWIZZLE_128B rows pitch 16, base_offset=0" : (mode == 3 ? "SWIZZLE_128B rows DENSE pitch 10, base_offset=0" : "SWIZZLE_64B rows DENSE pitch 10, base_offset=0"))),
         bad, 128 * N, maxd, bad == 0 ? "OK" : "WRONG");
  cudaFree(dx); cudaFree(dw); cudaFree(dout);
}

// ------------------------------------------------------------------------------------------------ 3: MMA issue rate
// layout 0: SWIZZLE_NONE (A planes with pitch 10, shifted starts), 2: SWIZZLE_128B (aligned tiles)
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(int N, int layout, int iters, long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ uint64_t bar;
  __shared__ uint32_t s_tmem;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  if (threadIdx.x < 32) { tmem_alloc(smem_u32(&s_tmem), 512); tmem_relinquish(); }
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw)))[i] = 0;
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;
  if (threadIdx.x == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(N >> 3) << 17) | (uint32_t(128 >> 4) << 24);
    const uint32_t sA = base, sB = base + 64 * 1024;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const int tap = it % 9, ks = it & 3;
      uint64_t da, db;
      if (layout == 0) {
        da = make_desc(sA + (uint32_t)(2 * ks) * (18 * 10 * 16) + (uint32_t)((tap / 3) * 10 + tap % 3) * 16, 18 * 10 * 16, 160, 0, 0);
        db = make_desc(sB + (uint32_t)(2 * ks) * (N * 16), N * 16, 128, 0, 0);
      } else {
        da = make_desc(sA + (uint32_t)ks * 32 + (uint32_t)(tap & 1) * 16384, 16, 1024, 2, 0);
        db = make_desc(sB + (uint32_t)ks * 32, 16, 1024, 2, 0);
      }
      umma_bf16(tmem + (uint32_t)((it & 1) * 256), da, db, idesc, 1u);
    }
    umma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    const long long t1 = clock64();
    cycles[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

// Unrolled issue loop with precomputed descriptors: one or two issuing warps (each with its own accumulator).
template <int N>
__global__ void __launch_bounds__(128, 1) mma_rate2_kernel(int n_issuers, int iters, int rowbytes, long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ uint64_t bar[2];
  __shared__ uint32_t s_tmem;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar[0]), 1); mbar_init(smem_u32(&bar[1]), 1); fence_barrier_init(); }
  if (threadIdx.x < 32) { tmem_alloc(smem_u32(&s_tmem), 512); tmem_relinquish(); }
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw)))[i] = 0;
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;
  const int warp = threadIdx.x >> 5;
  if (warp < n_issuers) {   // warp-uniform control flow + elect.sync keeps the descriptors in uniform registers
    const bool leader = elect_one() != 0;
    constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(N >> 3) << 17) | (uint32_t(128 >> 4) << 24);
    const uint32_t lay = rowbytes == 128 ? 2u : 4u;
    const uint64_t a0 = make_desc(base + warp * 32768, 16, 10 * rowbytes, lay, 0);
    const uint64_t b0 = make_desc(base + 96 * 1024, 16, 8 * rowbytes, lay, 0);
    const uint32_t tm = tmem + (uint32_t)(warp * 256);
    const int kk = rowbytes / 32;   // k16 steps per row
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        const uint32_t aoff = (uint32_t)(((tap / 3) * 10 + tap % 3) * rowbytes) >> 4;
        const uint32_t boff = (uint32_t)(tap * N * rowbytes) >> 4;
#pragma unroll 4
        for (int ks = 0; ks < kk; ++ks)
          if (leader) umma_bf16(tm, a0 + aoff + 2 * ks, b0 + boff + 2 * ks, idesc, 1u);
      }
    }
    if (leader) umma_commit(smem_u32(&bar[warp]));
    mbar_wait(smem_u32(&bar[warp]), 0);
    const long long t1 = clock64();
    if (leader) cycles[blockIdx.x * 2 + warp] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

// A operand in SWIZZLE_NONE planes ([cgroup][18][10][16B], C = 64 -> 8 planes), B in SWIZZLE_128B rows.
// mode 0: one issuer, one accumulator; 1: one issuer alternating two accumulators; 2: two issuers
template <int N>
__global__ void __launch_bounds__(128, 1) mma_rate3_kernel(int mode, int a_none, int iters, long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ uint64_t bar[2];
  __shared__ uint32_t s_tmem;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar[0]), 1); mbar_init(smem_u32(&bar[1]), 1); fence_barrier_init(); }
  if (threadIdx.x < 32) { tmem_alloc(smem_u32(&s_tmem), 512); tmem_relinquish(); }
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw)))[i] = 0;
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;
  const int warp = threadIdx.x >> 5;
  const int n_issuers = mode == 2 ? 2 : 1;
  if (warp < n_issuers) {
    const bool leader = elect_one() != 0;
    constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(N >> 3) << 17) | (uint32_t(128 >> 4) << 24);
    const uint32_t plane = 18 * 10 * 16;
    const uint64_t a0 = a_none ? make_desc(base + warp * 32768, plane, 160, 0, 0) : make_desc(base + warp * 32768, 16, 1280, 2, 0);
    const uint64_t b0 = make_desc(base + 96 * 1024, 16, 1024, 2, 0);
    const uint32_t tm = tmem + (uint32_t)(warp * 256);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const uint32_t tmi = tm + (mode == 1 ? (uint32_t)((it & 1) * 256) : 0u);
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        const uint32_t aoff = a_none ? (uint32_t)(((tap / 3) * 10 + tap % 3) * 16) >> 4 : (uint32_t)(((tap / 3) * 10 + tap % 3) * 128) >> 4;
        const uint32_t boff = (uint32_t)(tap * N * 128) >> 4;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          if (leader) umma_bf16(tmi, a0 + aoff + (a_none ? (uint32_t)(2 * ks) * (plane >> 4) : (uint32_t)(2 * ks)), b0 + boff + 2 * ks, idesc, 1u);
      }
    }
    if (leader) umma_commit(smem_u32(&bar[warp]));
    mbar_wait(smem_u32(&bar[warp]), 0);
    const long long t1 = clock64();
    if (leader) cycles[blockIdx.x * 2 + warp] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

template <int N>
static void run_rate3(int sms, long long* dcyc) {
  const int smem = 200 * 1024;
  CK(cudaFuncSetAttribute(mma_rate3_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  for (int a_none = 0; a_none <= 1; ++a_none)
    for (int mode = 0; mode <= 2; ++mode) {
      const int iters = 200;
      CK(cudaMemset(dcyc, 0, sms * 2 * sizeof(long long)));
      mma_rate3_kernel<N><<<sms, 128, smem>>>(mode, a_none, iters, dcyc);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("MMA rate3 N=%d failed: %s\n", N, cudaGetErrorString(e)); exit(3); }
      std::vector<long long> h(sms * 2);
      CK(cudaMemcpy(h.data(), dcyc, sms * 2 * sizeof(long long), cudaMemcpyDeviceToHost));
      long long mx = 0;
      for (long long c : h) mx = c > mx ? c : mx;
      const double n_mma = (double)iters * 36 * (mode == 2 ? 2 : 1);
      const double cyc = (double)mx / n_mma;
      printf("MMA_RATE3 A=%s mode=%s N=%3d: %.1f clk/MMA -> %.0f MAC/clk/SM (floor %d clk); smem operand %.0f B/clk\n", a_none ? "NONE " : "SW128",
             mode == 0 ? "1 issuer/1 acc " : (mode == 1 ? "1 issuer/2 accs" : "2 issuers      "), N, cyc, 128.0 * N * 16 / cyc, 128 * N / 256, (4096 + N * 32) / cyc);
    }
}

template <int N>
static void run_rate2(int sms, long long* dcyc) {
  const int smem = 200 * 1024;
  CK(cudaFuncSetAttribute(mma_rate2_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  for (int rowbytes = 64; rowbytes <= 128; rowbytes *= 2)
    for (int issuers = 1; issuers <= 2; ++issuers) {
      if (N * rowbytes * 9 > 100 * 1024) continue;
      const int iters = 200;
      CK(cudaMemset(dcyc, 0, sms * 2 * sizeof(long long)));
      mma_rate2_kernel<N><<<sms, 128, smem>>>(issuers, iters, rowbytes, dcyc);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("MMA rate2 N=%d failed: %s\n", N, cudaGetErrorString(e)); exit(3); }
      std::vector<long long> h(sms * 2);
      CK(cudaMemcpy(h.data(), dcyc, sms * 2 * sizeof(long long), cudaMemcpyDeviceToHost));
      long long mx = 0;
      for (long long c : h) mx = c > mx ? c : mx;
      const double n_mma = (double)iters * 9 * (rowbytes / 32) * issuers;
      const double cyc = (double)mx / n_mma;
      printf("MMA_RATE2 row=%3dB issuers=%d M=128 N=%3d K=16: %.1f clk/MMA -> %.0f MAC/clk/SM (floor %d clk); smem operand %.0f B/clk\n", rowbytes, issuers, N, cyc,
             128.0 * N * 16 / cyc, 128 * N / 256, (4096 + N * 32) / cyc);
    }
}

static void run_mma_rate(int sms) {
  long long* dcyc;
  CK(cudaMalloc(&dcyc, sms * 2 * sizeof(long long)));
  run_rate3<32>(sms, dcyc);
  run_rate3<64>(sms, dcyc);
  run_rate3<128>(sms, dcyc);
  run_rate2<32>(sms, dcyc);
  run_rate2<64>(sms, dcyc);
  run_rate2<128>(sms, dcyc);
  run_rate2<256>(sms, dcyc);
  const int smem = 200 * 1024;
  CK(cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int iters = 4096;
  for (int layout = 0; layout <= 2; layout += 2)
    for (int N = 32; N <= 256; N *= 2) {
      mma_rate_kernel<<<sms, 128, smem>>>(N, layout, iters, dcyc);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("MMA rate N=%d layout=%d failed: %s\n", N, layout, cudaGetErrorString(e)); exit(3); }
      std::vector<long long> h(sms);
      CK(cudaMemcpy(h.data(), dcyc, sms * sizeof(long long), cudaMemcpyDeviceToHost));
      long long mx = 0;
      for (long long c : h) mx = c > mx ? c : mx;
      const double cyc = (double)mx / iters;
      printf("MMA_RATE layout=%s M=128 N=%3d K=16: %.1f clk/MMA -> %.0f MAC/clk/SM (floor %d clk), smem operand bytes/MMA %d -> %.0f B/clk\n",
             layout == 0 ? "NONE  " : "SW128 ", N, cyc, 128.0 * N * 16 / cyc, 128 * N / 256, 4096 + N * 32, (4096 + N * 32) / cyc);
    }
  cudaFree(dcyc);
}

// ------------------------------------------------------------------------------------------------ 3b: swapped operand roles
// Next-round question ("what comes next"): with the WEIGHTS as the A operand (M = 64 or 128 rows, from
// shared memory or resident in TMEM) and the PIXELS as the B operand (N = 128 / 256 row-shifted SWIZZLE_128B views of an
// 8-pixel-wide halo tile), is the MMA rate bound by the B stream alone?  Timing only (operands are zeros / whatever TMEM holds).
__device__ __forceinline__ void umma_bf16_tmem_a(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

template <int M, int N>
__global__ void __launch_bounds__(128, 1) mma_rate_ws_kernel(int a_in_tmem, int iters, long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ uint64_t bar;
  __shared__ uint32_t s_tmem;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  if (threadIdx.x < 32) { tmem_alloc(smem_u32(&s_tmem), 512); tmem_relinquish(); }
  for (int i = threadIdx.x; i < 190 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw)))[i] = 0;
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;
  if (threadIdx.x < 32) {
    const bool leader = elect_one() != 0;
    constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(N >> 3) << 17) | (uint32_t(M >> 4) << 24);
    // B: pixels.  N / 8 groups of 8 x-adjacent pixels, one group per tile row of a 10-pixel-pitch halo (SBO = 10 rows)
    const uint64_t b0 = make_desc(base, 16, 1280, 2, 0);
    // A: weights [tap][M rows][64 channels] as 128-byte SWIZZLE_128B rows, behind the halo region (N/8 + 2 rows of 10 pixels)
    const uint32_t a_base = base + 64 * 1024;
    const uint64_t a0 = make_desc(a_base, 16, 1024, 2, 0);
    const uint32_t tm_a = tmem + 256u;     // TMEM-resident A: one K = 16 slice is 8 columns; 9 taps x 4 slices = 288 columns
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        const uint32_t boff = (uint32_t)(((tap / 3) * 10 + tap % 3) * 128) >> 4;
        const uint32_t aoff = (uint32_t)(tap * M * 128) >> 4;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          if (leader) {
            if (a_in_tmem) umma_bf16_tmem_a(tmem, tm_a + (uint32_t)((tap * 4 + ks) % 28) * 8u, b0 + boff + 2 * ks, idesc, 1u);
            else umma_bf16(tmem, a0 + aoff + 2 * ks, b0 + boff + 2 * ks, idesc, 1u);
          }
        }
      }
    }
    if (leader) umma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    const long long t1 = clock64();
    if (leader) cycles[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

template <int M, int N>
static void run_ws_one(int sms, long long* dcyc) {
  const int smem = 200 * 1024;
  CK(cudaFuncSetAttribute(mma_rate_ws_kernel<M, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  for (int a_in_tmem = 0; a_in_tmem <= 1; ++a_in_tmem) {
    const int iters = 200;
    CK(cudaMemset(dcyc, 0, sms * sizeof(long long)));
    mma_rate_ws_kernel<M, N><<<sms, 128, smem>>>(a_in_tmem, iters, dcyc);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("MMA_WS M=%d N=%d A=%s failed: %s\n", M, N, a_in_tmem ? "tmem" : "smem", cudaGetErrorString(e)); exit(3); }
    std::vector<long long> h(sms);
    CK(cudaMemcpy(h.data(), dcyc, sms * sizeof(long long), cudaMemcpyDeviceToHost));
    long long mx = 0;
    for (long long c : h) mx = c > mx ? c : mx;
    const double cyc = (double)mx / (iters * 36.0);
    printf("MMA_WS weights-as-A M=%3d (%s) pixels-as-B N=%3d: %.1f clk/MMA -> %.0f MAC/clk/SM; smem operand bytes/MMA %d -> %.0f B/clk\n", M,
           a_in_tmem ? "TMEM" : "smem", N, cyc, (double)M * N * 16 / cyc, (a_in_tmem ? 0 : M * 32) + N * 32,
           ((a_in_tmem ? 0 : M * 32) + N * 32) / cyc);
  }
}

static void run_ws(int sms) {
  long long* dcyc;
  CK(cudaMalloc(&dcyc, sms * sizeof(long long)));
  run_ws_one<64, 128>(sms, dcyc);
  run_ws_one<64, 256>(sms, dcyc);
  run_ws_one<128, 128>(sms, dcyc);
  run_ws_one<128, 256>(sms, dcyc);
  cudaFree(dcyc);
}

// ------------------------------------------------------------------------------------------------ 3b': swapped roles, correctness
// D[co][pixel] = sum_tap sum_ci W[tap][co][ci] * X[pixel + shift(tap)][ci] with the weights as the A operand (M rows = output
// channels) and NP = 128 pixels (16 rows of 8, pitch-10 halo) as the B operand through row-shifted SWIZZLE_128B views.  The
// kernel dumps all 128 TMEM lanes x NP columns; the host finds which lane holds which output channel (the M = 64
// accumulator layout is not in the guides at hand) and checks the values.
template <int M>
__global__ void __launch_bounds__(128, 1) wst_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ w, float* __restrict__ out) {
  constexpr int C = 64, P = 10, NP = 128, RB = 128;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sp = smem_raw + (base - smem_u32(smem_raw));
  __shared__ uint64_t bar;
  __shared__ uint32_t s_tmem;
  const uint32_t sX = base, sW = base + 32 * 1024;        // halo 18 x 10 rows of 128 B = 23 KB; weights 9 x M x 128 B
  for (int i = threadIdx.x; i < 18 * 10 * (C / 8); i += blockDim.x) {
    const int g = i % (C / 8), px = (i / (C / 8)) % 10, py = i / (C / 8) / 10;
    const uint4 v = *reinterpret_cast<const uint4*>(x + ((size_t)(py * 10 + px) * C + g * 8));
    const uint32_t row = (uint32_t)(py * P + px) * RB;
    *reinterpret_cast<uint4*>(sp + row + (uint32_t)((g ^ ((row >> 7) & 7u)) * 16)) = v;
  }
  for (int i = threadIdx.x; i < 9 * M * (C / 8); i += blockDim.x) {
    const int g = i % (C / 8), co = (i / (C / 8)) % M, tap = i / (C / 8) / M;
    const uint4 v = *reinterpret_cast<const uint4*>(w + ((size_t)(tap * M + co) * C + g * 8));
    const uint32_t row = (uint32_t)tap * (M * RB) + (uint32_t)co * RB;
    *reinterpret_cast<uint4*>(sp + (sW - base) + row + (uint32_t)((g ^ ((row >> 7) & 7u)) * 16)) = v;
  }
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  if (threadIdx.x < 32) { tmem_alloc(smem_u32(&s_tmem), 128); tmem_relinquish(); }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;
  if (threadIdx.x == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(NP >> 3) << 17) | (uint32_t(M >> 4) << 24);
    int first = 1;
    for (int tap = 0; tap < 9; ++tap) {
      const int ty = tap / 3, tx = tap % 3;
      for (int ks = 0; ks < C / 16; ++ks) {
        const uint64_t da = make_desc(sW + (uint32_t)tap * (M * RB) + (uint32_t)ks * 32, 16, 8 * RB, 2, 0);          // weights: dense rows
        const uint64_t db = make_desc(sX + (uint32_t)(ty * P + tx) * RB + (uint32_t)ks * 32, 16, P * RB, 2, 0);      // pixels: shifted view
        umma_bf16(tmem, da, db, idesc, first ? 0u : 1u);
        first = 0;
      }
    }
    umma_commit(smem_u32(&bar));
  }
  mbar_wait(smem_u32(&bar), 0);
  tc_fence_after();
  const int warp = threadIdx.x >> 5;
  for (int cb = 0; cb < NP; cb += 32) {
    uint32_t v[32];
    tmem_ld_32x32b_x32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)cb, v);
    tmem_ld_wait();
    for (int c = 0; c < 32; ++c) out[threadIdx.x * NP + cb + c] = __uint_as_float(v[c]);
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 128);
}

template <int M>
static void run_wst() {
  const int C = 64, NP = 128;
  std::vector<__nv_bfloat16> hx(18 * 10 * C), hw(9 * M * C);
  std::vector<float> fx(hx.size()), fw(hw.size());
  srand(4321 + M);
  for (size_t i = 0; i < hx.size(); ++i) { const float v = (float)(rand() % 17 - 8) / 8.0f; hx[i] = __float2bfloat16(v); fx[i] = __bfloat162float(hx[i]); }
  for (size_t i = 0; i < hw.size(); ++i) { const float v = (float)(rand() % 13 - 6) / 16.0f; hw[i] = __float2bfloat16(v); fw[i] = __bfloat162float(hw[i]); }
  __nv_bfloat16 *dx, *dw;
  float* dout;
  CK(cudaMalloc(&dx, hx.size() * 2)); CK(cudaMalloc(&dw, hw.size() * 2)); CK(cudaMalloc(&dout, 128 * NP * 4));
  CK(cudaMemcpy(dx, hx.data(), hx.size() * 2, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dw, hw.data(), hw.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(dout, 0xff, 128 * NP * 4));
  const int smem = 160 * 1024;
  CK(cudaFuncSetAttribute(wst_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  wst_kernel<M><<<1, 128, smem>>>(dx, dw, dout);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("WST M=%d: kernel failed: %s\n", M, cudaGetErrorString(e)); exit(2); }
  std::vector<float> got(128 * NP);
  CK(cudaMemcpy(got.data(), dout, got.size() * 4, cudaMemcpyDeviceToHost));
  // reference D[co][pixel], pixel = y * 8 + x of the 16 x 8 tile inside the 18 x 10 halo
  std::vector<float> ref((size_t)M * NP, 0.f);
  for (int co = 0; co < M; ++co)
    for (int p = 0; p < NP; ++p) {
      const int y = p / 8, xq = p % 8;
      float acc = 0.f;
      for (int tap = 0; tap < 9; ++tap)
        for (int ci = 0; ci < C; ++ci) acc += fw[((size_t)tap * M + co) * C + ci] * fx[((size_t)(y + tap / 3) * 10 + xq + tap % 3) * C + ci];
      ref[(size_t)co * NP + p] = acc;
    }
  // which TMEM lane holds which output channel?
  int mapped = 0, bad_values = 0;
  printf("WST weights-as-A M=%d, 128 pixels as B: lane -> channel:", M);
  for (int lane = 0; lane < 128; ++lane) {
    int best = -1;
    for (int co = 0; co < M && best < 0; ++co) {
      int ok = 1;
      for (int p = 0; p < NP && ok; ++p) ok = fabsf(got[(size_t)lane * NP + p] - ref[(size_t)co * NP + p]) <= 1e-2f * (1.f + fabsf(ref[(size_t)co * NP + p]));
      if (ok) best = co;
    }
    if (best >= 0) { ++mapped; if (lane < 8 || lane % 16 == 0) printf(" %d->%d", lane, best); }
  }
  for (int co = 0; co < M; ++co) {   // every channel must be somewhere
    int found = 0;
    for (int lane = 0; lane < 128 && !found; ++lane) {
      int ok = 1;
      for (int p = 0; p < NP && ok; ++p) ok = fabsf(got[(size_t)lane * NP + p] - ref[(size_t)co * NP + p]) <= 1e-2f * (1.f + fabsf(ref[(size_t)co * NP + p]));
      found = ok;
    }
    bad_values += !found;
  }
  printf("\nWST M=%d: %d lanes hold a channel, %d of %d channels not found -> %s\n", M, mapped, bad_values, M, bad_values == 0 ? "OK" : "WRONG");
  cudaFree(dx); cudaFree(dw); cudaFree(dout);
}

// ------------------------------------------------------------------------------------------------ 3c: A-operand collector reuse
// Up-path phases share input views: consecutive MMAs with the SAME A tile and different B (weights of another phase) can
// keep A in the tensor pipe's collector (.collector::a::fill / ::use / ::lastuse -> SASS A_KEEP / A_REUSE).  Does a reused A
// take the 32 clk of its shared-memory read off the MMA?  Timing only.
template <int MODE>   // 0: plain, 1: fill, 2: use, 3: lastuse
__device__ __forceinline__ void umma_bf16_coll(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc) {
  if (MODE == 0) { umma_bf16(tmem_d, desc_a, desc_b, idesc, 1u); return; }
#define RCU_COLL_MMA(Q)                                                                                                   \
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::f16.collector::a::" Q        \
               " [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc) : "memory")
  if (MODE == 1) RCU_COLL_MMA("fill");
  if (MODE == 2) RCU_COLL_MMA("use");
  if (MODE == 3) RCU_COLL_MMA("lastuse");
#undef RCU_COLL_MMA
}

template <int N, int GROUP>   // GROUP consecutive MMAs share one A tile
__global__ void __launch_bounds__(128, 1) mma_rate_coll_kernel(int use_collector, int iters, long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ uint64_t bar;
  __shared__ uint32_t s_tmem;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  if (threadIdx.x < 32) { tmem_alloc(smem_u32(&s_tmem), 512); tmem_relinquish(); }
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw)))[i] = 0;
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;
  if (threadIdx.x < 32) {
    const bool leader = elect_one() != 0;
    constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(N >> 3) << 17) | (uint32_t(128 >> 4) << 24);
    const uint64_t a0 = make_desc(base, 16, 1280, 2, 0);
    const uint64_t b0 = make_desc(base + 96 * 1024, 16, 1024, 2, 0);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int view = 0; view < 9; ++view) {
        const uint32_t aoff = (uint32_t)(((view / 3) * 10 + view % 3) * 128) >> 4;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
          for (int g = 0; g < GROUP; ++g) {     // phase g: its own accumulator and its own weights, the same A tile
            const uint32_t d = tmem + (uint32_t)(g * N);
            const uint64_t b = b0 + ((uint32_t)(g * N * 128) >> 4) + 2 * ks;
            if (leader) {
              if (!use_collector || GROUP == 1) umma_bf16_coll<0>(d, a0 + aoff + 2 * ks, b, idesc);
              else if (g == 0) umma_bf16_coll<1>(d, a0 + aoff + 2 * ks, b, idesc);
              else if (g + 1 < GROUP) umma_bf16_coll<2>(d, a0 + aoff + 2 * ks, b, idesc);
              else umma_bf16_coll<3>(d, a0 + aoff + 2 * ks, b, idesc);
            }
          }
        }
      }
    }
    if (leader) umma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    const long long t1 = clock64();
    if (leader) cycles[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

template <int N, int GROUP>
static void run_coll_one(int sms, long long* dcyc) {
  const int smem = 200 * 1024;
  CK(cudaFuncSetAttribute(mma_rate_coll_kernel<N, GROUP>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  for (int use = 0; use <= 1; ++use) {
    const int iters = 100;
    CK(cudaMemset(dcyc, 0, sms * sizeof(long long)));
    mma_rate_coll_kernel<N, GROUP><<<sms, 128, smem>>>(use, iters, dcyc);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("MMA_COLL N=%d group=%d failed: %s\n", N, GROUP, cudaGetErrorString(e)); exit(3); }
    std::vector<long long> h(sms);
    CK(cudaMemcpy(h.data(), dcyc, sms * sizeof(long long), cudaMemcpyDeviceToHost));
    long long mx = 0;
    for (long long c : h) mx = c > mx ? c : mx;
    printf("MMA_COLL M=128 N=%3d, %d MMAs per A tile, collector %s: %.1f clk/MMA\n", N, GROUP, use ? "fill/use/lastuse" : "off             ",
           (double)mx / (iters * 36.0 * GROUP));
  }
}

static void run_coll(int sms) {
  long long* dcyc;
  CK(cudaMalloc(&dcyc, sms * sizeof(long long)));
  run_coll_one<32, 2>(sms, dcyc);
  run_coll_one<32, 4>(sms, dcyc);
  run_coll_one<64, 2>(sms, dcyc);
  run_coll_one<64, 4>(sms, dcyc);
  cudaFree(dcyc);
}

// ------------------------------------------------------------------------------------------------ 4: L2 -> smem bandwidth
__global__ void __launch_bounds__(256, 1) cpasync_bw_kernel(const uint8_t* __restrict__ src, size_t bytes_per_cta, int rounds, long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  const uint8_t* my = src + (size_t)blockIdx.x * bytes_per_cta;
  const uint32_t s0 = smem_u32(smem_raw);
  const int chunk = 256 * 16;  // bytes per block-wide cp.async
  const int per_round = (int)(bytes_per_cta / chunk);
  __syncthreads();
  const long long t0 = clock64();
  for (int r = 0; r < rounds; ++r) {
    for (int i = 0; i < per_round; ++i) {
      const uint32_t dst = s0 + (uint32_t)((i & 15) * chunk) + threadIdx.x * 16;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(my + (size_t)i * chunk + threadIdx.x * 16) : "memory");
      if ((i & 3) == 3) asm volatile("cp.async.commit_group;" ::: "memory");
      if ((i & 3) == 3) asm volatile("cp.async.wait_group 3;" ::: "memory");
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

__global__ void __launch_bounds__(32, 1) bulk_bw_kernel(const uint8_t* __restrict__ src, size_t bytes_per_cta, int rounds, int piece, long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bars[8];
  const uint8_t* my = src + (size_t)blockIdx.x * bytes_per_cta;
  const uint32_t s0 = (smem_u32(smem_raw) + 127u) & ~127u;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) mbar_init(smem_u32(&bars[i]), 1);
    fence_barrier_init();
    const int per_round = (int)(bytes_per_cta / piece);
    const long long t0 = clock64();
    long long n = 0;
    for (int r = 0; r < rounds; ++r)
      for (int i = 0; i < per_round; ++i, ++n) {
        const int slot = (int)(n & 7);
        if (n >= 8) mbar_wait(smem_u32(&bars[slot]), (uint32_t)(((n >> 3) - 1) & 1));
        mbar_expect_tx(smem_u32(&bars[slot]), (uint32_t)piece);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s0 + (uint32_t)(slot * piece)),
                     "l"(my + (size_t)i * piece), "r"(piece), "r"(smem_u32(&bars[slot]))
                     : "memory");
      }
    for (long long k = (n >= 8 ? n - 8 : 0); k < n; ++k) mbar_wait(smem_u32(&bars[k & 7]), (uint32_t)((k >> 3) & 1));
    const long long t1 = clock64();
    cycles[blockIdx.x] = t1 - t0;
  }
}

static void run_bw(int sms) {
  const size_t per_cta = 192 * 1024;           // 148 * 192 KB = 28 MB, L2 resident after the first round
  uint8_t* src;
  long long* dcyc;
  CK(cudaMalloc(&src, per_cta * sms));
  CK(cudaMemset(src, 1, per_cta * sms));
  CK(cudaMalloc(&dcyc, sms * sizeof(long long)));
  CK(cudaFuncSetAttribute(cpasync_bw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  CK(cudaFuncSetAttribute(bulk_bw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 16384 + 256));
  std::vector<long long> h(sms);
  for (int rep = 0; rep < 2; ++rep) {
    const int rounds = 20;
    cpasync_bw_kernel<<<sms, 256, 64 * 1024>>>(src, per_cta, rounds, dcyc);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h.data(), dcyc, sms * sizeof(long long), cudaMemcpyDeviceToHost));
    long long mx = 0;
    for (long long c : h) mx = c > mx ? c : mx;
    if (rep) printf("L2_BW cp.async 16B x 256 thr, %d SMs: %.1f B/clk/SM (slowest SM), chip %.0f B/clk\n", sms, (double)per_cta * rounds / mx, (double)per_cta * rounds / mx * sms);
  }
  for (int piece = 2048; piece <= 16384; piece *= 2) {
    const int rounds = 20;
    for (int rep = 0; rep < 2; ++rep) {
      bulk_bw_kernel<<<sms, 32, 8 * 16384 + 256>>>(src, per_cta, rounds, piece, dcyc);
      CK(cudaDeviceSynchronize());
    }
    CK(cudaMemcpy(h.data(), dcyc, sms * sizeof(long long), cudaMemcpyDeviceToHost));
    long long mx = 0;
    for (long long c : h) mx = c > mx ? c : mx;
    printf("L2_BW cp.async.bulk %5d B pieces, 8 in flight, %d SMs: %.1f B/clk/SM, chip %.0f B/clk\n", piece, sms, (double)per_cta * rounds / mx, (double)per_cta * rounds / mx * sms);
  }
  cudaFree(src); cudaFree(dcyc);
}

// ------------------------------------------------------------------------------------------------ 5: what does a 5-D pair box look like in smem?
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void __launch_bounds__(128, 1) tma5d_kernel(const __grid_constant__ CUtensorMap map, uint16_t* out, int bytes, int x0, int y0) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sp = smem_raw + (base - smem_u32(smem_raw));
  __shared__ uint64_t bar;
  for (int i = threadIdx.x; i < bytes / 2; i += blockDim.x) reinterpret_cast<uint16_t*>(sp)[i] = 0xFFFF;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  fence_proxy_async();
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(smem_u32(&bar), (uint32_t)bytes);
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(base), "l"(reinterpret_cast<uint64_t>(&map)), "r"(smem_u32(&bar)), "r"(0), "r"(0), "r"(x0), "r"(y0), "r"(0)
        : "memory");
  }
  mbar_wait(smem_u32(&bar), 0);
  for (int i = threadIdx.x; i < bytes / 2; i += blockDim.x) out[i] = reinterpret_cast<uint16_t*>(sp)[i];
}

static void run_tma5d() {
  const int H = 20, W = 16, C = 32, BOXX = 10, BOXY = 17;
  // rows: one zero row above, H image rows, one zero row below; ids are unique 16-bit patterns (0 reserved for zero rows)
  std::vector<uint16_t> h((size_t)(H + 2) * W * C, 0);
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x)
      for (int c = 0; c < C; ++c) h[((size_t)(y + 1) * W + x) * C + c] = (uint16_t)(1 + (y * W + x) * C + c);
  uint16_t *d, *dout;
  CK(cudaMalloc(&d, h.size() * 2));
  CK(cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice));
  const int bytes = C * 2 * 2 * BOXX * BOXY;
  CK(cudaMalloc(&dout, bytes));
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
  EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(fp);
  const cuuint64_t row_bytes = (cuuint64_t)W * C * 2;
  CUtensorMap map;
  cuuint64_t dims[5] = {32, 2, (cuuint64_t)W, (cuuint64_t)H + 1, 1};
  cuuint64_t strides[4] = {row_bytes, (cuuint64_t)C * 2, row_bytes, (cuuint64_t)(H + 1) * row_bytes};
  cuuint32_t box[5] = {32, 2, BOXX, BOXY, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("TMA5D encode rc=%d\n", (int)r);
  if (r != CUDA_SUCCESS) return;
  CK(cudaFuncSetAttribute(tma5d_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  const int x0 = 3, y0 = 2;   // interior window: y' = 2..18, x = 3..12
  tma5d_kernel<<<1, 128, 64 * 1024>>>(map, dout, bytes, x0, y0);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("TMA5D kernel failed: %s\n", cudaGetErrorString(e)); exit(4); }
  std::vector<uint16_t> got(bytes / 2);
  CK(cudaMemcpy(got.data(), dout, bytes, cudaMemcpyDeviceToHost));
  // hypothesis A: dense [y'][x][pair][c] with 16-byte chunk index XOR ((byte_offset >> 7) & 7)
  int badA = 0;
  for (int yy = 0; yy < BOXY; ++yy)
    for (int xx = 0; xx < BOXX; ++xx)
      for (int pr = 0; pr < 2; ++pr)
        for (int c = 0; c < C; ++c) {
          const int img_row = (y0 + yy) + pr - 1;   // image row
          const uint16_t want = (img_row < 0 || img_row >= H) ? 0 : (uint16_t)(1 + (img_row * W + (x0 + xx)) * C + c);
          const int line = yy * BOXX + xx;
          const int chunk = pr * 4 + c / 8;
          const int off = line * 64 + ((chunk ^ (line & 7)) * 8) + (c & 7);   // in uint16 units
          if (got[off] != want) ++badA;
        }
  printf("TMA5D hypothesis A (dense rows, address-based XOR): %d mismatches of %d\n", badA, BOXY * BOXX * 2 * C);
  // dump where the first elements of the first two lines ended up
  for (int slot = 0; slot < 2 * 64; slot += 8) {
    const uint16_t v = got[slot];
    if (v == 0 || v == 0xFFFF) { printf("  smem u16[%3d] = %s\n", slot, v ? "untouched" : "zero"); continue; }
    const int id = v - 1, c = id % C, x = (id / C) % W, y = id / C / W;
    printf("  smem u16[%3d] <- image row %d, x %d, c %d\n", slot, y, x, c);
  }
  cudaFree(d); cudaFree(dout);
}

// ------------------------------------------------------------------------------------------------ 6: TMA halo-box load throughput
__global__ void __launch_bounds__(32, 1) tma_box_kernel(const __grid_constant__ CUtensorMap map, int n_img, int tiles_x, int tiles_y, int stages,
                                                        uint32_t box_bytes, uint32_t slot, long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ uint64_t bars[16];
  if (threadIdx.x == 0) {
    for (int i = 0; i < stages; ++i) mbar_init(smem_u32(&bars[i]), 1);
    fence_barrier_init();
    const int tiles_per_img = tiles_x * tiles_y;
    const long long total = (long long)n_img * tiles_per_img;
    const int t0i = (int)(total * blockIdx.x / gridDim.x), t1i = (int)(total * (blockIdx.x + 1) / gridDim.x);
    const long long t0 = clock64();
    int n = 0;
    for (int tile = t0i; tile < t1i; ++tile, ++n) {
      const int img = tile / tiles_per_img, rem = tile - img * tiles_per_img, ty = rem / tiles_x, tx = rem - ty * tiles_x;
      const int s = n % stages;
      if (n >= stages) mbar_wait(smem_u32(&bars[s]), (uint32_t)(((n / stages) - 1) & 1));
      mbar_expect_tx(smem_u32(&bars[s]), box_bytes);
      tma_load_4d(base + (uint32_t)s * slot, &map, smem_u32(&bars[s]), 0, tx * 8 - 1, ty * 16 - 1, img);
    }
    for (int k = (n >= stages ? n - stages : 0); k < n; ++k) mbar_wait(smem_u32(&bars[k % stages]), (uint32_t)((k / stages) & 1));
    cycles[blockIdx.x] = clock64() - t0;
  }
}

static void run_tma_box(int sms) {
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
  EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(fp);
  long long* dcyc;
  CK(cudaMalloc(&dcyc, sms * sizeof(long long)));
  CK(cudaFuncSetAttribute(tma_box_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  const int H = 240, W = 240;
  for (int C = 32; C <= 64; C *= 2)
    for (int n_img = 8; n_img <= 128; n_img *= 16) {    // 8 images: L2 resident after one pass; 128: streams from HBM
      uint8_t* d;
      const size_t bytes = (size_t)n_img * H * W * C * 2;
      CK(cudaMalloc(&d, bytes));
      CK(cudaMemset(d, 1, bytes));
      CUtensorMap map;
      cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)n_img};
      cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
      cuuint32_t box[4] = {(cuuint32_t)C, 10, 18, 1};
      cuuint32_t estr[4] = {1, 1, 1, 1};
      CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { printf("TMA_BOX encode failed %d\n", (int)r); exit(5); }
      for (int stages = 2; stages <= 8; stages *= 2) {
        std::vector<long long> h(sms);
        for (int rep = 0; rep < 2; ++rep) {
          tma_box_kernel<<<sms, 32, 200 * 1024>>>(map, n_img, 30, 15, stages, (uint32_t)(C * 2 * 180), 23552u, dcyc);
          CK(cudaDeviceSynchronize());
        }
        CK(cudaMemcpy(h.data(), dcyc, sms * sizeof(long long), cudaMemcpyDeviceToHost));
        long long mx = 0;
        for (long long c : h) mx = c > mx ? c : mx;
        const double tiles_per_sm = (double)n_img * 450 / sms;
        printf("TMA_BOX C=%d (%d B rows) images=%3d (%s) stages=%d: %.0f clk/box, %.1f B/clk/SM\n", C, C * 2, n_img,
               n_img <= 8 ? "L2" : "HBM", stages, mx / tiles_per_sm, (double)C * 2 * 180 * tiles_per_sm / mx);
      }
      cudaFree(d);
    }
  cudaFree(dcyc);
}

int main(int argc, char** argv) {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  printf("device %s, %d SMs, cc %d.%d\n", prop.name, prop.multiProcessorCount, prop.major, prop.minor);
  const char* what = argc > 1 ? argv[1] : "all";
  const bool all = std::string(what) == "all";
  if (all || std::string(what) == "shift0") run_shift(0);
  if (all || std::string(what) == "shift1") run_shift(1);
  if (all || std::string(what) == "shift2") run_shift(2);
  if (all || std::string(what) == "shift3") run_shift(3);
  if (all || std::string(what) == "shift4") run_shift(4);
  if (all || std::string(what) == "tma5d") run_tma5d();
  if (all || std::string(what) == "tmabox") run_tma_box(prop.multiProcessorCount);
  if (all || std::string(what) == "rate") run_mma_rate(prop.multiProcessorCount);
  if (all || std::string(what) == "ws") run_ws(prop.multiProcessorCount);
  if (all || std::string(what) == "coll") run_coll(prop.multiProcessorCount);
  if (all || std::string(what) == "wst") { run_wst<128>(); run_wst<64>(); }
  if (all || std::string(what) == "bw") run_bw(prop.multiProcessorCount);
  return 0;
}
