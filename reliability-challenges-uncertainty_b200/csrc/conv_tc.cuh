// Implicit-GEMM 3x3 / 2x2-phase convolution on the 5th-generation tensor cores (sm_100a):
//   TMA (cp.async.bulk.tensor, tiled mode, zero OOB fill = the conv padding) -> swizzled shared memory
//   -> tcgen05.mma (kind::f16, bf16 x bf16 -> fp32) accumulating in TMEM -> tcgen05.ld epilogue that applies the
//   folded (Dropout2d scale, BatchNorm, bias) per-(image, channel) coefficients + ReLU, and either stores bf16
//   NHWC activations or (for conv_cls.0) applies the 1x1 head and stores logits.
//
// GEMM view of one output tile:  D[128 pixels (8 rows x 16 cols of one image)][BLOCK_N out-channels]
//     = sum over taps (dy,dx) and K-chunks of KC input channels of
//       A[128 pixels][KC] (the input window shifted by the tap; one 4-D TMA box)  x  B[BLOCK_N][KC]^T (weights).
// The two sources of a decoder conv (cat((up, skip), 1), common/model/unet.py:118) are two TMA tensor maps
// walked back to back along K — the concatenated tensor is never materialised.  nearest-x2 + conv3x3
// (common/model/unet.py:105, helpers.py:5-16) runs as four 2x2-tap phase convolutions on the low-res input
// with pre-summed weights (n_phases = 4), writing the interleaved high-res output directly.
//
// Persistent, warp-specialised CTA (192 threads): warp 0 = TMA producer, warp 1 = MMA issuer + TMEM owner,
// warps 2..5 = epilogue (warp w may touch TMEM lanes 32*(w%4)..+31).  Two TMEM accumulator stages let the
// epilogue of tile i overlap the MMAs of tile i+1.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>

#include "common.cuh"

namespace rcu {

constexpr int kTileH = 8;
constexpr int kTileW = 16;
constexpr int kTileM = kTileH * kTileW;  // 128 = UMMA M
constexpr int kConvThreads = 192;
constexpr int kMaxTaps = 9;
constexpr int kMaxPhases = 4;

struct ConvParams {
  int n_img;                 // images in this launch
  int in_h, in_w;            // spatial size of the (low-res for phase convs) input = GEMM pixel grid
  int tiles_x, tiles_y;
  int n_tiles_n;             // c_out / BLOCK_N
  int kc0, kc1;              // K-chunks taken from source 0 / source 1
  int n_taps, n_phases;
  signed char dy[kMaxPhases][kMaxTaps];
  signed char dx[kMaxPhases][kMaxTaps];
  // output addressing: pixel (y, x) of phase (a, b) lands at (out_mul*y + a, out_mul*x + b)
  int out_mul;
  int out_h, out_w;
  int out_c;                 // pixel stride of the output tensor in elements (a channel offset is folded into `out`)
  long long out_img_stride;  // elements between images (padded tensors carry one extra row per image)
  __nv_bfloat16* out;
  // folded epilogue coefficients: float2 (scale, shift) per (image, channel)
  const float2* coef;
  long long coef_stride;     // float2 elements per image row
  int coef_off;              // first column of this layer
  int relu;
  // fused 1x1 head (only when BLOCK_N == c_out == 32): logits[img_global][pixel][2]
  const float* head;         // [2][32] weights then [2] bias, fp32; NULL when not fused
  float* logits;
  int head_diff;             // 1: the fused head stores l0 - l1 per pixel instead of the logit pair
  int chunk_slices;          // images are ordered [sample][slice-in-chunk]
  long long slice0, n_slices_total;
  // residual branch of a ConvResidualBlock (common/model/unet.py:57-59): out = bf16(float(out) + acc * scale + shift),
  // i.e. the 1x1 convolution of the block input is added onto the block output already in `out`
  int accumulate;
};

// ---------------------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// One arrival for the whole (converged) warp: every lane has finished its TMEM reads (tcgen05.wait::ld is warp-collective and
// each lane has issued tcgen05.fence::before_thread_sync), lane 0 signals.  A per-thread arrive is 32 serialised shared-memory
// atomics on one word per warp and tile — on the data pipe the tensor cores read their operands from.
#ifndef RCU_ARRIVE_PER_THREAD
#define RCU_ARRIVE_PER_THREAD 0      // A/B switch (tools/ab_variants.sh): 1 restores one arrival per epilogue thread
#endif
constexpr uint32_t kEpilogueArrivals = RCU_ARRIVE_PER_THREAD ? 128u : 4u;   // arrival count of the accumulator-free barriers
__device__ __forceinline__ void mbar_arrive_warp(uint32_t bar) {
#if RCU_ARRIVE_PER_THREAD
  mbar_arrive(bar);
#else
  __syncwarp();
  if ((threadIdx.x & 31) == 0) mbar_arrive(bar);
#endif
}
// RCU_WAIT_HINT_NS > 0 passes the suspend-time hint of mbarrier.try_wait (the thread may sleep in hardware up to that long
// instead of returning to the polling loop); RCU_SLACK_SLEEP_NS > 0 makes the waits that have slack (epilogue groups waiting
// for their accumulator, the producer waiting for a free stage, the patch warps) back off with nanosleep between polls.
// ncu counts 78 polls per epilogue warp and tile (12.8 M per pixel-pair launch), yet neither switch moves the forward time
// (A/B on B200, profiles/r02zz_ab_experiments.log: hint 1000 ns, sleep 50 / 200 / 500 ns all within 1 % of plain polling):
// the polls do not compete with the tensor cores' operand reads.  Both default to off.
#ifndef RCU_WAIT_HINT_NS
#define RCU_WAIT_HINT_NS 0
#endif
#ifndef RCU_SLACK_SLEEP_NS
#define RCU_SLACK_SLEEP_NS 0
#endif
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
#if RCU_WAIT_HINT_NS > 0
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity), "r"((uint32_t)RCU_WAIT_HINT_NS)
      : "memory");
#else
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
#endif
  return ok != 0;
}
// Bounded wait: a protocol bug must trap, never hang the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {
      printf("rcu conv_tc: mbarrier timeout (block %d thread %d bar 0x%x parity %u)\n", blockIdx.x, threadIdx.x, bar, parity);
      __trap();
    }
  }
}
// The same for waiters with slack: sleep between polls instead of hammering the barrier word.
__device__ __forceinline__ void mbar_wait_slack(uint32_t bar, uint32_t parity) {
#if RCU_SLACK_SLEEP_NS > 0
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (true) {
    __nanosleep(RCU_SLACK_SLEEP_NS);
    if (mbar_try_wait(bar, parity)) return;
    if (clock64() - t0 > 4000000000LL) {
      printf("rcu conv_tc: mbarrier timeout (block %d thread %d bar 0x%x parity %u)\n", blockIdx.x, threadIdx.x, bar, parity);
      __trap();
    }
  }
#else
  mbar_wait(bar, parity);
#endif
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Epilogue coefficients of two adjacent channels: one 16-byte broadcast read (the 8-byte reads the compiler emits for two
// float2 loads cost twice the shared-memory wavefronts).  RCU_COEF_LDS64=1 restores the separate loads (A/B).
#ifndef RCU_COEF_LDS64
#define RCU_COEF_LDS64 0
#endif
#if RCU_COEF_LDS64
#define RCU_COEF2(IDX) const float2 c0 = coef[(IDX)], c1 = coef[(IDX) + 1]
#else
#define RCU_COEF2(IDX)                                                      \
  const float4 cc_ = *reinterpret_cast<const float4*>(&coef[(IDX)]);        \
  const float2 c0 = make_float2(cc_.x, cc_.y), c1 = make_float2(cc_.z, cc_.w)
#endif

// K-major operand tile descriptor (PTX "matrix descriptor"; field layout as in cute::UMMA::SmemDescriptor):
//   [0,14) start address >> 4 | [16,30) leading byte offset >> 4 (unused for swizzled K-major) |
//   [32,46) stride byte offset >> 4 (distance between 8-row groups) | [46,48) version = 1 | [61,64) swizzle mode
// Rows are KC*2 bytes wide (128 B -> SWIZZLE_128B = 2, 64 B -> SWIZZLE_64B = 4), 8-row groups are dense.
template <int KC>
__device__ __forceinline__ uint64_t make_kmajor_desc(uint32_t smem_addr) {
  constexpr uint64_t row_bytes = KC * 2;
  constexpr uint64_t sbo = 8 * row_bytes;
  constexpr uint64_t layout = (KC == 64) ? 2 : 4;
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | (uint64_t(1) << 16) | ((sbo >> 4) << 32) | (uint64_t(1) << 46) | (layout << 61);
}

// Instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 (1) at [4,6), a/b format BF16 (1) at [7,10)/[10,13),
// both K-major, N>>3 at [17,23), M>>4 at [24,29).
template <int BLOCK_N>
__device__ __forceinline__ constexpr uint32_t make_idesc() {
  return (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(BLOCK_N >> 3) << 17) | (uint32_t(kTileM >> 4) << 24);
}

template <int BLOCK_N, int KC>
struct ConvSmem {
  static constexpr int kABytes = kTileM * KC * 2;
  static constexpr int kBBytes = BLOCK_N * KC * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kBudget = 200 * 1024;
  static constexpr int kStagesRaw = kBudget / kStageBytes;
  static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
  static constexpr int kCoefBytes = 2 * BLOCK_N * (int)sizeof(float2);
  static constexpr int kHeadBytes = 2 * 32 * 4 + 16;
  static constexpr int kBarBytes = (2 * kStages + 4) * 8 + 16;
  static constexpr int kTotal = 1024 /*alignment slack*/ + kStages * kStageBytes + kCoefBytes + kHeadBytes + kBarBytes;
  static constexpr int kTmemCols = (2 * BLOCK_N) < 32 ? 32 : (2 * BLOCK_N);
};

template <int BLOCK_N, int KC>
__global__ void __launch_bounds__(kConvThreads, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap map_a0, const __grid_constant__ CUtensorMap map_a1,
               const __grid_constant__ CUtensorMap map_w, const ConvParams prm) {
  using S = ConvSmem<BLOCK_N, KC>;
  static_assert(KC == 64 || KC == 32, "K chunk is one swizzle row: 64 (128B) or 32 (64B) bf16 channels");
  static_assert(BLOCK_N % 32 == 0 && BLOCK_N >= 32 && BLOCK_N <= 256, "BLOCK_N in {32,64,128,256}");
  extern __shared__ uint8_t smem_dyn[];
  const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_dyn + (base - smem_u32(smem_dyn));
  const uint32_t smem_a = base;
  const uint32_t smem_b = base + S::kStages * S::kABytes;
  float2* s_coef = reinterpret_cast<float2*>(base_ptr + S::kStages * S::kStageBytes);
  float* s_head = reinterpret_cast<float*>(base_ptr + S::kStages * S::kStageBytes + S::kCoefBytes);
  uint8_t* bar_ptr = base_ptr + S::kStages * S::kStageBytes + S::kCoefBytes + S::kHeadBytes;
  const uint32_t bar_full = smem_u32(bar_ptr);                       // [kStages]
  const uint32_t bar_empty = bar_full + S::kStages * 8;              // [kStages]
  const uint32_t bar_tfull = bar_empty + S::kStages * 8;             // [2]
  const uint32_t bar_tempty = bar_tfull + 16;                        // [2]
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bar_ptr + (2 * S::kStages + 4) * 8);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a0);
    tma_prefetch_desc(&map_a1);
    tma_prefetch_desc(&map_w);
    for (int s = 0; s < S::kStages; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar_tfull + 8 * s, 1);
      mbar_init(bar_tempty + 8 * s, 128);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(s_tmem), S::kTmemCols);
    tmem_relinquish();
  }
  if (prm.head != nullptr && threadIdx.x >= 64 && threadIdx.x < 64 + 66) s_head[threadIdx.x - 64] = prm.head[threadIdx.x - 64];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;

  const int k_iters_per_tap = prm.kc0 + prm.kc1;
  const int k_iters = prm.n_taps * k_iters_per_tap;
  const long long tiles_per_img = (long long)prm.n_phases * prm.tiles_y * prm.tiles_x * prm.n_tiles_n;
  const long long total_tiles = (long long)prm.n_img * tiles_per_img;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        long long r = tile;
        const int nt = (int)(r % prm.n_tiles_n); r /= prm.n_tiles_n;
        const int tx = (int)(r % prm.tiles_x); r /= prm.tiles_x;
        const int ty = (int)(r % prm.tiles_y); r /= prm.tiles_y;
        const int ph = (int)(r % prm.n_phases);
        const int img = (int)(r / prm.n_phases);
        const int x0 = tx * kTileW, y0 = ty * kTileH;
        for (int tap = 0; tap < prm.n_taps; ++tap) {
          const int xx = x0 + prm.dx[ph][tap], yy = y0 + prm.dy[ph][tap];
          for (int kc = 0; kc < k_iters_per_tap; ++kc) {
            mbar_wait_slack(bar_empty + 8 * stage, phase ^ 1u);
            mbar_expect_tx(bar_full + 8 * stage, S::kStageBytes);
            if (kc < prm.kc0) tma_load_4d(smem_a + stage * S::kABytes, &map_a0, bar_full + 8 * stage, kc * KC, xx, yy, img);
            else tma_load_4d(smem_a + stage * S::kABytes, &map_a1, bar_full + 8 * stage, (kc - prm.kc0) * KC, xx, yy, img);
            tma_load_3d(smem_b + stage * S::kBBytes, &map_w, bar_full + 8 * stage, kc * KC, nt * BLOCK_N, ph * prm.n_taps + tap);
            if (++stage == S::kStages) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc<BLOCK_N>();
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        mbar_wait(bar_tempty + 8 * acc, acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BLOCK_N);
        for (int it = 0; it < k_iters; ++it) {
          mbar_wait(bar_full + 8 * stage, phase);
          tc_fence_after();
          const uint64_t da = make_kmajor_desc<KC>(smem_a + stage * S::kABytes);
          const uint64_t db = make_kmajor_desc<KC>(smem_b + stage * S::kBBytes);
#pragma unroll
          for (int k = 0; k < KC / 16; ++k) {
            // advance 16 bf16 = 32 bytes along K inside the swizzle row: +2 in the (addr >> 4) field
            umma_bf16(tmem_d, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (it > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(bar_empty + 8 * stage);  // frees the smem slot once these MMAs have read it
          if (++stage == S::kStages) { stage = 0; phase ^= 1u; }
        }
        umma_commit(bar_tfull + 8 * acc);  // accumulator complete -> epilogue
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else {
    // ===================== epilogue (warps 2..5, 128 threads) =====================
    const int q = warp & 3;               // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;        // GEMM row = pixel inside the tile
    const int et = threadIdx.x - 64;      // 0..127 among epilogue threads
    int acc = 0;
    uint32_t acc_phase = 0;
    for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      long long r = tile;
      const int nt = (int)(r % prm.n_tiles_n); r /= prm.n_tiles_n;
      const int tx = (int)(r % prm.tiles_x); r /= prm.tiles_x;
      const int ty = (int)(r % prm.tiles_y); r /= prm.tiles_y;
      const int ph = (int)(r % prm.n_phases);
      const int img = (int)(r / prm.n_phases);

      // stage this tile's (scale, shift) pairs; double-buffered by accumulator stage
      float2* coef = s_coef + acc * BLOCK_N;
      for (int c = et; c < BLOCK_N; c += 128)
        coef[c] = __ldg(prm.coef + (long long)img * prm.coef_stride + prm.coef_off + nt * BLOCK_N + c);
      asm volatile("bar.sync 1, 128;" ::: "memory");

      mbar_wait_slack(bar_tfull + 8 * acc, acc_phase);
      tc_fence_after();

      const int y = ty * kTileH + (row >> 4), x = tx * kTileW + (row & 15);
      const bool valid = (y < prm.in_h) && (x < prm.in_w);
      const int oy = prm.out_mul * y + (ph >> 1), ox = prm.out_mul * x + (ph & 1);
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BLOCK_N);

      if (prm.head != nullptr) {
        // BLOCK_N == 32: conv_cls.0 epilogue + conv_cls.1 (1x1, 32 -> 2) + logits store
        uint32_t v[32];
        tmem_ld_32x32b_x32(taddr, v);
        tmem_ld_wait();
        float l0 = s_head[64], l1 = s_head[65];
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          const float2 cf = coef[c];
          float a = fmaf(__uint_as_float(v[c]), cf.x, cf.y);
          a = prm.relu ? fmaxf(a, 0.0f) : a;
          l0 = fmaf(a, s_head[c], l0);
          l1 = fmaf(a, s_head[32 + c], l1);
        }
        if (valid) {
          const int t = img / prm.chunk_slices, sl = img - t * prm.chunk_slices;
          const long long gimg = (long long)t * prm.n_slices_total + prm.slice0 + sl;
          if (prm.head_diff) {
            prm.logits[(gimg * prm.out_h + oy) * prm.out_w + ox] = l0 - l1;
          } else {
            float2* dst = reinterpret_cast<float2*>(prm.logits) + (gimg * prm.out_h + oy) * prm.out_w + ox;
            *dst = make_float2(l0, l1);
          }
        }
      } else {
        __nv_bfloat16* dst = prm.out + (long long)img * prm.out_img_stride + ((long long)oy * prm.out_w + ox) * prm.out_c + nt * BLOCK_N;
#pragma unroll 1
        for (int cb = 0; cb < BLOCK_N; cb += 32) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(taddr + (uint32_t)cb, v);
          tmem_ld_wait();
          uint32_t packed[16];
          if (prm.accumulate && valid) {
            const uint4* s4 = reinterpret_cast<const uint4*>(dst + cb);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const uint4 q = s4[i];
              packed[4 * i] = q.x; packed[4 * i + 1] = q.y; packed[4 * i + 2] = q.z; packed[4 * i + 3] = q.w;
            }
          }
#pragma unroll
          for (int c = 0; c < 32; c += 2) {
            const float2 c0 = coef[cb + c], c1 = coef[cb + c + 1];
            float a0 = fmaf(__uint_as_float(v[c]), c0.x, c0.y);
            float a1 = fmaf(__uint_as_float(v[c + 1]), c1.x, c1.y);
            if (prm.accumulate) {
              const float2 old = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&packed[c >> 1]));
              a0 += old.x;
              a1 += old.y;
            }
            if (prm.relu) { a0 = fmaxf(a0, 0.0f); a1 = fmaxf(a1, 0.0f); }
            __nv_bfloat162 b = __floats2bfloat162_rn(a0, a1);
            packed[c >> 1] = *reinterpret_cast<uint32_t*>(&b);
          }
          if (valid) {
            uint4* d4 = reinterpret_cast<uint4*>(dst + cb);
#pragma unroll
            for (int i = 0; i < 4; ++i) d4[i] = make_uint4(packed[4 * i], packed[4 * i + 1], packed[4 * i + 2], packed[4 * i + 3]);
          }
        }
      }
      tc_fence_before();
      mbar_arrive(bar_tempty + 8 * acc);
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, S::kTmemCols);
}

}  // namespace rcu
