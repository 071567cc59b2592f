// Wide-channel implicit-GEMM convolution for the deep levels (c_out a multiple of 128: the 60^2, 30^2 and 15^2 levels
// of the BraTS net), where weights no longer fit in shared memory and the per-tap kernel in conv_tc.cuh spends its
// time re-loading the activation window nine times through TMA.
//
//   unit of work  = 16 x 16 output pixels (two 16 x 8 sub-tiles = two M = 128 accumulators) x 128 output channels
//   activations   : ONE TMA box per (unit, 64-channel chunk): the 18 x 18 halo window, SWIZZLE_128B rows; the taps and
//                   the two sub-tiles are row-shifted descriptor views of it (see conv_halo.cuh)
//   weights       : streamed per (chunk, group of TPS taps) as pre-swizzled [128][64] tiles by plain bulk copies
//                   (cp.async.bulk has no per-row cost; a tiled TMA box costs ~3.5 clk per row); each weight tile
//                   feeds both sub-tiles, so L2 -> SM weight traffic is 16 KB per 512 MMA-clocks (32 B/clk/SM)
//   TMEM          : 2 accumulators x 128 columns per unit, double buffered (512 columns): the epilogue of unit i
//                   overlaps the MMAs of unit i+1
//   MMA           : N = 128 -> 64 clk per tcgen05.mma, the tensor-pipe floor (operand reads 128 B/clk, probe `rate`)
//
// Units are dealt round-robin with the output-channel tile slowest, so the SMs stream the same weight tiles at the
// same time (L2-hot) while each reads its own window.  The host picks chunk sizes that make the unit count a near
// multiple of the SM count.
#pragma once
#include "conv_halo.cuh"

namespace rcu {

constexpr int kWideTile = 16;                 // unit = 16 x 16 pixels
constexpr int kWidePitch = kWideTile + 2;     // window 18 x 18
constexpr int kWideN = 128;
constexpr uint32_t kWideABytes = kWidePitch * kWidePitch * 128;          // 41472
constexpr uint32_t kWideASlot = (kWideABytes + 1023u) & ~1023u;          // 41984
constexpr uint32_t kWideWTile = kWideN * 128;                            // 16 KB
constexpr int kWideThreads = 64 + 2 * 128;    // producer warp, MMA warp, two epilogue groups of four warps

struct WideParams {
  int n_img, in_h, in_w, tiles_x, tiles_y;
  int n_chunks;              // c_in / 64
  int n_ntiles;              // c_out / 128
  int n_phases;              // 1 (3x3 conv) or 4 (nearest-x2 + conv3x3 as 2x2 phase convs)
  const uint8_t* w_image;    // [phase][ntile][chunk][tap][128][64] pre-swizzled bf16 tiles
  int out_mul, out_h, out_w;
  int out_c;                 // pixel stride in elements
  long long out_img_stride;
  __nv_bfloat16* out;        // channel offset of the destination slice folded in
  const float2* coef;
  long long coef_stride;
  int coef_off;
  int relu;
  int tma_store;             // epilogue writes through swizzled smem staging + TMA tensor stores (see conv_halo.cuh)
  __nv_bfloat16* pool_out;   // when set (TAPS == 9): MaxPool2d(2) of the output rides in the epilogue (unet.py:92-95)
  long long pool_img_stride;
  int pool_c;                // pixel stride of the pooled tensor in elements
};

template <int TAPS>   // 9: conv3x3, 4: one up-path phase per unit
struct WideCfg {
#ifndef RCU_WIDE_TPS
#define RCU_WIDE_TPS 3
#endif
#ifndef RCU_WIDE_WSTAGES
#define RCU_WIDE_WSTAGES 2
#endif
  static constexpr int kTps = TAPS == 9 ? RCU_WIDE_TPS : 2;           // taps per weight stage
  static constexpr int kWg = TAPS / kTps + (TAPS % kTps ? 1 : 0);   // weight stages per chunk
  static constexpr uint32_t kWSlot = kTps * kWideWTile;    // 48 KB / 32 KB
  static constexpr int kAStages = 2;
  static constexpr int kWStages = TAPS == 9 ? RCU_WIDE_WSTAGES : 3;
  static constexpr int kCoefBytes = 2 * kWideN * (int)sizeof(float2);
  static constexpr int kBarBytes = (2 * kAStages + 2 * kWStages + 4) * 8 + 16;
  static constexpr int kOutSlot = 128 * 64;                // 128 pixels x 32 channels, bf16, one per epilogue group
  static constexpr int kSmem = 1024 + kAStages * kWideASlot + kWStages * kWSlot + 2 * kOutSlot + kCoefBytes + kBarBytes;
};

template <int TAPS>
__global__ void __launch_bounds__(kWideThreads, 1)
conv_wide_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ HaloOutMaps out_maps,
                 const __grid_constant__ WideParams prm) {
  using C = WideCfg<TAPS>;
  extern __shared__ uint8_t smem_dyn[];
  const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_dyn + (base - smem_u32(smem_dyn));
  const uint32_t smem_a = base;
  const uint32_t smem_w = base + C::kAStages * kWideASlot;
  const uint32_t smem_out = base + C::kAStages * kWideASlot + C::kWStages * C::kWSlot;   // 1024-aligned
  const uint32_t tail = C::kAStages * kWideASlot + C::kWStages * C::kWSlot + 2 * C::kOutSlot;
  float2* s_coef = reinterpret_cast<float2*>(base_ptr + tail);
  uint8_t* bar_ptr = base_ptr + tail + C::kCoefBytes;
  const uint32_t bar_afull = smem_u32(bar_ptr);
  const uint32_t bar_aempty = bar_afull + C::kAStages * 8;
  const uint32_t bar_wfull = bar_aempty + C::kAStages * 8;
  const uint32_t bar_wempty = bar_wfull + C::kWStages * 8;
  const uint32_t bar_tfull = bar_wempty + C::kWStages * 8;   // [2]
  const uint32_t bar_tempty = bar_tfull + 16;                 // [2]
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bar_ptr + (2 * C::kAStages + 2 * C::kWStages + 4) * 8);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) griddep_launch_dependents();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    for (int s = 0; s < C::kAStages; ++s) { mbar_init(bar_afull + 8 * s, 1); mbar_init(bar_aempty + 8 * s, 1); }
    for (int s = 0; s < C::kWStages; ++s) { mbar_init(bar_wfull + 8 * s, 1); mbar_init(bar_wempty + 8 * s, 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(bar_tfull + 8 * s, 1); mbar_init(bar_tempty + 8 * s, kEpilogueArrivals); }   // tempty: one arrival per epilogue warp (mbar_arrive_warp)
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(s_tmem), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;

  // unit -> (phase, ntile, img, ty, tx), tx fastest, (phase, ntile) slowest
  const int tiles_per_img = prm.tiles_y * prm.tiles_x;
  const int units_per_w = prm.n_img * tiles_per_img;                       // units sharing one weight stream
  const int total_units = units_per_w * prm.n_ntiles * prm.n_phases;

  if (warp == 0) {
    // ===================== producer =====================
    if (lane == 0) {
      griddep_wait();   // see conv_halo.cuh: everything above overlapped the previous layer's tail
      int as = 0, ws = 0;
      uint32_t aph = 0, wph = 0;
      for (int unit = blockIdx.x; unit < total_units; unit += gridDim.x) {
        const int wsel = unit / units_per_w;                 // phase * n_ntiles + ntile
        int r = unit - wsel * units_per_w;
        const int img = r / tiles_per_img; r -= img * tiles_per_img;
        const int ty = r / prm.tiles_x, tx = r - ty * prm.tiles_x;
        const uint8_t* wsrc = prm.w_image + (size_t)wsel * prm.n_chunks * TAPS * kWideWTile;
        for (int c = 0; c < prm.n_chunks; ++c) {
          mbar_wait_slack(bar_aempty + 8 * as, aph ^ 1u);
          mbar_expect_tx(bar_afull + 8 * as, kWideABytes);
          tma_load_4d(smem_a + (uint32_t)as * kWideASlot, &map_a, bar_afull + 8 * as, c * 64, tx * kWideTile - 1, ty * kWideTile - 1, img);
          if (++as == C::kAStages) { as = 0; aph ^= 1u; }
#pragma unroll
          for (int g = 0; g < C::kWg; ++g) {
            const int taps = (g + 1) * C::kTps <= TAPS ? C::kTps : TAPS - g * C::kTps;
            mbar_wait_slack(bar_wempty + 8 * ws, wph ^ 1u);
            mbar_expect_tx(bar_wfull + 8 * ws, (uint32_t)taps * kWideWTile);
            for (int t = 0; t < taps; ++t)
              bulk_load(smem_w + (uint32_t)ws * C::kWSlot + (uint32_t)t * kWideWTile,
                        wsrc + ((size_t)c * TAPS + g * C::kTps + t) * kWideWTile, kWideWTile, bar_wfull + 8 * ws);
            if (++ws == C::kWStages) { ws = 0; wph ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (warp-uniform flow, elect-predicated tcgen05) =====================
    const bool leader = elect_one() != 0;
    constexpr uint32_t idesc = make_idesc<kWideN>();
    constexpr uint32_t hi_a = (uint32_t)((kWidePitch * 128) >> 4) | (1u << 14) | (2u << 29);
    constexpr uint32_t hi_b = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);
    int as = 0, ws = 0, acc = 0;
    uint32_t aph = 0, wph = 0, acc_phase = 0;
    for (int unit = blockIdx.x; unit < total_units; unit += gridDim.x) {
      const int wsel = unit / units_per_w;
      const int ph = wsel / prm.n_ntiles;
      // up-path phase (a, b): taps (i, j) read window rows a + i, columns b + j (conv3x3: rows ty, columns tx)
      const uint32_t up_base16 = TAPS == 4 ? (uint32_t)((((ph >> 1) * kWidePitch) + (ph & 1)) * 8) : 0u;
      mbar_wait(bar_tempty + 8 * acc, acc_phase ^ 1u);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(acc * 2 * kWideN);
      for (int c = 0; c < prm.n_chunks; ++c) {
        mbar_wait(bar_afull + 8 * as, aph);
        tc_fence_after();
        const uint32_t lo_a0 = ((((smem_a + (uint32_t)as * kWideASlot) & 0x3FFFFu) >> 4) | (1u << 16)) + up_base16;
#pragma unroll
        for (int g = 0; g < C::kWg; ++g) {
          mbar_wait(bar_wfull + 8 * ws, wph);
          tc_fence_after();
          const uint32_t lo_b0 = (((smem_w + (uint32_t)ws * C::kWSlot) & 0x3FFFFu) >> 4) | (1u << 16);
#pragma unroll
          for (int t = 0; t < C::kTps; ++t) {
            const int tap = g * C::kTps + t;
            if (tap < TAPS) {
              const uint32_t a_off = TAPS == 4 ? (uint32_t)(((tap >> 1) * kWidePitch + (tap & 1)) * 8)
                                               : (uint32_t)(((tap / 3) * kWidePitch + tap % 3) * 8);
#pragma unroll
              for (int s = 0; s < 2; ++s) {
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                  const uint32_t accumulate = (tap > 0 || ks > 0) ? 1u : (c > 0 ? 1u : 0u);
                  if (leader)
                    umma_bf16(tmem_d + (uint32_t)(s * kWideN), desc_from(lo_a0 + a_off + (uint32_t)(s * 8 * 8) + 2 * ks, hi_a),
                              desc_from(lo_b0 + (uint32_t)t * (kWideWTile >> 4) + 2 * ks, hi_b), idesc, accumulate);
                }
              }
            }
          }
          if (leader) umma_commit(bar_wempty + 8 * ws);
          if (++ws == C::kWStages) { ws = 0; wph ^= 1u; }
        }
        if (leader) umma_commit(bar_aempty + 8 * as);
        if (++as == C::kAStages) { as = 0; aph ^= 1u; }
      }
      if (leader) umma_commit(bar_tfull + 8 * acc);
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
  } else {
    // ===================== epilogue: two groups of four warps, group g drains accumulator stage g =====================
    const int group = (warp - 2) >> 2;
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int gt = threadIdx.x - 64 - group * 128;
    float2* coef = s_coef + group * kWideN;
    uint32_t acc_phase = 0;
    int cur_key = -1;
    int it = 0;
    for (int unit = blockIdx.x; unit < total_units; unit += gridDim.x, ++it) {
      if ((it & 1) != group) continue;
      const int wsel = unit / units_per_w;
      const int ph = wsel / prm.n_ntiles, nt = wsel - ph * prm.n_ntiles;
      int r = unit - wsel * units_per_w;
      const int img = r / tiles_per_img; r -= img * tiles_per_img;
      const int ty = r / prm.tiles_x, tx = r - ty * prm.tiles_x;

      const int key = img * prm.n_ntiles + nt;
      if (key != cur_key) {   // group-uniform
        asm volatile("bar.sync %0, 128;" ::"r"(group + 1) : "memory");
        coef[gt] = __ldg(prm.coef + (long long)img * prm.coef_stride + prm.coef_off + nt * kWideN + gt);
        asm volatile("bar.sync %0, 128;" ::"r"(group + 1) : "memory");
        cur_key = key;
      }
      mbar_wait_slack(bar_tfull + 8 * group, acc_phase);
      tc_fence_after();
#pragma unroll 1
      for (int s = 0; s < 2; ++s) {
        const int y = ty * kWideTile + (row >> 3), x = tx * kWideTile + s * 8 + (row & 7);
        const bool valid = (y < prm.in_h) && (x < prm.in_w);
        const int oy = prm.out_mul * y + (TAPS == 4 ? (ph >> 1) : 0), ox = prm.out_mul * x + (TAPS == 4 ? (ph & 1) : 0);
        __nv_bfloat16* dst = prm.out + (long long)img * prm.out_img_stride + ((long long)oy * prm.out_w + ox) * prm.out_c + nt * kWideN;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(group * 2 * kWideN + s * kWideN);
#pragma unroll 1
        for (int cb = 0; cb < kWideN; cb += 32) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(taddr + (uint32_t)cb, v);
          tmem_ld_wait();
          if (s == 1 && cb + 32 == kWideN) {
            tc_fence_before();
            mbar_arrive_warp(bar_tempty + 8 * group);   // both accumulators are in registers / stored: release the stage
          }
          uint32_t packed[16];
#pragma unroll
          for (int c = 0; c < 32; c += 2) {
            RCU_COEF2(cb + c);
            float a0 = fmaf(__uint_as_float(v[c]), c0.x, c0.y);
            float a1 = fmaf(__uint_as_float(v[c + 1]), c1.x, c1.y);
            if (prm.relu) { a0 = fmaxf(a0, 0.0f); a1 = fmaxf(a1, 0.0f); }
            __nv_bfloat162 b = __floats2bfloat162_rn(a0, a1);
            packed[c >> 1] = *reinterpret_cast<uint32_t*>(&b);
          }
          if (prm.tma_store) {
            const uint32_t so = smem_out + (uint32_t)group * C::kOutSlot;
            if (gt == 0) bulk_wait_group_read0();
            asm volatile("bar.sync %0, 128;" ::"r"(group + 1) : "memory");
            const uint32_t rowa = so + (uint32_t)row * 64u;
            const uint32_t sw = (uint32_t)(row >> 1) & 3u;
#pragma unroll
            for (int i = 0; i < 4; ++i)
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rowa + (((uint32_t)i ^ sw) << 4)), "r"(packed[4 * i]),
                           "r"(packed[4 * i + 1]), "r"(packed[4 * i + 2]), "r"(packed[4 * i + 3])
                           : "memory");
            fence_proxy_async();
            asm volatile("bar.sync %0, 128;" ::"r"(group + 1) : "memory");
            if (gt == 0) {
              tma_store_4d(&out_maps.m[TAPS == 4 ? ph : 0], so, nt * kWideN + cb, tx * kWideTile + s * 8, ty * kWideTile, img);
              bulk_commit_group();
            }
          } else if (valid) {
            uint4* d4 = reinterpret_cast<uint4*>(dst + cb);
#pragma unroll
            for (int i = 0; i < 4; ++i) d4[i] = make_uint4(packed[4 * i], packed[4 * i + 1], packed[4 * i + 2], packed[4 * i + 3]);
          }
          if (TAPS == 9 && prm.pool_out != nullptr) {
            // 2x2 max over the quad (y^1, x^1): row = y * 8 + x within the accumulator, a warp holds four y rows, so the
            // partners are lanes ^1 and ^8.  h, w are even: out-of-image pixels only pair with out-of-image pixels.
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              __nv_bfloat162 m = *reinterpret_cast<__nv_bfloat162*>(&packed[i]);
              uint32_t o = __shfl_xor_sync(0xffffffffu, packed[i], 1);
              m = __hmax2(m, *reinterpret_cast<__nv_bfloat162*>(&o));
              uint32_t mm = *reinterpret_cast<uint32_t*>(&m);
              o = __shfl_xor_sync(0xffffffffu, mm, 8);
              m = __hmax2(m, *reinterpret_cast<__nv_bfloat162*>(&o));
              packed[i] = *reinterpret_cast<uint32_t*>(&m);
            }
            if (valid) {
              // the four lanes of a quad hold the same maxima: each stores one 16-byte quarter of the 32 channels
              const int part = ((y & 1) << 1) | (x & 1);
              __nv_bfloat16* pd = prm.pool_out + (long long)img * prm.pool_img_stride +
                                  ((long long)(y >> 1) * (prm.in_w >> 1) + (x >> 1)) * prm.pool_c + nt * kWideN + cb + part * 8;
              uint4 v4;
              v4.x = part == 0 ? packed[0] : part == 1 ? packed[4] : part == 2 ? packed[8] : packed[12];
              v4.y = part == 0 ? packed[1] : part == 1 ? packed[5] : part == 2 ? packed[9] : packed[13];
              v4.z = part == 0 ? packed[2] : part == 1 ? packed[6] : part == 2 ? packed[10] : packed[14];
              v4.w = part == 0 ? packed[3] : part == 1 ? packed[7] : part == 2 ? packed[11] : packed[15];
              *reinterpret_cast<uint4*>(pd) = v4;
            }
          }
        }
      }
      acc_phase ^= 1u;
    }
    if (prm.tma_store && gt == 0) bulk_wait_group_read0();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

}  // namespace rcu
