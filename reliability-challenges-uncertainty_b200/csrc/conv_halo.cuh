// Halo-tile implicit-GEMM convolution for the high-resolution, thin layers (c_out = 32 / 64, the 240^2 and 120^2
// levels of the BraTS net, where the per-tap kernel in conv_tc.cuh is bound by L2 -> shared-memory traffic).
//
//   * one TMA box per (tile, 64-channel chunk) brings the (16+2) x (8+2) pixel halo window into shared memory ONCE;
//     the 9 taps (4 for an up-path phase) are shifted VIEWS of that window: the tcgen05 shared-memory descriptor of
//     tap (dy, dx) is the window's descriptor with the start address advanced by ((dy+1) * 10 + (dx+1)) rows.
//     SWIZZLE_128B is a function of the absolute shared-memory address, so a row-shifted start is still a valid
//     K-major operand (checked on B200 by tools/ubench/umma_probe.cu, results in profiles/r01_umma_probe.log);
//   * GEMM row r of a tile is pixel (y, x) = (r / 8, r % 8): every 8-row core-matrix group is 8 x-adjacent pixels,
//     i.e. 8 consecutive 128-byte rows of the window, and consecutive groups are one window row (10 pixels) apart —
//     a uniform stride-byte-offset;
//   * 32-channel sources (64-byte pixel rows) use the same SWIZZLE_128B map with a 32-channel box: TMA then still
//     gives every pixel its own 128-byte swizzled row and fills its first 64 logical bytes (measured, probe `tma5d`),
//     so the descriptors are unchanged and a tap is two K = 16 steps instead of four — full-rate operand reads,
//     where SWIZZLE_64B rows would read at half rate (2-way bank conflicts, probe `rate`);
//   * the layer's weights (<= 144 KB as pre-swizzled [c_out][64] tiles) are copied into shared memory once per CTA
//     and stay resident while the CTA walks its tiles (persistent kernel, one CTA per SM);
//   * the MMA warps run warp-uniform control flow and predicate only the tcgen05.mma on elect.sync, so descriptors
//     live in uniform registers (N = 32/64 tiles are 40/48-clock MMAs, bound by the 128 B/clk shared-memory operand
//     read — see the probe).  The MMA sequence of a tile is ROLLED over the window rows (#pragma unroll 1): fully
//     unrolled, ptxas hoists every descriptor computation of the tile in front of its first UTCHMMA and spills 90 - 114
//     uniform registers into the vector registers the epilogue needs (profiles/r02zz_ab_experiments.log); the head
//     instantiation alone keeps the unrolled form, which measured faster for it;
//   * one kernel instantiation per role (HaloMode): pixel rows over 64- / 32-channel sources, up-path phases, and the
//     pixel-pair rows as four instantiations — plain, with the patch warps of the first-layer dedup, with the fused head
//     (store path compiled out), and over a 64-channel source.
//
// Epilogue contract is that of conv_tc.cuh (folded Dropout2d x BN coefficients + ReLU, bf16 NHWC store with a
// channel offset / pixel stride so producers write straight into the concat buffer, or the fused 1x1 head).
#pragma once
#include "conv_tc.cuh"

namespace rcu {

constexpr int kHaloTileH = 16;
constexpr int kHaloTileW = 8;
constexpr int kHaloPitch = kHaloTileW + 2;   // window pixels per row
constexpr int kHaloRows = kHaloTileH + 2;    // window rows
// 3x3 conv over 64-ch chunks / a 32-ch source / one up-path phase / pixel-pair rows over a 32-ch / a 64-ch source
enum HaloMode { HALO_CONV64 = 0, HALO_CONV32 = 1, HALO_UP64 = 2, HALO_PAIR32 = 3, HALO_PAIR64 = 4, HALO_PAIR32_PATCH = 5, HALO_PAIR32_HEAD = 6 };   // 5: HALO_PAIR32 + the patch warps of the first-layer dedup; 6: HALO_PAIR32 with the fused 1x1 head as its epilogue
// Pixel-pair formulation (c_out = 32 layers).  An N = 32 tcgen05.mma reads 4 KB of A and 1 KB of B from shared memory
// for 128 x 32 x 16 MACs: the 128 B/clk operand port, not the tensor array, bounds it (40 clk against a 16 clk floor,
// profiles/r01_umma_probe.log).  A GEMM row that holds TWO x-adjacent pixels (the dense NHWC tensor viewed as
// [h][w/2][2c]) doubles N for the same A bytes: output column (a_o, c_o) of pair i sums taps dx = 2*di + a_in - a_o over
// the pairs di = -1, 0, +1, so
//     di =  0 : every (a_in, a_o) combination is a tap of the 3x3 kernel   -> dense [64 x K] weight tile, N = 64
//     di = -1 : only a_in = 1 -> a_o = 0 (dx = -1)                        -> N = 32 into columns  0..31, K = the a_in = 1 half
//     di = +1 : only a_in = 0 -> a_o = 1 (dx = +1)                        -> N = 32 into columns 32..63, K = the a_in = 0 half
// No zero blocks are multiplied (executed MACs = algorithmic MACs); per 256 output pixels and 3x3 x 32 channels the MMAs
// read 3 x (4 x 6 KB + 4 x 5 KB) = 132 KB instead of 2 x 18 x 5 KB = 180 KB, and a 32-channel halo box has no padding.
#ifndef RCU_UP_PAIRED
#define RCU_UP_PAIRED 1      // A/B switch: 0 runs the up-path phases one by one (and packs their weights per phase)
#endif
constexpr bool halo_is_pair32(int mode) { return mode == HALO_PAIR32 || mode == HALO_PAIR32_PATCH || mode == HALO_PAIR32_HEAD; }
constexpr bool halo_is_pair(int mode) { return halo_is_pair32(mode) || mode == HALO_PAIR64; }
#ifndef RCU_HALO_PATCH_WARPS
#define RCU_HALO_PATCH_WARPS 4
#endif
#ifndef RCU_PAIR_CENTRE_FIRST
#define RCU_PAIR_CENTRE_FIRST 1  // pixel-pair MMA sequence: the 12 centre (N = 64) MMAs of a chunk first, then its 12 side (N = 32) MMAs (0: per window row)
#endif
#ifndef RCU_HALO_ROLLED
#define RCU_HALO_ROLLED 1   // pixel-row and up-path MMA sequences rolled over window rows as well (A/B)
#endif
#ifndef RCU_PAIR_ROLLED
#define RCU_PAIR_ROLLED 1
#endif
#ifndef RCU_HALO_PATCH_SPLIT
#define RCU_HALO_PATCH_SPLIT 0   // 1: every patch warp takes a share of the rows of EVERY tile (shortest hand-over); 0: whole tiles, one stage per warp
#endif
constexpr int kHaloPatchWarps = RCU_HALO_PATCH_WARPS;   // warps that patch dropped first-layer channels into landed tiles, stage s by warp s mod n
constexpr int halo_patch_threads(int mode) { return mode == HALO_PAIR32_PATCH ? 32 * kHaloPatchWarps : 0; }
constexpr int kHaloSmemBudget = 225 * 1024;

struct HaloParams {
  int n_img, in_h, in_w, tiles_x, tiles_y;
  int n_chunks;              // TMA boxes (pipeline slots) per tile: c_in / 64, or 1 for a 32-channel source
  uint32_t chunk_bytes;      // bytes one box delivers
  uint32_t chunk_stride;     // slot size (multiple of 1024)
  int n_stages;
  int tma_store;             // epilogue stages bf16 tiles in shared memory and writes them with TMA tensor stores
  int tiles_per_turn;        // consecutive tiles an MMA issuer handles per issue turn (1 or 2)
  // HALO_UP64: the PH up-path phases this launch computes per tile (one accumulator each).  Phase (a, b) reads the
  // window from (a * 10 + b) * 8 sixteen-byte units on and writes output pixel (2y + a, 2x + b).
  uint32_t up_base16[4];
  int up_dy[4], up_dx[4];
  const void* w_image;       // pre-swizzled weight tiles in global memory, in the order the MMA loop walks them
  uint32_t w_bytes;
  // output addressing: pixel (y, x) lands at (out_mul*y + dy, out_mul*x + dx), (dy, dx) = the phase offset (0 for plain convs)
  int out_mul, out_h, out_w;
  int out_c;                 // pixel stride in elements (the channel offset is folded into `out`)
  long long out_img_stride;  // elements between images
  __nv_bfloat16* out;
  // optional fused nn.MaxPool2d(2) (common/model/unet.py:90) of this layer's output: dense [img][h/2][w/2][N] tensor
  __nv_bfloat16* pool_out;
  long long pool_img_stride;
  const float2* coef;
  long long coef_stride;
  int coef_off;
  int relu;
  const float* head;         // non-NULL: fused 1x1 head (conv_cls.1 / conv_sigma.1); the weights themselves travel in head_w
  float* logits;
  int head_diff;             // 1: the head stores l0 - l1 (float32 per pixel, softmax is a function of the difference alone) instead of the logit pair
  int chunk_slices;
  long long slice0, n_slices_total;
  // HALO_PAIR32 reading the FIRST convolution's output stored once per slice instead of once per (sample, slice) (unet.cu):
  //   0 off; 1 every image reads variant 0 (no dropout in that unit); 2 sample 0 reads variant 0 (the deterministic pass), the
  //   others variant 1; 3 all read variant 1.  Variant 1 holds every channel as if kept; the channels Dropout2d dropped for an
  //   image are constants (relu of the folded bias) that the patch warp writes into the landed tile before the MMAs read it.
  int dedup_mode;
  const float2* patch_coef;  // coefficient rows of the producing unit, [image][coef_stride] + patch_off: x == 0 marks a dropped channel, relu(y) is its value
  int patch_off;
  // [2][32] weights + [2] bias of the fused head, in the kernel's constant bank: every FFMA of the head takes its weight
  // as a constant operand (the shared-memory copy cost three LDS per channel on the port the tensor pipe reads its operands from)
  float head_w[66];
};

__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred;
}

// Programmatic dependent launch: a kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may become
// resident while its predecessor in the stream drains.  Everything before griddep_wait() (barrier init, TMEM allocation,
// descriptor prefetch, the resident weight image: constant data) overlaps the predecessor's tail; the first load of
// an activation comes after it, and every global store of the kernel is downstream of those loads.
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}

struct HaloOutMaps { CUtensorMap m[4]; };   // destination views, one per up-path phase of the launch (m[0] for plain convs)

__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(reinterpret_cast<uint64_t>(map)),
               "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_group_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

__device__ __forceinline__ uint64_t desc_from(uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | lo; }

template <int N, int G, int PH = 1>
struct HaloSmem {
  static constexpr int kCoefBytes = G * N * (int)sizeof(float2);
  static constexpr int kHeadBytes = 2 * 32 * 4 + 16;
  static constexpr int kOutSlot = 128 * 64;                    // one 128-pixel x 32-channel bf16 tile per epilogue group
  static constexpr int kOutBytes = G * kOutSlot;               // only reserved when the launch uses TMA stores
  static constexpr int kMaxStages = 8;
  static constexpr int kBarBytes = (2 * kMaxStages + 2 * G + 3) * 8 + 16 + kMaxStages * 8 + 64 * kHaloPatchWarps;   // ... + landed[kMaxStages] + 32 bf16 patch values per patch warp
  static constexpr int kFixed = 1024 + kCoefBytes + kHeadBytes + kBarBytes;
  static constexpr int kAccCols = PH * N;                      // TMEM columns of one accumulator stage
  static constexpr int kTmemCols = G * kAccCols < 32 ? 32 : G * kAccCols;   // 128, 256 or 512: powers of two
  static constexpr int kThreads = 96 + 128 * G;   // producer warp, two MMA warps, G epilogue groups of four warps
};

// G = number of TMEM accumulator stages = number of 4-warp epilogue groups (group g drains the tiles whose index in
// the CTA's range is g mod G), so G epilogues are in flight while the MMA warp works on the next tile.
template <int N, int G, int MODE, int PH = 1>
__global__ void __launch_bounds__(96 + 128 * G + halo_patch_threads(MODE), 1)
conv_halo_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ HaloOutMaps out_maps,
                 const __grid_constant__ HaloParams prm) {
  using S = HaloSmem<N, G, PH>;
  static_assert(PH == 1 || MODE == HALO_UP64, "several phases per tile only exist on the up path");
  static_assert(S::kTmemCols <= 512, "accumulator stages exceed TMEM");
  static_assert(N == 32 || N == 64, "halo kernel serves c_out = 32 / 64");
  static_assert(!halo_is_pair(MODE) || (N == 64 && PH == 1), "pair rows carry two 32-channel output pixels");
  constexpr bool PAIR = halo_is_pair(MODE);
  constexpr bool PATCH = halo_patch_threads(MODE) != 0;
  constexpr int kEpiWarp0 = 3 + (PATCH ? kHaloPatchWarps : 0);   // first epilogue warp
  extern __shared__ uint8_t smem_dyn[];
  const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_dyn + (base - smem_u32(smem_dyn));
  const uint32_t w_span = (prm.w_bytes + 1023u) & ~1023u;
  const uint32_t smem_w = base;
  const uint32_t smem_a = base + w_span;
  const uint32_t smem_out = base + w_span + (uint32_t)prm.n_stages * prm.chunk_stride;   // 1024-aligned
  const uint32_t tail = w_span + (uint32_t)prm.n_stages * prm.chunk_stride + (prm.tma_store ? (uint32_t)S::kOutBytes : 0u);
  float2* s_coef = reinterpret_cast<float2*>(base_ptr + tail);
  uint8_t* bar_ptr = base_ptr + tail + S::kCoefBytes + S::kHeadBytes;
  const uint32_t bar_full = smem_u32(bar_ptr);                      // [kMaxStages]
  const uint32_t bar_empty = bar_full + S::kMaxStages * 8;          // [kMaxStages]
  const uint32_t bar_tfull = bar_empty + S::kMaxStages * 8;         // [G]
  const uint32_t bar_tempty = bar_tfull + G * 8;                    // [G]
  const uint32_t bar_w = bar_tempty + G * 8;                        // [1]
  const uint32_t bar_turn = bar_w + 8;                              // [2] issue-turn hand-off between the two MMA warps
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bar_ptr + (2 * S::kMaxStages + 2 * G + 3) * 8);
  const uint32_t bar_land = bar_turn + 16 + 16;                     // [kMaxStages] TMA landing barriers of the patch path
  unsigned short* s_patch = reinterpret_cast<unsigned short*>(bar_ptr + (2 * S::kMaxStages + 2 * G + 3) * 8 + 16 + S::kMaxStages * 8);   // [32]
  const bool patching = PATCH && prm.dedup_mode >= 2;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) griddep_launch_dependents();   // the next layer's CTA may take this SM as soon as this one exits
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    if (prm.tma_store)
      for (int i = 0; i < PH; ++i) tma_prefetch_desc(&out_maps.m[i]);
    for (int s = 0; s < S::kMaxStages; ++s) {
      mbar_init(bar_full + 8 * s, (patching && RCU_HALO_PATCH_SPLIT) ? kHaloPatchWarps : 1);   // split patching: one arrival per patch warp
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int s = 0; s < G; ++s) {
      mbar_init(bar_tfull + 8 * s, 1);
      mbar_init(bar_tempty + 8 * s, kEpilogueArrivals);   // one arrival per epilogue warp (mbar_arrive_warp)
    }
    mbar_init(bar_w, 1);
    mbar_init(bar_turn, 1);
    mbar_init(bar_turn + 8, 1);
    if (PATCH)
      for (int s = 0; s < S::kMaxStages; ++s) mbar_init(bar_land + 8 * s, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(s_tmem), S::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;

  // contiguous tile range of this CTA: neighbouring tiles share halo rows in L2 and, above all, the same image, so the
  // per-image epilogue coefficients are re-staged only a handful of times per launch
  const int tiles_per_img = prm.tiles_y * prm.tiles_x;
  const long long total_tiles = (long long)prm.n_img * tiles_per_img;   // < 2^31 (checked on the host)
  const int t_begin = (int)(total_tiles * blockIdx.x / gridDim.x);
  const int t_end = (int)(total_tiles * (blockIdx.x + 1) / gridDim.x);

  if (warp == 0) {
    // ===================== producer: weights once, then one halo box per (tile, chunk) =====================
    if (lane == 0) {
      mbar_expect_tx(bar_w, prm.w_bytes);
      for (uint32_t off = 0; off < prm.w_bytes; off += 32768u) {
        const uint32_t n = prm.w_bytes - off < 32768u ? prm.w_bytes - off : 32768u;
        bulk_load(smem_w + off, reinterpret_cast<const uint8_t*>(prm.w_image) + off, n, bar_w);
      }
      griddep_wait();                                    // the previous layer's output is complete and visible from here on
      int stage = 0;
      uint32_t phase = 0;
      int img = t_begin / tiles_per_img;
      int ty = (t_begin - img * tiles_per_img) / prm.tiles_x;
      int tx = t_begin - img * tiles_per_img - ty * prm.tiles_x;
      for (int tile = t_begin; tile < t_end; ++tile) {
        const int x0 = tx * kHaloTileW - 1, y0 = ty * kHaloTileH;
        int src_img = img;
        if (PATCH && prm.dedup_mode != 0) {   // the source holds [variant][slice of the chunk], not [sample][slice]
          const int t = img / prm.chunk_slices, sl = img - t * prm.chunk_slices;
          const int variant = prm.dedup_mode == 1 ? 0 : ((prm.dedup_mode == 2 && t == 0) ? 0 : 1);
          src_img = variant * prm.chunk_slices + sl;
        }
        for (int j = 0; j < prm.n_chunks; ++j) {
          mbar_wait_slack(bar_empty + 8 * stage, phase ^ 1u);
          const uint32_t dst = smem_a + (uint32_t)stage * prm.chunk_stride;
          const uint32_t bar_tx = (patching ? bar_land : bar_full) + 8 * stage;   // patched tiles reach the MMA warps through the patch warp
          mbar_expect_tx(bar_tx, prm.chunk_bytes);
          tma_load_4d(dst, &map_a, bar_tx, j * 64, x0, y0 - 1, src_img);
          if (++stage == prm.n_stages) { stage = 0; phase ^= 1u; }
        }
        if (++tx == prm.tiles_x) { tx = 0; if (++ty == prm.tiles_y) { ty = 0; ++img; } }
      }
    }
  } else if (warp <= 2) {
    // ===================== two MMA issuers (warps 1, 2), alternating tiles =====================
    // One warp's per-tile fixed cost (two mbarrier waits at ~90 clk each even when complete, two commits, loop
    // overhead: ~500 clk measured) is as long as the MMAs of a thin tile (18 x 40 clk) and the tensor pipe's queue is
    // shallow, so a single issuer leaves the pipe idle half of the time; with two issuers one warp's bookkeeping
    // hides behind the other's MMAs.  Control flow is warp-uniform, only the tcgen05 instructions are predicated on
    // elect.sync, so descriptors live in uniform registers and MMAs issue back to back.
#ifndef RCU_HALO_ISSUERS
#define RCU_HALO_ISSUERS 2
#endif
    constexpr int n_issuers = RCU_HALO_ISSUERS;   // 1: warp 1 alone issues every tile (A/B)
    const int mw = warp - 1;
    const bool leader = elect_one() != 0;
    constexpr uint32_t idesc = make_idesc<N>();
    // descriptor halves: [0,14) addr>>4, [16,30) LBO>>4 (=1, unused), hi: [0,14) SBO>>4, [14,16) version 1, [29,32) SWIZZLE_128B
    constexpr uint32_t hi_a = (uint32_t)((kHaloPitch * 128) >> 4) | (1u << 14) | (2u << 29);
    constexpr uint32_t hi_b = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);
    const uint32_t lo_b0 = ((smem_w & 0x3FFFFu) >> 4) | (1u << 16);
    mbar_wait(bar_w, 0);
    tc_fence_after();
    constexpr int kTaps = MODE == HALO_UP64 ? 4 : 9;
    constexpr int kK16 = MODE == HALO_CONV32 ? 2 : 4;
    constexpr uint32_t kTile16 = (uint32_t)(N * 128) >> 4;              // one [N][64] weight tile in 16-byte units
    constexpr uint32_t kChunkW16 = (MODE == HALO_CONV32 ? 5u : (uint32_t)kTaps) * kTile16;
    // The two warps never issue at the same time (concurrent tcgen05.mma streams from two warps of one CTA fault
    // intermittently on B200): an issuer owns the issue turn for TT consecutive tiles and hands it over right after
    // its last MMA; its commits and the barrier waits for its NEXT turn then overlap the other warp's MMAs.  The
    // hand-off itself costs ~200 clk of idle tensor pipe, so TT = 2 when the smem ring is deep enough.
    const int TT = prm.tiles_per_turn;
    const int n_tiles = t_end - t_begin;
    uint32_t turn_phase = 0;
    // ring positions of this issuer's next tile, advanced incrementally (no divisions on the issue path)
    int stage = (mw * TT * prm.n_chunks) % prm.n_stages;
    uint32_t phase = (uint32_t)((mw * TT * prm.n_chunks) / prm.n_stages) & 1u;
    int acc = (mw * TT) % G;
    uint32_t acc_phase = (uint32_t)((mw * TT) / G) & 1u;
    for (int k0 = mw * TT; k0 < n_tiles && mw < n_issuers; k0 += n_issuers * TT) {
      const int nt = n_tiles - k0 < TT ? n_tiles - k0 : TT;
      // everything this turn needs, before asking for the turn
      {
        int st = stage, ac = acc;
        uint32_t ph = phase, aph = acc_phase;
        for (int i = 0; i < nt; ++i) {
          mbar_wait(bar_tempty + 8 * ac, aph ^ 1u);
          mbar_wait(bar_full + 8 * st, ph);
          st += prm.n_chunks;
          if (st >= prm.n_stages) { st -= prm.n_stages; ph ^= 1u; }
          if (++ac == G) { ac = 0; aph ^= 1u; }
        }
      }
      if (n_issuers == 2 && k0 != 0) {
        mbar_wait(bar_turn + 8 * mw, turn_phase);
        turn_phase ^= 1u;
      }
      tc_fence_after();
      for (int i = 0; i < nt; ++i) {
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * S::kAccCols);
        for (int j = 0; j < prm.n_chunks; ++j) {
          if (j > 0) {
            mbar_wait(bar_full + 8 * stage, phase);
            tc_fence_after();
          }
          const uint32_t lo_a = (((smem_a + (uint32_t)stage * prm.chunk_stride) & 0x3FFFFu) >> 4) | (1u << 16);
          // every offset below is a compile-time constant: the loops unroll into back-to-back UTCHMMA with uniform adds
          if constexpr (PAIR) {
            // weight tiles [chunk][dy]{T0: [64][64] centre, S: [32][64] side} = 768 sixteen-byte units per (chunk, dy)
            constexpr uint32_t idesc32 = make_idesc<32>();
            auto pair_taps = [&](const int jj) {
              if constexpr (RCU_PAIR_ROLLED && MODE != HALO_PAIR32_HEAD) {
              // ROLLED over the three window rows (#pragma unroll 1): with the 24 MMAs of a chunk fully unrolled ptxas hoists every
              // descriptor computation in front of the first UTCHMMA (~130 uniform instructions, half of them uniform-register
              // spills: 96 live descriptor registers against 63) and the tensor pipe idles through that preamble at the start of
              // every issue turn.  A rolled row loop bounds the hoisting window to four MMAs; the next row's descriptor math then
              // runs while the queued MMAs execute.  Centre MMAs (N = 64) of the chunk first, then its side MMAs (N = 32).
              // Measured: 64->32 -2.5 %, the patching kernel -7 % (its 80-register budget spilled), 32->32 unchanged and conv_cls.0 + head
              // 7 % SLOWER (its light epilogue leaves the issue loop exposed): the head instantiation keeps the unrolled sequence.
              {
                uint32_t a_row = lo_a + 8u, b_t0 = lo_b0 + (uint32_t)(jj * 3 * 768);
#pragma unroll 1
                for (int dyi = 0; dyi < 3; ++dyi, a_row += (uint32_t)(kHaloPitch * 8), b_t0 += 768u) {
#pragma unroll
                  for (int ks = 0; ks < 4; ++ks)
                    if (leader)
                      umma_bf16(tmem_d, desc_from(a_row + 2 * ks, hi_a), desc_from(b_t0 + 2 * ks, hi_b), idesc, (jj > 0 || dyi > 0 || ks > 0) ? 1u : 0u);
                }
              }
              {
                uint32_t a_row = lo_a, b_s = lo_b0 + (uint32_t)(jj * 3 * 768) + 512u;
#pragma unroll 1
                for (int dyi = 0; dyi < 3; ++dyi, a_row += (uint32_t)(kHaloPitch * 8), b_s += 768u) {
                  if (halo_is_pair32(MODE)) {
                    // one chunk holds both pixels of the pair: K 32..63 is a_in = 1 (left neighbour pair -> a_o = 0),
                    // K 0..31 is a_in = 0 (right neighbour pair -> a_o = 1)
#pragma unroll
                    for (int ks = 2; ks < 4; ++ks)
                      if (leader) umma_bf16(tmem_d, desc_from(a_row + 2 * ks, hi_a), desc_from(b_s + 2 * ks, hi_b), idesc32, 1u);
#pragma unroll
                    for (int ks = 0; ks < 2; ++ks)
                      if (leader) umma_bf16(tmem_d + 32u, desc_from(a_row + 16u + 2 * ks, hi_a), desc_from(b_s + 2 * ks, hi_b), idesc32, 1u);
                  } else if (jj == 0) {   // chunk 0 = pixel a_in = 0: right neighbour pair -> a_o = 1
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks)
                      if (leader) umma_bf16(tmem_d + 32u, desc_from(a_row + 16u + 2 * ks, hi_a), desc_from(b_s + 2 * ks, hi_b), idesc32, 1u);
                  } else {                // chunk 1 = pixel a_in = 1: left neighbour pair -> a_o = 0
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks)
                      if (leader) umma_bf16(tmem_d, desc_from(a_row + 2 * ks, hi_a), desc_from(b_s + 2 * ks, hi_b), idesc32, 1u);
                  }
                }
              }
              return;
              }

#if RCU_PAIR_CENTRE_FIRST
              // all centre MMAs of the chunk first, then all side MMAs: one shape switch per chunk instead of six (32->32: -5 %)
#pragma unroll
              for (int dyi = 0; dyi < 3; ++dyi) {
                const uint32_t a_row = lo_a + (uint32_t)(dyi * kHaloPitch * 8);
                const uint32_t b_t0 = lo_b0 + (uint32_t)((jj * 3 + dyi) * 768);
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                  if (leader)
                    umma_bf16(tmem_d, desc_from(a_row + 8u + 2 * ks, hi_a), desc_from(b_t0 + 2 * ks, hi_b), idesc,
                              (jj > 0 || dyi > 0 || ks > 0) ? 1u : 0u);
              }
#endif
#pragma unroll
              for (int dyi = 0; dyi < 3; ++dyi) {
                const uint32_t a_row = lo_a + (uint32_t)(dyi * kHaloPitch * 8);
                const uint32_t b_t0 = lo_b0 + (uint32_t)((jj * 3 + dyi) * 768);
                const uint32_t b_s = b_t0 + 512u;
#if !RCU_PAIR_CENTRE_FIRST
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                  if (leader)
                    umma_bf16(tmem_d, desc_from(a_row + 8u + 2 * ks, hi_a), desc_from(b_t0 + 2 * ks, hi_b), idesc,
                              (jj > 0 || dyi > 0 || ks > 0) ? 1u : 0u);
#endif
                if (halo_is_pair32(MODE)) {
                  // one chunk holds both pixels of the pair: K 32..63 is a_in = 1 (left neighbour pair -> a_o = 0),
                  // K 0..31 is a_in = 0 (right neighbour pair -> a_o = 1)
#pragma unroll
                  for (int ks = 2; ks < 4; ++ks)
                    if (leader) umma_bf16(tmem_d, desc_from(a_row + 2 * ks, hi_a), desc_from(b_s + 2 * ks, hi_b), idesc32, 1u);
#pragma unroll
                  for (int ks = 0; ks < 2; ++ks)
                    if (leader) umma_bf16(tmem_d + 32u, desc_from(a_row + 16u + 2 * ks, hi_a), desc_from(b_s + 2 * ks, hi_b), idesc32, 1u);
                } else if (jj == 0) {   // chunk 0 = pixel a_in = 0: right neighbour pair -> a_o = 1
#pragma unroll
                  for (int ks = 0; ks < 4; ++ks)
                    if (leader) umma_bf16(tmem_d + 32u, desc_from(a_row + 16u + 2 * ks, hi_a), desc_from(b_s + 2 * ks, hi_b), idesc32, 1u);
                } else {                // chunk 1 = pixel a_in = 1: left neighbour pair -> a_o = 0
#pragma unroll
                  for (int ks = 0; ks < 4; ++ks)
                    if (leader) umma_bf16(tmem_d, desc_from(a_row + 2 * ks, hi_a), desc_from(b_s + 2 * ks, hi_b), idesc32, 1u);
                }
              }
            };
            if (j == 0) pair_taps(0); else pair_taps(1);
          } else if constexpr (MODE == HALO_UP64 && PH >= 2 && RCU_UP_PAIRED) {
            // Up-path phases in x-PAIRS: the phases (a, 0) and (a, 1) read window columns {0, 1} and {1, 2} of the same
            // rows, and their accumulators are adjacent TMEM columns.  Column 1 therefore feeds ONE N' = 2N MMA with both
            // phases' weights stacked, columns 0 and 2 one N MMA each into their phase's accumulator: 6 MMAs per K step and
            // row pair instead of 8, 20 % (N = 32) / 17 % (N = 64) fewer operand wavefronts.  Weight image per
            // (phase pair, chunk, window row i2): [2N][64] both | [N][64] column 0 -> b = 0 | [N][64] column 2 -> b = 1.
            constexpr uint32_t idesc2 = make_idesc<2 * N>();
#pragma unroll
            for (int pp = 0; pp < PH / 2; ++pp) {
              const uint32_t base_a = lo_a + prm.up_base16[2 * pp];                      // phase (a, 0): window row a, column 0
              const uint32_t wpp = lo_b0 + (uint32_t)((pp * prm.n_chunks + j) * 2) * (4u * kTile16);
#if RCU_HALO_ROLLED
#pragma unroll 1
#else
#pragma unroll
#endif
              for (int i2 = 0; i2 < 2; ++i2) {
                const uint32_t a_row = base_a + (uint32_t)(i2 * kHaloPitch * 8);
                const uint32_t wt = wpp + (uint32_t)i2 * (4u * kTile16);
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                  if (leader)
                    umma_bf16(tmem_d + (uint32_t)(2 * pp * N), desc_from(a_row + 8u + 2 * ks, hi_a), desc_from(wt + 2 * ks, hi_b), idesc2,
                              (j > 0 || i2 > 0 || ks > 0) ? 1u : 0u);
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                  if (leader) umma_bf16(tmem_d + (uint32_t)(2 * pp * N), desc_from(a_row + 2 * ks, hi_a), desc_from(wt + 2u * kTile16 + 2 * ks, hi_b), idesc, 1u);
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                  if (leader)
                    umma_bf16(tmem_d + (uint32_t)((2 * pp + 1) * N), desc_from(a_row + 16u + 2 * ks, hi_a), desc_from(wt + 3u * kTile16 + 2 * ks, hi_b), idesc, 1u);
              }
            }
          } else {
#pragma unroll
          for (int p = 0; p < PH; ++p) {
            const uint32_t lo_a0 = lo_a + (MODE == HALO_UP64 ? prm.up_base16[p] : 0u);
            // weight tiles are stored [phase][chunk][tap]
            const uint32_t lo_bj = lo_b0 + (uint32_t)(p * prm.n_chunks + j) * kChunkW16;
            if constexpr (RCU_HALO_ROLLED && MODE != HALO_UP64) {
              // rolled over the three window rows (see the pixel-pair path): three taps x kK16 MMAs per iteration
#pragma unroll 1
              for (int ty3 = 0; ty3 < 3; ++ty3) {
#pragma unroll
                for (int tx3 = 0; tx3 < 3; ++tx3) {
                  const int tap = ty3 * 3 + tx3;
                  const uint32_t a_off = (uint32_t)((ty3 * kHaloPitch + tx3) * 8);
                  const uint32_t b_off = MODE == HALO_CONV32 ? (uint32_t)(tap >> 1) * kTile16 + (uint32_t)(tap & 1) * 4u : (uint32_t)tap * kTile16;
#pragma unroll
                  for (int ks = 0; ks < kK16; ++ks) {
                    const uint32_t accumulate = (tap > 0 || ks > 0) ? 1u : (j > 0 ? 1u : 0u);
                    if (leader)
                      umma_bf16(tmem_d + (uint32_t)(p * N), desc_from(lo_a0 + a_off + 2 * ks, hi_a), desc_from(lo_bj + b_off + 2 * ks, hi_b), idesc, accumulate);
                  }
                }
              }
            } else {
#pragma unroll
            for (int tap = 0; tap < kTaps; ++tap) {
              const uint32_t a_off = MODE == HALO_UP64 ? (uint32_t)(((tap >> 1) * kHaloPitch + (tap & 1)) * 8)
                                                       : (uint32_t)(((tap / 3) * kHaloPitch + tap % 3) * 8);
              const uint32_t b_off = MODE == HALO_CONV32 ? (uint32_t)(tap >> 1) * kTile16 + (uint32_t)(tap & 1) * 4u : (uint32_t)tap * kTile16;
#pragma unroll
              for (int ks = 0; ks < kK16; ++ks) {
                const uint32_t accumulate = (tap > 0 || ks > 0) ? 1u : (j > 0 ? 1u : 0u);
                if (leader)
                  umma_bf16(tmem_d + (uint32_t)(p * N), desc_from(lo_a0 + a_off + 2 * ks, hi_a), desc_from(lo_bj + b_off + 2 * ks, hi_b), idesc, accumulate);
              }
            }
            }
          }
          }
          if (i + 1 == nt && j + 1 == prm.n_chunks && n_issuers == 2 && leader) mbar_arrive(bar_turn + 8 * (mw ^ 1));   // hand the turn over
          if (leader) umma_commit(bar_empty + 8 * stage);
          if (++stage == prm.n_stages) { stage = 0; phase ^= 1u; }
        }
        if (leader) umma_commit(bar_tfull + 8 * acc);
        if (++acc == G) { acc = 0; acc_phase ^= 1u; }
      }
      // skip the other issuer's turn
      if (n_issuers == 2) {
        int skip = TT * prm.n_chunks;
        while (skip > 0) {
          const int step = skip < prm.n_stages - stage ? skip : prm.n_stages - stage;
          stage += step; skip -= step;
          if (stage == prm.n_stages) { stage = 0; phase ^= 1u; }
        }
        for (int i = 0; i < TT; ++i)
          if (++acc == G) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else if (PATCH && warp < kEpiWarp0) {
    // ===================== patch warps (first-layer dedup): dropped channels of the landed tiles =====================
    // A tile holds the "every channel kept" variant of the first convolution's output for this slice.  Dropout2d sits
    // between conv + bias and BatchNorm (unet.py:13-19), so a dropped channel of a (sample, slice) is relu(d_c) at every
    // pixel: its two 2-byte elements per window row (pixel 0 / 1 of the pair) are overwritten — inside the image only,
    // the zero padding stays — then the tile is handed to the MMA warps.  One box per tile (32-channel source); tile i of
    // the CTA's range lands in stage i mod n_stages.  One warp alone cannot keep up with the tensor pipe: either every patch warp
    // takes a quarter of the rows of every tile (RCU_HALO_PATCH_SPLIT, bar_full then counts one arrival per warp), or a stage
    // belongs to patch warp stage mod kHaloPatchWarps (per-stage ownership keeps the parity waits unambiguous for any n_stages).
    if (patching) {
      const int pw = warp - 3;
      unsigned short* my_patch = s_patch + 32 * pw;
      // this lane's window rows: position, byte offset and swizzle term never change.  Split mode: the rows of a tile are dealt
      // out over all patch warps (row = 32 * (k * warps + pw) + lane), whole-tile mode: lane, lane + 32, ... of the warp's own tiles
      constexpr int kRowStep = RCU_HALO_PATCH_SPLIT ? 32 * kHaloPatchWarps : 32;
      constexpr int kRowsPerLane = (kHaloRows * kHaloPitch + kRowStep - 1) / kRowStep;
      const int row0 = RCU_HALO_PATCH_SPLIT ? 32 * pw + lane : lane;
      int wy[kRowsPerLane], wx[kRowsPerLane];
      uint32_t roff[kRowsPerLane], rsw[kRowsPerLane];
#pragma unroll
      for (int k = 0; k < kRowsPerLane; ++k) {
        const int r = row0 + kRowStep * k;
        wy[k] = r / kHaloPitch; wx[k] = r - wy[k] * kHaloPitch;
        roff[k] = (uint32_t)r * 128u; rsw[k] = (uint32_t)(r & 7);
      }
      int cur_img = -1;
      uint32_t mask = 0;
      const int n_tiles = t_end - t_begin;
      int stage = 0;
      uint32_t phase = 0;
      for (int i = 0; i < n_tiles; ++i, stage = (stage + 1 == prm.n_stages ? 0 : stage + 1), phase ^= (stage == 0 ? 1u : 0u)) {
        if (!RCU_HALO_PATCH_SPLIT && stage % kHaloPatchWarps != pw) continue;   // a stage always belongs to the same warp: it sees every phase of its barrier in order
        const int tile = t_begin + i;
        const int img = tile / tiles_per_img;
        const int rem = tile - img * tiles_per_img;
        const int ty = rem / prm.tiles_x, tx = rem - ty * prm.tiles_x;
        if (img != cur_img) {
          const float2 c = __ldg(prm.patch_coef + (long long)img * prm.coef_stride + prm.patch_off + lane);
          const bool stochastic = !(prm.dedup_mode == 2 && img / prm.chunk_slices == 0);
          mask = __ballot_sync(0xffffffffu, stochastic && c.x == 0.0f);
          __syncwarp();
          my_patch[lane] = __bfloat16_as_ushort(__float2bfloat16_rn(fmaxf(c.y, 0.0f)));
          __syncwarp();
          cur_img = img;
        }
        mbar_wait_slack(bar_land + 8 * stage, phase);
        if (mask != 0u) {
          uint8_t* tile_ptr = base_ptr + (w_span + (uint32_t)stage * prm.chunk_stride);
          const int yb = ty * kHaloTileH - 1, xb = tx * kHaloTileW - 1;
          for (uint32_t m = mask; m != 0u; m &= m - 1u) {
            const uint32_t c = (uint32_t)__ffs(m) - 1u;
            const unsigned short v = my_patch[c];
            const uint32_t ch = c >> 3, cl = (c & 7u) << 1;
#pragma unroll
            for (int k = 0; k < kRowsPerLane; ++k) {
              const int y = yb + wy[k], xp = xb + wx[k];
              if (row0 + kRowStep * k < kHaloRows * kHaloPitch && y >= 0 && y < prm.in_h && xp >= 0 && xp < prm.in_w) {
                // SWIZZLE_128B: the 16-byte piece index of an element is XORed with the row index modulo 8
                const uint32_t o = roff[k] + (((ch ^ rsw[k]) << 4) | cl);
#ifndef RCU_PATCH_EXP_NOSTORE
                *reinterpret_cast<unsigned short*>(tile_ptr + o) = v;            // pixel 0 of the pair: K = c
                *reinterpret_cast<unsigned short*>(tile_ptr + (o ^ 64u)) = v;     // pixel 1: K = 32 + c, four pieces further
#endif
              }
            }
          }
#ifndef RCU_PATCH_EXP_NOFENCE
          fence_proxy_async();   // generic-proxy stores -> visible to the tensor cores' async-proxy reads
#endif
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_full + 8 * stage);
      }
    }
  } else {
    // ===================== epilogue: G groups of 4 warps; group g owns accumulator stage g =====================
    const int group = (warp - kEpiWarp0) >> 2;
    const int q = warp & 3;               // TMEM lane quarter this warp may access (each group holds all four)
    const int row = q * 32 + lane;
    const int gt = threadIdx.x - 32 * kEpiWarp0 - group * 128;   // 0..127 inside the group
    float2* coef = s_coef + group * N;
    const uint32_t taddr0 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(group * S::kAccCols);
    uint32_t acc_phase = 0;
    int cur_img = -1;
    for (int tile = t_begin + group; tile < t_end; tile += G) {
      const int img = tile / tiles_per_img;
      const int rem = tile - img * tiles_per_img;
      const int ty = rem / prm.tiles_x, tx = rem - ty * prm.tiles_x;

      if (img != cur_img) {   // group-uniform
        asm volatile("bar.sync %0, 128;" ::"r"(group + 1) : "memory");   // everyone is done with the old coefficients
        if (gt < N) coef[gt] = __ldg(prm.coef + (long long)img * prm.coef_stride + prm.coef_off + (PAIR ? (gt & 31) : gt));   // pair rows: both pixels share the 32 channel coefficients
        asm volatile("bar.sync %0, 128;" ::"r"(group + 1) : "memory");
        cur_img = img;
      }

      mbar_wait_slack(bar_tfull + 8 * group, acc_phase);
      tc_fence_after();

      const int y = ty * kHaloTileH + (row >> 3), x = tx * kHaloTileW + (row & 7);
      const bool valid = (y < prm.in_h) && (x < prm.in_w);
      if constexpr (PAIR) {
        // row = pixel pair (y, x): accumulator columns [0, 32) are output pixel 2x, [32, 64) pixel 2x + 1 (prm.in_w counts pairs)
        if constexpr (MODE == HALO_PAIR32_HEAD) {   // the head layer has its own instantiation (own register allocation, own MMA sequence)
          // 16 channels of BOTH pixels per step: a channel's coefficients are read once for the pair (the broadcast reads of the
          // coefficients were a quarter of the kernel's shared-memory wavefronts when every pixel read them again)
          float lg[4] = {prm.head_w[64], prm.head_w[65], prm.head_w[64], prm.head_w[65]};
#pragma unroll
          for (int cc = 0; cc < 2; ++cc) {
            uint32_t v0[16], v1[16];
            tmem_ld_32x32b_x16(taddr0 + (uint32_t)(cc * 16), v0);
            tmem_ld_32x32b_x16(taddr0 + (uint32_t)(32 + cc * 16), v1);
            tmem_ld_wait();
            if (cc == 1) {
              tc_fence_before();
              mbar_arrive_warp(bar_tempty + 8 * group);
            }
#pragma unroll
            for (int c = 0; c < 16; c += 2) {
              const int ch = cc * 16 + c;
              const float4 k = *reinterpret_cast<const float4*>(&coef[ch]);   // two channels per 16-byte broadcast read
              float a0 = fmaf(__uint_as_float(v0[c]), k.x, k.y), a1 = fmaf(__uint_as_float(v0[c + 1]), k.z, k.w);
              float b0 = fmaf(__uint_as_float(v1[c]), k.x, k.y), b1 = fmaf(__uint_as_float(v1[c + 1]), k.z, k.w);
              if (prm.relu) { a0 = fmaxf(a0, 0.0f); a1 = fmaxf(a1, 0.0f); b0 = fmaxf(b0, 0.0f); b1 = fmaxf(b1, 0.0f); }
              lg[0] = fmaf(a0, prm.head_w[ch], lg[0]);
              lg[1] = fmaf(a0, prm.head_w[32 + ch], lg[1]);
              lg[0] = fmaf(a1, prm.head_w[ch + 1], lg[0]);
              lg[1] = fmaf(a1, prm.head_w[33 + ch], lg[1]);
              lg[2] = fmaf(b0, prm.head_w[ch], lg[2]);
              lg[3] = fmaf(b0, prm.head_w[32 + ch], lg[3]);
              lg[2] = fmaf(b1, prm.head_w[ch + 1], lg[2]);
              lg[3] = fmaf(b1, prm.head_w[33 + ch], lg[3]);
            }
          }
          if (valid) {
            const int t = img / prm.chunk_slices, sl = img - t * prm.chunk_slices;
            const long long gimg = (long long)t * prm.n_slices_total + prm.slice0 + sl;
            if (prm.head_diff) {
              float* dst = prm.logits + (gimg * prm.out_h + y) * prm.out_w + 2 * x;
              *reinterpret_cast<float2*>(dst) = make_float2(lg[0] - lg[1], lg[2] - lg[3]);
            } else {
              float2* dst = reinterpret_cast<float2*>(prm.logits) + (gimg * prm.out_h + y) * prm.out_w + 2 * x;   // out_w is even: 16-byte aligned
              *reinterpret_cast<float4*>(dst) = make_float4(lg[0], lg[1], lg[2], lg[3]);
            }
          }
        } else {
          uint32_t packed[2][16];
          // 16 channels of BOTH pixels per step: a channel's coefficients are read once for the pair
#pragma unroll
          for (int cc = 0; cc < 2; ++cc) {
            uint32_t v0[16], v1[16];
            tmem_ld_32x32b_x16(taddr0 + (uint32_t)(cc * 16), v0);
            tmem_ld_32x32b_x16(taddr0 + (uint32_t)(32 + cc * 16), v1);
            tmem_ld_wait();
            if (cc == 1) {
              tc_fence_before();
              mbar_arrive_warp(bar_tempty + 8 * group);
            }
#pragma unroll
            for (int c = 0; c < 16; c += 2) {
              RCU_COEF2(cc * 16 + c);
              float a0 = fmaf(__uint_as_float(v0[c]), c0.x, c0.y), a1 = fmaf(__uint_as_float(v0[c + 1]), c1.x, c1.y);
              float b0 = fmaf(__uint_as_float(v1[c]), c0.x, c0.y), b1 = fmaf(__uint_as_float(v1[c + 1]), c1.x, c1.y);
              if (prm.relu) { a0 = fmaxf(a0, 0.0f); a1 = fmaxf(a1, 0.0f); b0 = fmaxf(b0, 0.0f); b1 = fmaxf(b1, 0.0f); }
              __nv_bfloat162 pa = __floats2bfloat162_rn(a0, a1), pb = __floats2bfloat162_rn(b0, b1);
              packed[0][cc * 8 + (c >> 1)] = *reinterpret_cast<uint32_t*>(&pa);
              packed[1][cc * 8 + (c >> 1)] = *reinterpret_cast<uint32_t*>(&pb);
            }
          }
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            if (prm.tma_store) {
              // the even / odd output pixels of the tile are two strided views of the destination (out_maps.m[half])
              const uint32_t so = smem_out + (uint32_t)group * S::kOutSlot;
              if (gt == 0) bulk_wait_group_read0();
              asm volatile("bar.sync %0, 128;" ::"r"(group + 1) : "memory");
              const uint32_t rowa = so + (uint32_t)row * 64u;
              const uint32_t sw = (uint32_t)(row >> 1) & 3u;
#pragma unroll
              for (int i = 0; i < 4; ++i)
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rowa + (((uint32_t)i ^ sw) << 4)), "r"(packed[half][4 * i]),
                             "r"(packed[half][4 * i + 1]), "r"(packed[half][4 * i + 2]), "r"(packed[half][4 * i + 3])
                             : "memory");
              fence_proxy_async();
              asm volatile("bar.sync %0, 128;" ::"r"(group + 1) : "memory");
              if (gt == 0) {
                tma_store_4d(&out_maps.m[half], so, 0, tx * kHaloTileW, ty * kHaloTileH, img);
                bulk_commit_group();
              }
            } else if (valid) {
              uint4* d4 = reinterpret_cast<uint4*>(prm.out + (long long)img * prm.out_img_stride + ((long long)y * prm.out_w + 2 * x + half) * prm.out_c);
#pragma unroll
              for (int i = 0; i < 4; ++i) d4[i] = make_uint4(packed[half][4 * i], packed[half][4 * i + 1], packed[half][4 * i + 2], packed[half][4 * i + 3]);
            }
          }
          if (prm.pool_out != nullptr) {
            // 2x2 max: the x neighbour is the other half of this row, the y neighbour is lane ^ 8 (lane = (y & 3) * 8 + x)
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              __nv_bfloat162 m = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&packed[0][i]), *reinterpret_cast<__nv_bfloat162*>(&packed[1][i]));
              uint32_t mm = *reinterpret_cast<uint32_t*>(&m);
              uint32_t o = __shfl_xor_sync(0xffffffffu, mm, 8);
              m = __hmax2(m, *reinterpret_cast<__nv_bfloat162*>(&o));
              packed[0][i] = *reinterpret_cast<uint32_t*>(&m);
            }
            if (valid) {
              // both lanes of a quad hold the 32 maxima: the even row stores channels 0..15, the odd row 16..31
              const int part = y & 1;
              uint4* pd = reinterpret_cast<uint4*>(prm.pool_out + (long long)img * prm.pool_img_stride +
                                                   ((long long)(y >> 1) * prm.in_w + x) * 32 + part * 16);
              pd[0] = part == 0 ? make_uint4(packed[0][0], packed[0][1], packed[0][2], packed[0][3])
                                : make_uint4(packed[0][8], packed[0][9], packed[0][10], packed[0][11]);
              pd[1] = part == 0 ? make_uint4(packed[0][4], packed[0][5], packed[0][6], packed[0][7])
                                : make_uint4(packed[0][12], packed[0][13], packed[0][14], packed[0][15]);
            }
          }
        }
      } else {
#pragma unroll 1
      for (int p = 0; p < PH; ++p) {
      const uint32_t taddr = taddr0 + (uint32_t)(p * N);
      const int oy = prm.out_mul * y + (MODE == HALO_UP64 ? prm.up_dy[p] : 0), ox = prm.out_mul * x + (MODE == HALO_UP64 ? prm.up_dx[p] : 0);

      if (prm.head != nullptr) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(taddr, v);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive_warp(bar_tempty + 8 * group);   // accumulator is in registers: the MMA warp may reuse the stage
        float l0 = prm.head_w[64], l1 = prm.head_w[65];
#pragma unroll
        for (int c = 0; c < 32; c += 2) {
          const float4 cc = *reinterpret_cast<const float4*>(&coef[c]);
          float a0 = fmaf(__uint_as_float(v[c]), cc.x, cc.y), a1 = fmaf(__uint_as_float(v[c + 1]), cc.z, cc.w);
          if (prm.relu) { a0 = fmaxf(a0, 0.0f); a1 = fmaxf(a1, 0.0f); }
          l0 = fmaf(a0, prm.head_w[c], l0);
          l1 = fmaf(a0, prm.head_w[32 + c], l1);
          l0 = fmaf(a1, prm.head_w[c + 1], l0);
          l1 = fmaf(a1, prm.head_w[33 + c], l1);
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
        __nv_bfloat16* dst = prm.out + (long long)img * prm.out_img_stride + ((long long)oy * prm.out_w + ox) * prm.out_c;
#pragma unroll 1
        for (int cb = 0; cb < N; cb += 32) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(taddr + (uint32_t)cb, v);
          tmem_ld_wait();
          if (cb + 32 == N && p + 1 == PH) {
            tc_fence_before();
            mbar_arrive_warp(bar_tempty + 8 * group);
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
            // stage the 128 x 32 tile in shared memory (64-byte rows, SWIZZLE_64B: 16-byte chunk ^= (row >> 1) & 3, which
            // also makes the 4 x STS.128 of a warp conflict-free) and let ONE TMA tensor store write it: the direct path
            // costs 32 LSU wavefronts per STG.128 (every lane its own line), and TMA clips tiles that overhang the image
            const uint32_t so = smem_out + (uint32_t)group * S::kOutSlot;
            if (gt == 0) bulk_wait_group_read0();                         // the previous store has finished reading the slot
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
              tma_store_4d(&out_maps.m[p], so, cb, tx * kHaloTileW, ty * kHaloTileH, img);
              bulk_commit_group();
            }
          } else if (valid) {
            uint4* d4 = reinterpret_cast<uint4*>(dst + cb);
#pragma unroll
            for (int i = 0; i < 4; ++i) d4[i] = make_uint4(packed[4 * i], packed[4 * i + 1], packed[4 * i + 2], packed[4 * i + 3]);
          }
          if (prm.pool_out != nullptr) {
            // 2x2 max over the quad (y^1, x^1): lane = (y & 3) * 8 + x, so the partners are lanes ^1 and ^8 of this warp.
            // Out-of-image pixels only ever pair with out-of-image pixels (h, w are even), their quads are not stored.
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
                                  ((long long)(y >> 1) * (prm.in_w >> 1) + (x >> 1)) * N + cb + part * 8;
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
      }  // phases
      }
      acc_phase ^= 1u;
    }
    if (prm.tma_store && gt == 0) bulk_wait_group_read0();   // shared memory must outlive the last store's read
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, S::kTmemCols);
}

}  // namespace rcu
