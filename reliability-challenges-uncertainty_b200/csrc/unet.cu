// U-Net forward engine behind rcu_unet_* (include/rcu_b200.h).
//
// Restates UNet.forward (common/model/unet.py:166-186) for residual=False, sigma_out=False, bn=True,
// transpose=False (the only configuration the reference ships, unet.py:157) as a fixed schedule of kernels over
// bf16 NHWC activations with the MC samples folded into the batch ([sample][slice] image order inside a chunk).
#include <cuda.h>
#include <cuda_bf16.h>

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "common.cuh"
#include "conv_tc.cuh"
#include "conv_halo.cuh"
#include "conv_wide.cuh"
#include "philox.cuh"
#include "unet_kernels.cuh"
#include "postnet.cuh"

namespace rcu {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) {
    cudaGetLastError();
    return nullptr;
  }
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

enum OpKind { OP_FIRST = 0, OP_CONV = 1, OP_POOL = 2, OP_RES_FIRST = 3 };

// A channel slice of a bf16 NHWC tensor inside the arena.  Producers write slices of the concat buffers directly
// (torch.cat((up, skip), 1), common/model/unet.py:118, is never materialised by a copy).  `padded` tensors carry one
// zero row above every image (unused at present: every Act is dense).
struct Act {
  __nv_bfloat16* base = nullptr;   // element (image 0, y 0, x 0, first channel of the slice)
  int c = 0;                       // channels of the slice
  int c_total = 0;                 // pixel stride in elements
  int h = 0, w = 0;
  long long img_stride = 0;        // elements between images
  bool padded = false;
};

struct HaloPack {                  // create-time description of a layer the halo kernel can run
  bool ok = false;
  bool pair = false;               // 32-channel source: 64-byte pixel rows, two K = 16 steps per tap
  int n_chunks = 0;
  int n_phases = 1;
  uint8_t* d_wimg[kMaxPhases] = {nullptr, nullptr, nullptr, nullptr};   // per-phase views into one contiguous image
  uint32_t w_bytes = 0;            // bytes of ONE phase
  int phases_per_launch = 1;       // up path: 4, 2 or 1 phases share a launch (weights of all of them resident)
  int pair_mode = 0;               // 1 / 2: pixel-pair rows over a 32- / 64-channel source (HALO_PAIR32 / HALO_PAIR64)
};

struct WidePack {                  // create-time description of a layer the wide kernel can run (c_out % 128 == 0)
  bool ok = false;
  int n_chunks = 0, n_ntiles = 0, n_phases = 1;
  uint8_t* d_wimg = nullptr;       // [phase][ntile][chunk][tap][128][64] pre-swizzled tiles
};

struct ConvLayer {
  int c0 = 0, c1 = 0, c_out = 0;        // c1 is always 0 now (concat buffers), kept for the kernel's two-map interface
  int n_taps = 9, n_phases = 1, out_mul = 1, relu = 1;
  signed char dy[kMaxPhases][kMaxTaps];
  signed char dx[kMaxPhases][kMaxTaps];
  __nv_bfloat16* d_weights = nullptr;  // [n_phases * n_taps][c_out][c0 + c1]
  int coef_off = 0;
  int block_n = 0, kc = 0;
  bool head = false;
  bool accumulate = false;               // residual branch: a 1x1 conv of the block input added onto the block output (per-tap kernel)
  const float* d_head = nullptr;         // fused 1x1 head of this layer: [2][c_out] + [2]
  float h_head[66] = {0};                // host copy (kernel parameter of the halo kernels)
  int out_slot = 0;                      // 0: class logits (conv_cls), 1: sigma logits (conv_sigma, unet.py:162-164)
  HaloPack halo;
  HaloPack pairp;                        // pixel-pair variant of the same layer (c_out = 32, dense source), preferred when planned
  WidePack wide;
  // plan-time
  int in_h = 0, in_w = 0;
  Act src, dst;
  Act pool_dst;                          // when set: the halo kernel also writes MaxPool2d(2) of the output here
  CUtensorMap map_a0, map_a1, map_w, map_halo, map_wide;
  CUtensorMap map_out[kMaxPhases];       // halo kernel: TMA-store views of the destination (one per up-path phase)
  bool tma_store = false;
  int halo_stages = 0;
  bool use_pair = false;                 // plan-time: run the pixel-pair kernel
  bool pair_tma_store = false;
  int pair_stages = 0;
  CUtensorMap map_halo_pair;             // source viewed as [h][w/2][2c]
  CUtensorMap map_out_pair[2];           // even / odd output pixels of the destination
};

struct Op {
  OpKind kind;
  int conv = -1;                         // index into convs for OP_CONV
  Act in;                                // OP_POOL input
  Act out;                               // output (debug view)
  int h = 0, w = 0, c = 0;               // output dims (for POOL: input dims in in_h/in_w)
  int in_h = 0, in_w = 0;
};

}  // namespace rcu

using namespace rcu;

struct rcu_unet {
  int device = 0;
  int in_channels = 0, depth = 0, start_filters = 0;
  float p_drop = 0.f;
  int n_sites = 0, total_dropout_channels = 0;
  std::vector<int> site_channels;
  // device constants
  float* d_first_w = nullptr;            // [c_in*9][sf]
  int first_coef_off = 0;
  float* d_head = nullptr;               // conv_cls.1:   [2][sf] + [2]
  float* d_sigma_head = nullptr;         // conv_sigma.1: [2][sf] + [2] (sigma_out nets only)
  bool has_sigma = false;
  bool residual = false;                 // residual=True nets (ConvResidualBlock, common/model/unet.py:42-60)
  float* d_res0_w = nullptr;             // first block's residual 1x1 conv on the float32 input: [sf][c_in], and its bias [sf]
  float* d_res0_b = nullptr;
  Act res0_dst;
  std::vector<ConvLayer> convs;          // execution order
  CoefColumns cols{};
  std::vector<void*> owned;              // device allocations freed in destroy
  int n_cols = 0;
  // plan
  int H = 0, W = 0, max_images = 0;
  void* arena = nullptr;                 // caller-owned activation workspace (rcu_unet_bind_workspace)
  size_t arena_bytes = 0;
  size_t planned_bytes = 0;
  float2* d_coef = nullptr;
  Act first_out;
  Act head_feat;                         // features of conv_cls.0 for the cross-check path
  Act sigma_feat;                        // same for conv_sigma.0
  Act features;                          // input of conv_cls = `UNet.features` (unet.py:178-179)
  std::vector<Op> ops;
  int conv_impl = 0;
  unsigned long long halo_mask = ~0ull;  // debug: bit i enables the halo kernel for conv i (execution order)
  bool pair_enabled = true;              // debug: rcu_unet_set_conv_impl(net, 3) runs the pixel-row halo kernel where the pair kernel would
  bool first_dedup = true;               // store the first convolution's output once per slice (rcu_unet_set_first_layer_dedup)
  bool unit0_dropout = false;            // the first unit has a Dropout2d site
  float2* d_e0coef = nullptr;            // [2][start_filters] epilogue coefficients of the first conv: deterministic / every channel kept
  int last_dedup_mode = 0;
  long long last_launches = 0;
  int last_n_img = 0;
  // optional per-op timing
  bool timing = false;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev_pool;
  std::vector<int> ev_op;      // op index of every recorded pair since the last read
  size_t ev_used = 0;
};

struct rcu_postnet {
  int device = 0;
  int n_convs = 0;
  rcu::PostNetWeights w;   // host copy; passed by value as a __grid_constant__ kernel parameter
};

namespace rcu {
static int launch_postnet(const rcu_postnet* post, bool src_bf16, const void* src, int px_stride, long long img_stride, int hw, int n_img,
                          int chunk_slices, long long slice0, long long n_slices_total, float* logits, cudaStream_t st) {
  const long long total = (long long)n_img * hw;
  if (total == 0) return RCU_OK;
  long long blocks = (total + 127) / 128;
  if (blocks > (long long)sm_count() * 64) blocks = (long long)sm_count() * 64;
#define RCU_POSTNET_CASE(NC)                                                                                                      \
  case NC:                                                                                                                        \
    if (src_bf16) postnet_kernel<NC, true><<<(unsigned)blocks, 128, 0, st>>>(post->w, src, px_stride, img_stride, hw, n_img,      \
                                                                              chunk_slices, slice0, n_slices_total, logits);      \
    else postnet_kernel<NC, false><<<(unsigned)blocks, 128, 0, st>>>(post->w, src, px_stride, img_stride, hw, n_img, chunk_slices, \
                                                                     slice0, n_slices_total, logits);                              \
    break;
  switch (post->n_convs) {
    RCU_POSTNET_CASE(1) RCU_POSTNET_CASE(2) RCU_POSTNET_CASE(3) RCU_POSTNET_CASE(4)
    default: set_error("PostNet with %d convs", post->n_convs); return RCU_ENOTSUP;
  }
#undef RCU_POSTNET_CASE
  RCU_LAUNCH_CHECK();
  return RCU_OK;
}
}  // namespace rcu

namespace rcu {
struct OpTimer {  // brackets one launch with events when timing is on
  rcu_unet* net; cudaStream_t st; int slot = -1;
  OpTimer(rcu_unet* n, int op, cudaStream_t s) : net(n), st(s) {
    if (!net->timing) return;
    if (net->ev_used == net->ev_pool.size()) {
      cudaEvent_t a, b;
      if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) { cudaGetLastError(); return; }
      net->ev_pool.push_back({a, b});
    }
    slot = (int)net->ev_used++;
    net->ev_op.push_back(op);
    cudaEventRecord(net->ev_pool[slot].first, st);
  }
  ~OpTimer() { if (slot >= 0) cudaEventRecord(net->ev_pool[slot].second, st); }
};
}  // namespace rcu

namespace rcu {

template <typename T>
static int dev_upload(rcu_unet* net, const std::vector<T>& host, T** out) {
  void* d = nullptr;
  RCU_CUDA(cudaMalloc(&d, host.size() * sizeof(T) + 16));
  net->owned.push_back(d);
  RCU_CUDA(cudaMemcpy(d, host.data(), host.size() * sizeof(T), cudaMemcpyHostToDevice));
  *out = reinterpret_cast<T*>(d);
  return RCU_OK;
}

static inline uint16_t f32_to_bf16_rn(float f) {
  uint32_t u;
  std::memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40u);  // NaN
  const uint32_t lsb = (u >> 16) & 1u;
  u += 0x7fffu + lsb;
  return (uint16_t)(u >> 16);
}

static void fold_unit(const rcu_conv_unit& u, float eps, std::vector<float>& fa, std::vector<float>& fba, std::vector<float>& fd) {
  for (int c = 0; c < u.c_out; ++c) {
    if (u.bn_weight) {
      const float a = u.bn_weight[c] / std::sqrt(u.bn_var[c] + eps);
      fa.push_back(a);
      fba.push_back(u.bias[c] * a);
      fd.push_back(u.bn_bias[c] - a * u.bn_mean[c]);
    } else {
      fa.push_back(1.0f);
      fba.push_back(u.bias[c]);
      fd.push_back(0.0f);
    }
  }
}

static int pick_tiles(int c_src_min, int c_out, int* block_n, int* kc) {
  *kc = (c_src_min % 64 == 0) ? 64 : 32;
  *block_n = c_out >= 256 ? 256 : c_out;
  if (c_src_min % 32 != 0 || (*block_n != 32 && *block_n != 64 && *block_n != 128 && *block_n != 256) || c_out % *block_n != 0) {
    set_error("channel configuration (c_in multiple of %d, c_out=%d) is outside the tcgen05 tile set", c_src_min, c_out);
    return RCU_ENOTSUP;
  }
  return RCU_OK;
}

// weights [c_out][c_in][3][3] fp32 -> bf16 [tap][c_out][c_in]
static std::vector<uint16_t> pack_conv3x3(const rcu_conv_unit& u) {
  std::vector<uint16_t> w((size_t)9 * u.c_out * u.c_in);
  for (int tap = 0; tap < 9; ++tap)
    for (int co = 0; co < u.c_out; ++co)
      for (int ci = 0; ci < u.c_in; ++ci)
        w[((size_t)tap * u.c_out + co) * u.c_in + ci] = f32_to_bf16_rn(u.weight[((size_t)co * u.c_in + ci) * 9 + tap]);
  return w;
}

// nearest-x2 followed by conv3x3(pad 1) == four 2x2-tap convolutions on the low-res input, one per output parity
// (a, b): rows {2y+a-1, 2y+a, 2y+a+1} of the upsampled image collapse onto low-res rows
//   a = 0: {y-1 | y, y}    -> taps i=0: dy=-1, ky {0};   i=1: dy=0,  ky {1,2}
//   a = 1: {y, y | y+1}    -> taps i=0: dy=0,  ky {0,1}; i=1: dy=+1, ky {2}
// (same along x).  Zero padding of the upsampled image coincides with zero OOB fill of the low-res image.
static std::vector<uint16_t> pack_upconv_phases(const rcu_conv_unit& u, ConvLayer& L) {
  std::vector<uint16_t> w((size_t)16 * u.c_out * u.c_in);
  for (int a = 0; a < 2; ++a)
    for (int b = 0; b < 2; ++b) {
      const int ph = a * 2 + b;
      for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 2; ++j) {
          const int tap = i * 2 + j;
          L.dy[ph][tap] = (signed char)(a == 0 ? i - 1 : i);
          L.dx[ph][tap] = (signed char)(b == 0 ? j - 1 : j);
          int ky0, ky1, kx0, kx1;
          if (a == 0) { ky0 = i == 0 ? 0 : 1; ky1 = i == 0 ? 0 : 2; } else { ky0 = i == 0 ? 0 : 2; ky1 = i == 0 ? 1 : 2; }
          if (b == 0) { kx0 = j == 0 ? 0 : 1; kx1 = j == 0 ? 0 : 2; } else { kx0 = j == 0 ? 0 : 2; kx1 = j == 0 ? 1 : 2; }
          for (int co = 0; co < u.c_out; ++co)
            for (int ci = 0; ci < u.c_in; ++ci) {
              float s = 0.0f;
              for (int ky = ky0; ky <= ky1; ++ky)
                for (int kx = kx0; kx <= kx1; ++kx) s += u.weight[((size_t)co * u.c_in + ci) * 9 + ky * 3 + kx];
              w[((size_t)(ph * 4 + tap) * u.c_out + co) * u.c_in + ci] = f32_to_bf16_rn(s);
            }
        }
    }
  return w;
}

static int make_act_map(CUtensorMap* map, const Act& a, int n, int kc) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available (driver too old?)"); return RCU_ECUDA; }
  cuuint64_t dims[4] = {(cuuint64_t)a.c, (cuuint64_t)a.w, (cuuint64_t)a.h, (cuuint64_t)n};
  cuuint64_t strides[3] = {(cuuint64_t)a.c_total * 2, (cuuint64_t)a.w * a.c_total * 2, (cuuint64_t)a.img_stride * 2};
  cuuint32_t box[4] = {(cuuint32_t)kc, (cuuint32_t)kTileW, (cuuint32_t)kTileH, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUtensorMapSwizzle sw = kc == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, a.base, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(activation c=%d w=%d h=%d n=%d kc=%d) failed: %d", a.c, a.w, a.h, n, kc, (int)r); return RCU_ECUDA; }
  return RCU_OK;
}

// Halo window of the halo kernel: (16+2) x (8+2) pixels x box_c channels, SWIZZLE_128B, zero fill outside the image.
// box_c = 64 fills whole 128-byte rows.  box_c = 32 (32-channel sources): TMA still gives every pixel its own
// 128-byte swizzled row and fills the first 64 logical bytes (measured: tools/ubench/umma_probe.cu `tma5d`), so the
// same SWIZZLE_128B descriptors read it at full rate with two K = 16 steps per tap instead of four.
static int make_halo_map(CUtensorMap* map, const Act& a, int n, int box_c) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available (driver too old?)"); return RCU_ECUDA; }
  cuuint64_t dims[4] = {(cuuint64_t)a.c, (cuuint64_t)a.w, (cuuint64_t)a.h, (cuuint64_t)n};
  cuuint64_t strides[3] = {(cuuint64_t)a.c_total * 2, (cuuint64_t)a.w * a.c_total * 2, (cuuint64_t)a.img_stride * 2};
  cuuint32_t box[4] = {(cuuint32_t)box_c, (cuuint32_t)kHaloPitch, (cuuint32_t)kHaloRows, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, a.base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(halo c=%d w=%d h=%d n=%d) failed: %d", a.c, a.w, a.h, n, (int)r); return RCU_ECUDA; }
  return RCU_OK;
}

// Destination view for the halo kernel's TMA tensor stores: 128-pixel x 32-channel boxes, SWIZZLE_64B.  For an up-path
// phase (a, b) the view is the (2y + a, 2x + b) sub-lattice of the high-resolution tensor (doubled strides).
static int make_out_map(CUtensorMap* map, const Act& d, int n, int out_mul, int a, int b) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available (driver too old?)"); return RCU_ECUDA; }
  cuuint64_t dims[4] = {(cuuint64_t)d.c, (cuuint64_t)(d.w / out_mul), (cuuint64_t)(d.h / out_mul), (cuuint64_t)n};
  cuuint64_t strides[3] = {(cuuint64_t)d.c_total * 2 * out_mul, (cuuint64_t)d.w * d.c_total * 2 * out_mul, (cuuint64_t)d.img_stride * 2};
  cuuint32_t box[4] = {32, (cuuint32_t)kHaloTileW, (cuuint32_t)kHaloTileH, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  void* base = d.base + ((size_t)a * d.w + b) * d.c_total;
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(out c=%d w=%d h=%d n=%d) failed: %d", d.c, d.w, d.h, n, (int)r); return RCU_ECUDA; }
  return RCU_OK;
}

static int make_weight_map(CUtensorMap* map, const void* base, int c_in, int c_out, int taps, int kc, int block_n) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available (driver too old?)"); return RCU_ECUDA; }
  cuuint64_t dims[3] = {(cuuint64_t)c_in, (cuuint64_t)c_out, (cuuint64_t)taps};
  cuuint64_t strides[2] = {(cuuint64_t)c_in * 2, (cuuint64_t)c_out * c_in * 2};
  cuuint32_t box[3] = {(cuuint32_t)kc, (cuuint32_t)block_n, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  const CUtensorMapSwizzle sw = kc == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(weights c_in=%d c_out=%d taps=%d) failed: %d", c_in, c_out, taps, (int)r); return RCU_ECUDA; }
  return RCU_OK;
}

template <int BLOCK_N, int KC>
static int launch_conv_tc(const ConvLayer& L, const ConvParams& prm, cudaStream_t st) {
  using S = ConvSmem<BLOCK_N, KC>;
  auto kern = conv_tc_kernel<BLOCK_N, KC>;
  static int ctas_per_sm[64] = {0};
  int dev = 0;
  RCU_CUDA(cudaGetDevice(&dev));
  dev = dev < 0 || dev >= 64 ? 0 : dev;
  if (ctas_per_sm[dev] == 0) {
    RCU_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal));
    int occ = 0;
    RCU_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kConvThreads, S::kTotal));
    // TMEM: 512 columns per SM shared by the resident CTAs
    const int tmem_limit = 512 / S::kTmemCols;
    if (occ > tmem_limit) occ = tmem_limit;
    if (occ > 2) occ = 2;
    if (occ < 1) { set_error("conv_tc_kernel<%d,%d> does not fit on an SM", BLOCK_N, KC); return RCU_ECUDA; }
    ctas_per_sm[dev] = occ;
  }
  const long long total_tiles = (long long)prm.n_img * prm.n_phases * prm.tiles_y * prm.tiles_x * prm.n_tiles_n;
  long long grid = (long long)sm_count() * ctas_per_sm[dev];
  if (grid > total_tiles) grid = total_tiles;
  if (grid < 1) return RCU_OK;
  kern<<<(unsigned)grid, kConvThreads, S::kTotal, st>>>(L.map_a0, L.map_a1, L.map_w, prm);
  RCU_LAUNCH_CHECK();
  return RCU_OK;
}

static int dispatch_conv_tc(const ConvLayer& L, const ConvParams& prm, cudaStream_t st) {
  if (L.kc == 32) {
    if (L.block_n == 32) return launch_conv_tc<32, 32>(L, prm, st);
    if (L.block_n == 64) return launch_conv_tc<64, 32>(L, prm, st);
  } else {
    if (L.block_n == 32) return launch_conv_tc<32, 64>(L, prm, st);
    if (L.block_n == 64) return launch_conv_tc<64, 64>(L, prm, st);
    if (L.block_n == 128) return launch_conv_tc<128, 64>(L, prm, st);
    if (L.block_n == 256) return launch_conv_tc<256, 64>(L, prm, st);
  }
  set_error("no tcgen05 conv instantiation for BLOCK_N=%d KC=%d", L.block_n, L.kc);
  return RCU_ENOTSUP;
}

// ---------------------------------------------------------------------------------------------------------
// Halo-kernel weight images: a sequence of [c_out][64] bf16 tiles in the SWIZZLE_128B K-major shared-memory layout
// (16-byte chunk index XOR (row & 7)), in the order the kernel's MMA loop walks them ([phase][chunk][tap]; two taps per tile for a
// 32-channel source), copied verbatim into shared memory by the kernel.
// ---------------------------------------------------------------------------------------------------------
static inline size_t sw128_index(int n, int k) {  // element index inside a tile
  return (size_t)n * 64 + (size_t)((((k >> 3) ^ (n & 7)) << 3) + (k & 7));
}

constexpr uint32_t kHaloMaxWeightBytes = 150 * 1024;

// plain 3x3 conv: c_in == 32 (pair source) or c_in in {64, 128}
static int pack_halo_conv3x3(rcu_unet* net, const rcu_conv_unit& u, HaloPack& hp) {
  const int N = u.c_out;
  if (!(N == 32 || N == 64)) return RCU_OK;
  std::vector<uint16_t> img;
  int n_tiles = 0;
  auto new_tile = [&]() { img.resize((size_t)(n_tiles + 1) * N * 64, 0); return n_tiles++; };
  if (u.c_in == 32) {
    // 64-byte pixel rows: two taps share one [c_out][64] weight tile (k < 32: even tap, k >= 32: odd tap)
    hp.pair = true; hp.n_chunks = 1;
    for (int tap = 0; tap < 9; ++tap) {
      if ((tap & 1) == 0) new_tile();
      const int t = n_tiles - 1;
      for (int n = 0; n < N; ++n)
        for (int k = 0; k < 32; ++k)
          img[(size_t)t * N * 64 + sw128_index(n, (tap & 1) * 32 + k)] = f32_to_bf16_rn(u.weight[((size_t)n * 32 + k) * 9 + tap]);
    }
  } else if (u.c_in == 64 || u.c_in == 128) {
    hp.pair = false; hp.n_chunks = u.c_in / 64;
    for (int j = 0; j < hp.n_chunks; ++j) {
      for (int tap = 0; tap < 9; ++tap) {
        const int t = new_tile();
        for (int n = 0; n < N; ++n)
          for (int k = 0; k < 64; ++k)
            img[(size_t)t * N * 64 + sw128_index(n, k)] = f32_to_bf16_rn(u.weight[((size_t)n * u.c_in + j * 64 + k) * 9 + tap]);
      }
    }
  } else {
    return RCU_OK;
  }
  hp.w_bytes = (uint32_t)(img.size() * 2);
  if (hp.w_bytes > kHaloMaxWeightBytes) return RCU_OK;
  uint16_t* d;
  int rc = dev_upload(net, img, &d);
  if (rc) return rc;
  hp.d_wimg[0] = reinterpret_cast<uint8_t*>(d);
  hp.n_phases = 1;
  hp.ok = true;
  return RCU_OK;
}

// Pixel-pair rows (conv_halo.cuh, HALO_PAIR32 / HALO_PAIR64): GEMM row = the two pixels (2i, 2i+1) of a dense NHWC source
// viewed as [h][w/2][2c]; output column n = a_o * 32 + c_o.  Per (chunk j, dy) the image holds
//   T0 [64 n][64 k]  pair tap di = 0: element = W[dy][dx = a_in - a_o][c_o][c_i]   (every combination is inside the 3x3 kernel)
//   S  [32 n][64 k]  the two side taps, n = c_o: a_in = 0 columns carry W[dy][+1] (right neighbour pair -> a_o = 1),
//                    a_in = 1 columns carry W[dy][-1] (left neighbour pair -> a_o = 0)
// with k = a_in * 32 + c_i for a 32-channel source (one chunk) and k = c_i, a_in = j for a 64-channel source (two chunks).
static int pack_halo_pair(rcu_unet* net, const rcu_conv_unit& u, HaloPack& hp) {
  const char* e = std::getenv("RCU_HALO_PAIR");
  if ((e && e[0] == '0') || u.c_out != 32 || !(u.c_in == 32 || u.c_in == 64)) return RCU_OK;
  const int nj = u.c_in / 32;
  auto W = [&](int co, int ci, int dy, int dx) { return u.weight[((size_t)co * u.c_in + ci) * 9 + (dy + 1) * 3 + (dx + 1)]; };
  std::vector<uint16_t> img((size_t)nj * 3 * 96 * 64, 0);
  for (int j = 0; j < nj; ++j)
    for (int dyi = 0; dyi < 3; ++dyi) {
      uint16_t* t0 = img.data() + (size_t)(j * 3 + dyi) * 96 * 64;
      uint16_t* sd = t0 + 64 * 64;
      for (int k = 0; k < 64; ++k) {
        const int a_in = nj == 1 ? (k >> 5) : j, ci = nj == 1 ? (k & 31) : k;
        for (int n = 0; n < 64; ++n) t0[sw128_index(n, k)] = f32_to_bf16_rn(W(n & 31, ci, dyi - 1, a_in - (n >> 5)));
        for (int n = 0; n < 32; ++n) sd[sw128_index(n, k)] = f32_to_bf16_rn(W(n, ci, dyi - 1, a_in == 0 ? 1 : -1));
      }
    }
  hp.pair_mode = nj; hp.pair = false; hp.n_chunks = nj; hp.n_phases = 1;
  hp.w_bytes = (uint32_t)(img.size() * 2);
  uint16_t* d;
  int rc = dev_upload(net, img, &d);
  if (rc) return rc;
  hp.d_wimg[0] = reinterpret_cast<uint8_t*>(d);
  hp.ok = true;
  return RCU_OK;
}

// nearest-x2 + conv3x3 as four 2x2-tap phase convolutions (see pack_upconv_phases): one weight image per phase
constexpr int kHaloGroups = 4;   // TMEM accumulator stages = epilogue warp groups
static size_t halo_chunk_stride_bytes() { return ((size_t)kHaloRows * kHaloPitch * 128 + 1023) & ~size_t(1023); }

static int pack_halo_upconv(rcu_unet* net, const rcu_conv_unit& u, HaloPack& hp) {
  const int N = u.c_out;
  if (!(N == 32 || N == 64) || !(u.c_in == 64 || u.c_in == 128)) return RCU_OK;
  hp.pair = false; hp.n_chunks = u.c_in / 64; hp.n_phases = 4;
  hp.w_bytes = (uint32_t)(hp.n_chunks * 4 * N * 128);
  if (hp.w_bytes > kHaloMaxWeightBytes) return RCU_OK;
  // as many phases per launch as TMEM (4 stages x PH x N <= 512 columns) and shared memory (weights + >= 3 halo slots) allow
  hp.phases_per_launch = 1;
  for (int phs = 4; phs >= 2; phs >>= 1)
    if (kHaloGroups * phs * N <= 512 && (size_t)phs * hp.w_bytes + 3 * halo_chunk_stride_bytes() + 4096 <= (size_t)kHaloSmemBudget) { hp.phases_per_launch = phs; break; }
  std::vector<uint16_t> all((size_t)4 * hp.n_chunks * 4 * N * 64, 0);
  // tap (i2, j2) of phase (a, b) reads window row a + i2, column b + j2; its weight is the sum of the 3x3 taps that
  // nearest-x2 folds onto that low-resolution pixel (see pack_upconv_phases)
  auto phase_weight = [&](int a, int b, int i2, int j2, int n, int ci) {
    int ky0, ky1, kx0, kx1;
    if (a == 0) { ky0 = i2 == 0 ? 0 : 1; ky1 = i2 == 0 ? 0 : 2; } else { ky0 = i2 == 0 ? 0 : 2; ky1 = i2 == 0 ? 1 : 2; }
    if (b == 0) { kx0 = j2 == 0 ? 0 : 1; kx1 = j2 == 0 ? 0 : 2; } else { kx0 = j2 == 0 ? 0 : 2; kx1 = j2 == 0 ? 1 : 2; }
    float sum = 0.0f;
    for (int ky = ky0; ky <= ky1; ++ky)
      for (int kx = kx0; kx <= kx1; ++kx) sum += u.weight[((size_t)n * u.c_in + ci) * 9 + ky * 3 + kx];
    return f32_to_bf16_rn(sum);
  };
  if (hp.phases_per_launch >= 2 && RCU_UP_PAIRED) {
    // x-paired layout (conv_halo.cuh, HALO_UP64 with PH >= 2): per (a, chunk, i2) a [2N][64] tile for window column 1
    // (rows 0..N-1: phase (a, 0) tap j2 = 1, rows N..2N-1: phase (a, 1) tap j2 = 0), then [N][64] for column 0 (phase (a, 0),
    // j2 = 0) and [N][64] for column 2 (phase (a, 1), j2 = 1).  Same bytes as the per-phase layout: 4N rows per (a, chunk, i2).
    for (int a = 0; a < 2; ++a)
      for (int j = 0; j < hp.n_chunks; ++j)
        for (int i2 = 0; i2 < 2; ++i2) {
          uint16_t* blk = all.data() + ((size_t)((a * hp.n_chunks + j) * 2 + i2)) * 4 * N * 64;
          for (int k = 0; k < 64; ++k) {
            const int ci = j * 64 + k;
            for (int n = 0; n < 2 * N; ++n) blk[sw128_index(n, k)] = phase_weight(a, n / N, i2, n < N ? 1 : 0, n % N, ci);
            for (int n = 0; n < N; ++n) {
              blk[(size_t)2 * N * 64 + sw128_index(n, k)] = phase_weight(a, 0, i2, 0, n, ci);
              blk[(size_t)3 * N * 64 + sw128_index(n, k)] = phase_weight(a, 1, i2, 1, n, ci);
            }
          }
        }
    uint16_t* dp;
    int rcp = dev_upload(net, all, &dp);
    if (rcp) return rcp;
    for (int ph = 0; ph < 4; ++ph) hp.d_wimg[ph] = reinterpret_cast<uint8_t*>(dp) + (size_t)ph * hp.w_bytes;
    hp.ok = true;
    return RCU_OK;
  }
  for (int a = 0; a < 2; ++a)
    for (int b = 0; b < 2; ++b) {
      const int ph = a * 2 + b;
      uint16_t* img = all.data() + (size_t)ph * hp.n_chunks * 4 * N * 64;
      int t = 0;
      for (int j = 0; j < hp.n_chunks; ++j) {
        for (int i2 = 0; i2 < 2; ++i2)
          for (int j2 = 0; j2 < 2; ++j2) {
            // tap (i2, j2) of phase (a, b) reads window row a + i2, column b + j2 (see pack_upconv_phases for the kernel sums)
            int ky0, ky1, kx0, kx1;
            if (a == 0) { ky0 = i2 == 0 ? 0 : 1; ky1 = i2 == 0 ? 0 : 2; } else { ky0 = i2 == 0 ? 0 : 2; ky1 = i2 == 0 ? 1 : 2; }
            if (b == 0) { kx0 = j2 == 0 ? 0 : 1; kx1 = j2 == 0 ? 0 : 2; } else { kx0 = j2 == 0 ? 0 : 2; kx1 = j2 == 0 ? 1 : 2; }
            for (int n = 0; n < N; ++n)
              for (int k = 0; k < 64; ++k) {
                float sum = 0.0f;
                for (int ky = ky0; ky <= ky1; ++ky)
                  for (int kx = kx0; kx <= kx1; ++kx) sum += u.weight[((size_t)n * u.c_in + j * 64 + k) * 9 + ky * 3 + kx];
                img[(size_t)t * N * 64 + sw128_index(n, k)] = f32_to_bf16_rn(sum);
              }
            ++t;
          }
      }
    }
  uint16_t* d;
  int rc = dev_upload(net, all, &d);
  if (rc) return rc;
  for (int ph = 0; ph < 4; ++ph) hp.d_wimg[ph] = reinterpret_cast<uint8_t*>(d) + (size_t)ph * hp.w_bytes;
  hp.ok = true;
  return RCU_OK;
}

// Wide-kernel weight image: [phase][ntile][chunk][tap] tiles of [128][64] bf16 in the SWIZZLE_128B smem layout.
static int pack_wide(rcu_unet* net, const rcu_conv_unit& u, bool upconv, WidePack& wp) {
  if (u.c_out % kWideN != 0 || u.c_in % 64 != 0) return RCU_OK;
  wp.n_chunks = u.c_in / 64; wp.n_ntiles = u.c_out / kWideN; wp.n_phases = upconv ? 4 : 1;
  const int taps = upconv ? 4 : 9;
  std::vector<uint16_t> img((size_t)wp.n_phases * wp.n_ntiles * wp.n_chunks * taps * kWideN * 64);
  size_t tile = 0;
  for (int ph = 0; ph < wp.n_phases; ++ph) {
    const int a = ph >> 1, b = ph & 1;
    for (int nt = 0; nt < wp.n_ntiles; ++nt)
      for (int c = 0; c < wp.n_chunks; ++c)
        for (int tap = 0; tap < taps; ++tap, ++tile) {
          int ky0 = tap / 3, ky1 = tap / 3, kx0 = tap % 3, kx1 = tap % 3;
          if (upconv) {   // pre-summed 2x2 phase weights, see pack_upconv_phases
            const int i2 = tap >> 1, j2 = tap & 1;
            if (a == 0) { ky0 = i2 == 0 ? 0 : 1; ky1 = i2 == 0 ? 0 : 2; } else { ky0 = i2 == 0 ? 0 : 2; ky1 = i2 == 0 ? 1 : 2; }
            if (b == 0) { kx0 = j2 == 0 ? 0 : 1; kx1 = j2 == 0 ? 0 : 2; } else { kx0 = j2 == 0 ? 0 : 2; kx1 = j2 == 0 ? 1 : 2; }
          }
          uint16_t* t = img.data() + tile * kWideN * 64;
          for (int n = 0; n < kWideN; ++n)
            for (int k = 0; k < 64; ++k) {
              float sum = 0.0f;
              for (int ky = ky0; ky <= ky1; ++ky)
                for (int kx = kx0; kx <= kx1; ++kx) sum += u.weight[((size_t)(nt * kWideN + n) * u.c_in + c * 64 + k) * 9 + ky * 3 + kx];
              t[sw128_index(n, k)] = f32_to_bf16_rn(sum);
            }
        }
  }
  uint16_t* d;
  int rc = dev_upload(net, img, &d);
  if (rc) return rc;
  wp.d_wimg = reinterpret_cast<uint8_t*>(d);
  wp.ok = true;
  return RCU_OK;
}

static int make_wide_map(CUtensorMap* map, const Act& a, int n) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available (driver too old?)"); return RCU_ECUDA; }
  cuuint64_t dims[4] = {(cuuint64_t)a.c, (cuuint64_t)a.w, (cuuint64_t)a.h, (cuuint64_t)n};
  cuuint64_t strides[3] = {(cuuint64_t)a.c_total * 2, (cuuint64_t)a.w * a.c_total * 2, (cuuint64_t)a.img_stride * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)kWidePitch, (cuuint32_t)kWidePitch, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, a.base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(wide c=%d w=%d h=%d n=%d) failed: %d", a.c, a.w, a.h, n, (int)r); return RCU_ECUDA; }
  return RCU_OK;
}

// Launch with programmatic stream serialization (the kernels call griddepcontrol.wait before their first dependent
// access, see conv_halo.cuh).  RCU_PDL=0 falls back to plain stream order.
template <typename Kern, typename... Args>
static int launch_pdl(Kern kern, unsigned grid, unsigned threads, size_t smem, cudaStream_t st, const Args&... args) {
  static const bool pdl = [] { const char* e = std::getenv("RCU_PDL"); return !(e && e[0] == '0'); }();
  cudaLaunchConfig_t cfg;
  std::memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  RCU_CUDA(cudaLaunchKernelEx(&cfg, kern, args...));
  return RCU_OK;
}

template <int TAPS>
static int launch_conv_wide(const ConvLayer& L, const WideParams& prm, const HaloOutMaps& maps, cudaStream_t st) {
  auto kern = conv_wide_kernel<TAPS>;
  using C = WideCfg<TAPS>;
  static bool configured[64] = {false};
  int dev = 0;
  RCU_CUDA(cudaGetDevice(&dev));
  dev = dev < 0 || dev >= 64 ? 0 : dev;
  if (!configured[dev]) {
    RCU_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmem));
    configured[dev] = true;
  }
  const long long total = (long long)prm.n_img * prm.tiles_y * prm.tiles_x * prm.n_ntiles * prm.n_phases;
  long long grid = sm_count();
  if (grid > total) grid = total;
  if (grid < 1) return RCU_OK;
  return launch_pdl(kern, (unsigned)grid, kWideThreads, (size_t)C::kSmem, st, L.map_wide, maps, prm);
}

static uint32_t halo_chunk_stride(bool /*half_rows*/) {
  const uint32_t bytes = (uint32_t)(kHaloRows * kHaloPitch * 128);   // footprint: 128-byte rows even when only 64 are filled
  return (bytes + 1023u) & ~1023u;
}

template <int N>
static int halo_stage_count(uint32_t w_bytes, bool pair, bool tma_store = false) {
  const int64_t room = (int64_t)kHaloSmemBudget - HaloSmem<N, kHaloGroups>::kFixed - (int64_t)((w_bytes + 1023u) & ~1023u) -
                       (tma_store ? HaloSmem<N, kHaloGroups>::kOutBytes : 0);
  int64_t st = room / halo_chunk_stride(pair);
  if (st > HaloSmem<N, kHaloGroups>::kMaxStages) st = HaloSmem<N, kHaloGroups>::kMaxStages;
  return (int)st;
}

template <int N, int MODE, int PH = 1>
static int launch_conv_halo(const ConvLayer& L, const HaloParams& prm, const HaloOutMaps& maps, cudaStream_t st) {
  auto kern = conv_halo_kernel<N, kHaloGroups, MODE, PH>;
  using HS = HaloSmem<N, kHaloGroups, PH>;
  static bool configured[64] = {false};
  int dev = 0;
  RCU_CUDA(cudaGetDevice(&dev));
  dev = dev < 0 || dev >= 64 ? 0 : dev;
  if (!configured[dev]) {
    RCU_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kHaloSmemBudget));
    configured[dev] = true;
  }
  const size_t smem = (size_t)HS::kFixed + ((prm.w_bytes + 1023u) & ~1023u) + (size_t)prm.n_stages * prm.chunk_stride +
                      (prm.tma_store ? HS::kOutBytes : 0);
  const long long total_tiles = (long long)prm.n_img * prm.tiles_y * prm.tiles_x;
  long long grid = sm_count();
  if (grid > total_tiles) grid = total_tiles;
  if (grid < 1) return RCU_OK;
  return launch_pdl(kern, (unsigned)grid, HS::kThreads + halo_patch_threads(MODE), smem, st, L.map_halo, maps, prm);
}

}  // namespace rcu

extern "C" int rcu_unet_create(const rcu_unet_desc* d, int device, rcu_unet** out) {
  RCU_CHECK_ARG(d != nullptr && out != nullptr, "NULL argument");
  *out = nullptr;
  int rc = rcu_device_check(device);
  if (rc) return rc;
  if (d->nb_classes != 2) { set_error("nb_classes=%d: the hot path is binary (2 classes)", d->nb_classes); return RCU_ENOTSUP; }
  RCU_CHECK_ARG(d->depth >= 1 && d->depth <= 6, "depth %d out of range", d->depth);
  RCU_CHECK_ARG(d->n_units == 4 * d->depth + 3, "expected %d conv units for depth %d, got %d", 4 * d->depth + 3, d->depth, d->n_units);
  RCU_CHECK_ARG(d->n_upconvs == d->depth, "expected %d upconvs, got %d", d->depth, d->n_upconvs);
  RCU_CHECK_ARG(d->in_channels >= 1 && d->in_channels <= 8, "in_channels %d out of range [1, 8]", d->in_channels);
  if (d->start_filters != 32 && d->start_filters != 64) { set_error("start_filters=%d: supported 32 or 64", d->start_filters); return RCU_ENOTSUP; }
  RCU_CHECK_ARG(d->p_drop >= 0.f && d->p_drop < 1.f, "dropout p out of range");
  RCU_CUDA(cudaSetDevice(device));

  rcu_unet* net = new rcu_unet();
  net->device = device;
  net->in_channels = d->in_channels;
  net->depth = d->depth;
  net->start_filters = d->start_filters;
  net->p_drop = d->p_drop;
  const int sf = d->start_filters;

  // ---- column tables (units first, then upconvs) ----
  std::vector<float> fa, fba, fd;
  std::vector<int> site, chin, scol;
  // sigma_out nets (unet.py:162-164) carry one more Conv2dBnRelu + 1x1 head on the same features; its unit (and its
  // Dropout2d site) come after conv_cls.0, the order the reference's forward visits them in (unet.py:181-185)
  const bool residual = d->residuals != nullptr;
  if (residual && d->n_residuals != 2 * d->depth + 1) { set_error("expected %d residual convs for depth %d, got %d", 2 * d->depth + 1, d->depth, d->n_residuals); delete net; return RCU_EINVAL; }
  net->residual = residual;
  const bool has_sigma = d->sigma_unit != nullptr;
  if (has_sigma && d->sigma_head == nullptr) { set_error("sigma_unit without sigma_head"); rcu_unet_destroy(net); return RCU_EINVAL; }
  net->has_sigma = has_sigma;
  const int n_all_units = d->n_units + (has_sigma ? 1 : 0);
  auto unit_at = [&](int u) -> const rcu_conv_unit& { return u < d->n_units ? d->units[u] : *d->sigma_unit; };
  std::vector<int> unit_off(n_all_units), up_off(d->n_upconvs);
  int n_sites = 0, drop_cols = 0;
  for (int u = 0; u < n_all_units; ++u) {
    const rcu_conv_unit& cu = unit_at(u);
    if (!(cu.weight && cu.bias && cu.bn_weight && cu.bn_bias && cu.bn_mean && cu.bn_var)) {
      set_error("unit %d: NULL weight/bn pointer (bn=False nets are outside the hot path)", u);
      rcu_unet_destroy(net);
      return RCU_EINVAL;
    }
    unit_off[u] = (int)fa.size();
    fold_unit(cu, d->bn_eps, fa, fba, fd);
    for (int c = 0; c < cu.c_out; ++c) {
      site.push_back(cu.has_dropout ? n_sites : -1);
      chin.push_back(c);
      scol.push_back(cu.has_dropout ? drop_cols + c : 0);
    }
    if (cu.has_dropout) {
      net->site_channels.push_back(cu.c_out);
      ++n_sites;
      drop_cols += cu.c_out;
    }
  }
  for (int u = 0; u < d->n_upconvs; ++u) {
    up_off[u] = (int)fa.size();
    fold_unit(d->upconvs[u], d->bn_eps, fa, fba, fd);
    for (int c = 0; c < d->upconvs[u].c_out; ++c) { site.push_back(-1); chin.push_back(c); scol.push_back(0); }
  }
  std::vector<int> res_off(residual ? d->n_residuals : 0);
  for (int r = 0; r < (int)res_off.size(); ++r) {   // bias-only columns, like the upconvs
    const rcu_conv_unit& ru = d->residuals[r];
    if (!(ru.weight && ru.bias)) { set_error("residual %d: NULL weight / bias", r); rcu_unet_destroy(net); return RCU_EINVAL; }
    res_off[r] = (int)fa.size();
    fold_unit(ru, d->bn_eps, fa, fba, fd);
    for (int c = 0; c < ru.c_out; ++c) { site.push_back(-1); chin.push_back(c); scol.push_back(0); }
  }
  net->n_sites = n_sites;
  net->total_dropout_channels = drop_cols;
  net->n_cols = (int)fa.size();

#define RCU_TRY(expr) do { int rc__ = (expr); if (rc__) { rcu_unet_destroy(net); return rc__; } } while (0)
  float *dfa, *dfba, *dfd; int *dsite, *dchin, *dscol;
  RCU_TRY(dev_upload(net, fa, &dfa)); RCU_TRY(dev_upload(net, fba, &dfba)); RCU_TRY(dev_upload(net, fd, &dfd));
  RCU_TRY(dev_upload(net, site, &dsite)); RCU_TRY(dev_upload(net, chin, &dchin)); RCU_TRY(dev_upload(net, scol, &dscol));
  net->cols = CoefColumns{dfa, dfba, dfd, dsite, dchin, dscol, net->n_cols};

  // ---- first conv: fp32 [c_in*9][sf] ----
  {
    const rcu_conv_unit& u0 = d->units[0];
    if (u0.c_in != d->in_channels || u0.c_out != sf) { set_error("unit 0 must map in_channels -> start_filters"); rcu_unet_destroy(net); return RCU_EINVAL; }
    std::vector<float> w((size_t)u0.c_in * 9 * sf);
    for (int co = 0; co < sf; ++co)
      for (int ci = 0; ci < u0.c_in; ++ci)
        for (int k = 0; k < 9; ++k) w[((size_t)ci * 9 + k) * sf + co] = u0.weight[((size_t)co * u0.c_in + ci) * 9 + k];
    RCU_TRY(dev_upload(net, w, &net->d_first_w));
    net->first_coef_off = unit_off[0];
    // the two sample-independent coefficient rows of the first unit, in coef_kernel's own arithmetic (s = 1 and s = 1 / (1 - p))
    net->unit0_dropout = u0.has_dropout != 0;
    const float inv_keep0 = 1.0f / (1.0f - d->p_drop);
    std::vector<float2> e0((size_t)2 * sf);
    for (int c = 0; c < sf; ++c) {
      const float a = fa[unit_off[0] + c], ba = fba[unit_off[0] + c], dd = fd[unit_off[0] + c];
      e0[c] = make_float2(a * 1.0f, std::fmaf(ba, 1.0f, dd));
      e0[sf + c] = make_float2(a * inv_keep0, std::fmaf(ba, inv_keep0, dd));
    }
    RCU_TRY(dev_upload(net, e0, &net->d_e0coef));
  }
  // ---- 1x1 heads ----
  std::vector<float> host_head[2];
  auto upload_head = [&](const rcu_conv_unit& h, float** out_ptr, int slot) -> int {
    if (!(h.weight && h.bias) || h.c_in != sf || h.c_out != 2) { set_error("head must be a 1x1 conv start_filters -> 2"); return RCU_EINVAL; }
    std::vector<float> hw((size_t)2 * sf + 2);
    for (int i = 0; i < 2 * sf; ++i) hw[i] = h.weight[i];
    hw[2 * sf] = h.bias[0];
    hw[2 * sf + 1] = h.bias[1];
    host_head[slot] = hw;
    return dev_upload(net, hw, out_ptr);
  };
  RCU_TRY(upload_head(d->head, &net->d_head, 0));
  if (has_sigma) RCU_TRY(upload_head(*d->sigma_head, &net->d_sigma_head, 1));
  // ---- tensor-core conv layers in execution order ----
  auto add_unit = [&](int u, int c0, int c1, bool head, int out_slot = 0, int relu = 1) -> int {
    const rcu_conv_unit& cu = unit_at(u);
    if (cu.c_in != c0 + c1) { set_error("unit %d: c_in=%d does not match the topology (%d)", u, cu.c_in, c0 + c1); return RCU_EINVAL; }
    ConvLayer L;
    L.c0 = c0; L.c1 = c1; L.c_out = cu.c_out; L.n_taps = 9; L.n_phases = 1; L.out_mul = 1; L.relu = relu; L.head = head;
    for (int t = 0; t < 9; ++t) { L.dy[0][t] = (signed char)(t / 3 - 1); L.dx[0][t] = (signed char)(t % 3 - 1); }
    L.coef_off = unit_off[u];
    L.out_slot = out_slot;
    L.d_head = out_slot ? net->d_sigma_head : net->d_head;
    if (head && sf == 32) std::memcpy(L.h_head, host_head[out_slot ? 1 : 0].data(), sizeof(float) * 66);
    int rc2 = pick_tiles(c1 > 0 ? (c0 < c1 ? c0 : c1) : c0, cu.c_out, &L.block_n, &L.kc);
    if (rc2) return rc2;
    if (head && (L.block_n != 32 || cu.c_out != 32)) { set_error("fused head needs start_filters == 32"); return RCU_ENOTSUP; }
    std::vector<uint16_t> w = pack_conv3x3(cu);
    uint16_t* dw;
    rc2 = dev_upload(net, w, &dw);
    if (rc2) return rc2;
    L.d_weights = reinterpret_cast<__nv_bfloat16*>(dw);
    rc2 = pack_halo_conv3x3(net, cu, L.halo);
    if (rc2) return rc2;
    rc2 = pack_halo_pair(net, cu, L.pairp);
    if (rc2) return rc2;
    rc2 = pack_wide(net, cu, false, L.wide);
    if (rc2) return rc2;
    net->convs.push_back(L);
    return RCU_OK;
  };
  auto add_upconv = [&](int j, int c_in, int c_out) -> int {
    const rcu_conv_unit& cu = d->upconvs[j];
    if (cu.c_in != c_in || cu.c_out != c_out || !cu.weight || !cu.bias) { set_error("upconv %d: shape mismatch", j); return RCU_EINVAL; }
    ConvLayer L;
    L.c0 = c_in; L.c1 = 0; L.c_out = c_out; L.n_taps = 4; L.n_phases = 4; L.out_mul = 2; L.relu = 0;
    L.coef_off = up_off[j];
    int rc2 = pick_tiles(c_in, c_out, &L.block_n, &L.kc);
    if (rc2) return rc2;
    std::vector<uint16_t> w = pack_upconv_phases(cu, L);
    uint16_t* dw;
    rc2 = dev_upload(net, w, &dw);
    if (rc2) return rc2;
    L.d_weights = reinterpret_cast<__nv_bfloat16*>(dw);
    rc2 = pack_halo_upconv(net, cu, L.halo);
    if (rc2) return rc2;
    rc2 = pack_wide(net, cu, true, L.wide);
    if (rc2) return rc2;
    net->convs.push_back(L);
    return RCU_OK;
  };
  // residual branch of block r (forward order: down 0.., bottom, up 0..): block output += conv1x1(block input) + bias,
  // and the block's last Conv2dBnRelu has no activation (unet.py:52-53)
  auto add_residual = [&](int r, int c_in, int c_out) -> int {
    const rcu_conv_unit& ru = d->residuals[r];
    if (ru.c_in != c_in || ru.c_out != c_out) { set_error("residual %d: %d -> %d channels, topology says %d -> %d", r, ru.c_in, ru.c_out, c_in, c_out); return RCU_EINVAL; }
    if (r == 0) {   // the block input is the float32 image: CUDA-core kernel (first_residual_kernel)
      std::vector<float> w(ru.weight, ru.weight + (size_t)c_out * c_in), b(ru.bias, ru.bias + c_out);
      int rc2 = dev_upload(net, w, &net->d_res0_w);
      if (rc2) return rc2;
      return dev_upload(net, b, &net->d_res0_b);
    }
    ConvLayer L;
    L.c0 = c_in; L.c1 = 0; L.c_out = c_out; L.n_taps = 1; L.n_phases = 1; L.out_mul = 1; L.relu = 0; L.accumulate = true;
    L.dy[0][0] = 0; L.dx[0][0] = 0;
    L.coef_off = res_off[r];
    int rc2 = pick_tiles(c_in, c_out, &L.block_n, &L.kc);
    if (rc2) return rc2;
    std::vector<uint16_t> w((size_t)c_out * c_in);
    for (size_t i = 0; i < w.size(); ++i) w[i] = f32_to_bf16_rn(ru.weight[i]);   // [1 tap][c_out][c_in]
    uint16_t* dw;
    rc2 = dev_upload(net, w, &dw);
    if (rc2) return rc2;
    L.d_weights = reinterpret_cast<__nv_bfloat16*>(dw);
    net->convs.push_back(L);
    return RCU_OK;
  };
  {
    int u = 1;  // unit 0 is the first conv
    int r = 0;
    int c = sf;
    const int last_relu = residual ? 0 : 1;
    RCU_TRY(add_unit(u++, c, 0, false, 0, last_relu));   // down 0, second conv
    if (residual) RCU_TRY(add_residual(r++, d->in_channels, c));
    for (int l = 1; l < d->depth; ++l) {
      RCU_TRY(add_unit(u++, c, 0, false));               // c -> 2c
      c *= 2;
      RCU_TRY(add_unit(u++, c, 0, false, 0, last_relu));
      if (residual) RCU_TRY(add_residual(r++, c / 2, c));
    }
    RCU_TRY(add_unit(u++, c, 0, false));                 // bottom
    c *= 2;
    RCU_TRY(add_unit(u++, c, 0, false, 0, last_relu));
    if (residual) RCU_TRY(add_residual(r++, c / 2, c));
    for (int j = 0; j < d->depth; ++j) {
      RCU_TRY(add_upconv(j, c, c / 2));
      c /= 2;
      RCU_TRY(add_unit(u++, 2 * c, 0, false));           // cat((up, skip), 1): one 2c-channel concat buffer
      RCU_TRY(add_unit(u++, c, 0, false, 0, last_relu));
      if (residual) RCU_TRY(add_residual(r++, 2 * c, c));
    }
    const bool fuse_head = (sf == 32);
    RCU_TRY(add_unit(u++, c, 0, fuse_head));             // conv_cls.0
    if (has_sigma) RCU_TRY(add_unit(u++, c, 0, fuse_head, 1));   // conv_sigma.0
  }
#undef RCU_TRY
  *out = net;
  return RCU_OK;
}

extern "C" void rcu_unet_destroy(rcu_unet* net) {
  if (!net) return;
  cudaSetDevice(net->device);
  for (auto& e : net->ev_pool) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
  for (void* p : net->owned) cudaFree(p);
  delete net;
}

extern "C" int rcu_unet_total_dropout_channels(const rcu_unet* net) { return net ? net->total_dropout_channels : 0; }
extern "C" int64_t rcu_unet_last_launch_count(const rcu_unet* net) { return net ? net->last_launches : 0; }
extern "C" int rcu_unet_set_conv_impl(rcu_unet* net, int impl) {
  RCU_CHECK_ARG(net != nullptr, "NULL handle");
  RCU_CHECK_ARG(impl >= 0 && impl <= 3, "conv impl must be 0 (tcgen05), 1 (cross-check), 2 (tcgen05, per-tap kernel only) or 3 (tcgen05 without the pixel-pair kernel)");
  net->pair_enabled = impl != 3;
  net->conv_impl = impl == 3 ? 0 : impl;
  return RCU_OK;
}

extern "C" int rcu_unet_set_first_layer_dedup(rcu_unet* net, int enable) {
  RCU_CHECK_ARG(net != nullptr, "NULL handle");
  net->first_dedup = enable != 0;
  return RCU_OK;
}

extern "C" int rcu_unet_set_halo_mask(rcu_unet* net, uint64_t mask) {
  RCU_CHECK_ARG(net != nullptr, "NULL handle");
  net->halo_mask = mask;
  return RCU_OK;
}

// Sizes the activation workspace (base == nullptr) or lays the tensors out inside the caller's buffer, encodes the TMA
// descriptors (they embed addresses) and builds the schedule.
static int plan_impl(rcu_unet* net, int height, int width, int max_images_per_chunk, uint8_t* base_in, size_t* workspace_bytes);

extern "C" int rcu_unet_plan(rcu_unet* net, int height, int width, int max_images_per_chunk, size_t* workspace_bytes) {
  RCU_CHECK_ARG(net != nullptr, "NULL handle");
  size_t bytes = 0;
  int rc = plan_impl(net, height, width, max_images_per_chunk, nullptr, &bytes);
  if (rc) return rc;
  net->planned_bytes = bytes;
  if (workspace_bytes) *workspace_bytes = bytes;
  return RCU_OK;
}

extern "C" int rcu_unet_bind_workspace(rcu_unet* net, void* workspace, size_t workspace_bytes) {
  RCU_CHECK_ARG(net != nullptr, "NULL handle");
  RCU_CHECK_ARG(net->planned_bytes > 0, "rcu_unet_plan has not been called");
  RCU_CHECK_ARG(workspace != nullptr && reinterpret_cast<uintptr_t>(workspace) % 1024 == 0, "workspace must be a 1024-byte aligned device buffer");
  RCU_CHECK_ARG(workspace_bytes >= net->planned_bytes, "workspace of %zu bytes is smaller than the planned %zu", workspace_bytes, net->planned_bytes);
  size_t bytes = 0;
  return plan_impl(net, net->H, net->W, net->max_images, reinterpret_cast<uint8_t*>(workspace), &bytes);
}

static int plan_impl(rcu_unet* net, int height, int width, int max_images_per_chunk, uint8_t* base_in, size_t* workspace_bytes) {
  RCU_CHECK_ARG(net != nullptr, "NULL handle");
  const int div = 1 << net->depth;
  if (height % div != 0 || width % div != 0 || height < div || width < div) {
    set_error("spatial size %dx%d must be a multiple of 2^depth = %d (the reference's F.pad branch is not on the hot path)", height, width, div);
    return RCU_ENOTSUP;
  }
  RCU_CHECK_ARG(max_images_per_chunk >= 1, "max_images_per_chunk must be >= 1");
  RCU_CUDA(cudaSetDevice(net->device));
  net->arena = nullptr;   // the workspace belongs to the caller (rcu_unet_bind_workspace)
  net->ops.clear();
  net->H = height; net->W = width; net->max_images = max_images_per_chunk;
  const long long N = max_images_per_chunk;
  const int sf = net->start_filters, depth = net->depth;

  // ---- tensors.  Two passes over the same code: pass 0 sizes the arena, pass 1 hands out pointers.
  // Level l (resolution H>>l, C_l = sf<<l channels):  E[l] first conv of the encoder block, CAT[l] = [up | skip]
  // (2*C_l channels; the encoder's second conv writes the skip half, the decoder's upconv the up half), P[l] pooled
  // skip (input of level l+1), DA/DB[l] the decoder block's two convs.
  std::vector<Act> E(depth + 1), CAT(depth), P(depth), DA(depth), DB(depth);
  Act SB, HF, HF2;
  size_t off = 0;
  uint8_t* base = base_in;
  auto tensor = [&](int c_total, int h, int w) {
    Act a;
    a.c = a.c_total = c_total; a.h = h; a.w = w;
    a.padded = false;
    const long long rows = a.padded ? h + 1 : h;
    a.img_stride = rows * w * c_total;
    const size_t lead = a.padded ? (size_t)w * c_total * 2 : 0;   // the zero row above image 0
    const size_t bytes = lead + (size_t)N * a.img_stride * 2;
    if (base) a.base = reinterpret_cast<__nv_bfloat16*>(base + off + lead);
    off += (bytes + 1023) & ~size_t(1023);
    return a;
  };
  auto slice = [](const Act& t, int ch_off, int c) {
    Act a = t;
    a.base = t.base ? t.base + ch_off : nullptr;
    a.c = c;
    return a;
  };
  size_t o_coef = 0;
  {
    off = 0;
    for (int l = 0; l <= depth; ++l) {
      const int h = height >> l, w = width >> l, c = sf << l;
      E[l] = tensor(c, h, w);
      if (l < depth) {
        CAT[l] = tensor(2 * c, h, w);
        P[l] = tensor(c, h / 2, w / 2);
        DA[l] = tensor(c, h, w);
        DB[l] = tensor(c, h, w);
      } else {
        SB = tensor(c, h, w);
      }
    }
    HF = tensor(sf, height, width);   // conv_cls.0 features (cross-check path only; the tcgen05 paths fuse the head)
    if (net->has_sigma) HF2 = tensor(sf, height, width);
    o_coef = off;
    off += ((size_t)N * net->n_cols * sizeof(float2) + 1023) & ~size_t(1023);
  }
  if (workspace_bytes) *workspace_bytes = off;
  if (base == nullptr) return RCU_OK;   // sizing pass: nothing else depends on addresses
  net->arena = base;
  net->arena_bytes = off;
  net->d_coef = reinterpret_cast<float2*>(base + o_coef);
  net->first_out = E[0];
  net->head_feat = HF;
  net->sigma_feat = HF2;

  // ---- schedule ----
  int ci = 0;
  auto push_conv = [&](const Act& src, const Act& dst) -> int {
    ConvLayer& L = net->convs[ci];
    if (src.c != L.c0) { set_error("conv %d: source has %d channels, layer expects %d", ci, src.c, L.c0); return RCU_EINVAL; }
    L.src = src; L.dst = dst; L.in_h = src.h; L.in_w = src.w;
    int rc = make_act_map(&L.map_a0, src, (int)N, L.kc);
    if (rc) return rc;
    L.map_a1 = L.map_a0;
    rc = make_weight_map(&L.map_w, L.d_weights, L.c0 + L.c1, L.c_out, L.n_taps * L.n_phases, L.kc, L.block_n);
    if (rc) return rc;
    if (L.halo.ok) {
      const uint32_t resident = L.halo.w_bytes * (uint32_t)(L.halo.n_phases == 4 ? L.halo.phases_per_launch : 1);
      // TMA-store epilogue when the staging slots still leave a deep enough ring
      static const bool allow_tma_store = [] { const char* e = std::getenv("RCU_HALO_TMA_STORE"); return !(e && e[0] == '0'); }();
      const int with_out = L.c_out == 32 ? halo_stage_count<32>(resident, L.halo.pair, true) : halo_stage_count<64>(resident, L.halo.pair, true);
      L.tma_store = allow_tma_store && !L.head && with_out >= 2 * L.halo.n_chunks + 2;
      L.halo_stages = L.tma_store ? with_out
                                  : (L.c_out == 32 ? halo_stage_count<32>(resident, L.halo.pair) : halo_stage_count<64>(resident, L.halo.pair));
      if (L.halo_stages < 2) {
        L.halo.ok = false;
      } else {
        rc = make_halo_map(&L.map_halo, src, (int)N, L.halo.pair ? 32 : 64);
        if (rc) return rc;
        for (int ph = 0; ph < L.halo.n_phases && L.tma_store; ++ph) {
          rc = make_out_map(&L.map_out[ph], dst, (int)N, L.out_mul, L.halo.n_phases == 4 ? (ph >> 1) : 0, L.halo.n_phases == 4 ? (ph & 1) : 0);
          if (rc) return rc;
        }
      }
    }
    L.use_pair = false;
    if (L.pairp.ok && src.c == src.c_total && src.w % 2 == 0 && !src.padded && dst.w == src.w) {
      static const bool allow_tma_store = [] { const char* e = std::getenv("RCU_HALO_TMA_STORE"); return !(e && e[0] == '0'); }();
      const int nj = L.pairp.n_chunks;
      const int with_out = halo_stage_count<64>(L.pairp.w_bytes, false, true);
      L.pair_tma_store = allow_tma_store && !L.head && with_out >= 2 * nj + 1;
      L.pair_stages = L.pair_tma_store ? with_out : halo_stage_count<64>(L.pairp.w_bytes, false);
      if (L.pair_stages >= nj + 1) {
        Act pv = src;                       // [h][w/2][2c]
        pv.c = pv.c_total = 2 * src.c_total; pv.w = src.w / 2;
        rc = make_halo_map(&L.map_halo_pair, pv, (int)N, 64);
        if (rc) return rc;
        for (int half = 0; half < 2 && L.pair_tma_store; ++half) {
          Act dv = dst;                     // the even / odd pixels of the destination slice
          dv.base = dst.base + (size_t)half * dst.c_total; dv.c_total = 2 * dst.c_total; dv.w = dst.w / 2;
          rc = make_out_map(&L.map_out_pair[half], dv, (int)N, 1, 0, 0);
          if (rc) return rc;
        }
        L.use_pair = true;
      }
    }
    if (L.wide.ok) {
      rc = make_wide_map(&L.map_wide, src, (int)N);
      if (rc) return rc;
      if (!L.halo.ok)
        for (int ph = 0; ph < L.wide.n_phases; ++ph) {
          rc = make_out_map(&L.map_out[ph], dst, (int)N, L.out_mul, L.wide.n_phases == 4 ? (ph >> 1) : 0, L.wide.n_phases == 4 ? (ph & 1) : 0);
          if (rc) return rc;
        }
    }
    Op op;
    op.kind = OP_CONV; op.conv = ci; op.out = dst;
    op.h = src.h * L.out_mul; op.w = src.w * L.out_mul; op.c = L.c_out;
    net->ops.push_back(op);
    ++ci;
    return RCU_OK;
  };
  auto push_pool = [&](const Act& in, const Act& out) {
    Op op;
    op.kind = OP_POOL; op.in = in; op.out = out; op.in_h = in.h; op.in_w = in.w; op.h = in.h / 2; op.w = in.w / 2; op.c = in.c;
    net->ops.push_back(op);
  };
  {
    Op first;
    first.kind = OP_FIRST; first.out = E[0]; first.h = height; first.w = width; first.c = sf;
    net->ops.push_back(first);
  }
  for (ConvLayer& L : net->convs) L.pool_dst = Act();
  const bool res = net->residual;
  int rc = push_conv(E[0], slice(CAT[0], sf, sf));
  if (rc) return rc;
  // the pool that follows rides in this conv's epilogue (halo and wide kernels) — unless a residual branch is still to be added
  if (!res) net->convs[ci - 1].pool_dst = P[0];
  if (res) {
    net->res0_dst = slice(CAT[0], sf, sf);
    Op op;
    op.kind = OP_RES_FIRST; op.out = net->res0_dst; op.h = height; op.w = width; op.c = sf;
    net->ops.push_back(op);
  }
  for (int l = 1; l <= depth; ++l) {
    const int cp = sf << (l - 1), c = sf << l;
    push_pool(slice(CAT[l - 1], cp, cp), P[l - 1]);
    const Act block_out = l < depth ? slice(CAT[l], c, c) : SB;
    if ((rc = push_conv(P[l - 1], E[l]))) return rc;
    if ((rc = push_conv(E[l], block_out))) return rc;
    if (l < depth && !res) net->convs[ci - 1].pool_dst = P[l];
    if (res && (rc = push_conv(P[l - 1], block_out))) return rc;   // + residual(block input)
  }
  Act cur = SB;
  for (int l = depth - 1; l >= 0; --l) {
    const int c = sf << l;
    if ((rc = push_conv(cur, slice(CAT[l], 0, c)))) return rc;     // upconv phases: low-res in, high-res up half of the concat
    if ((rc = push_conv(CAT[l], DA[l]))) return rc;                 // conv over cat((up, skip), 1)
    if ((rc = push_conv(DA[l], DB[l]))) return rc;
    if (res && (rc = push_conv(CAT[l], DB[l]))) return rc;          // + residual(cat((up, skip), 1))
    cur = DB[l];
  }
  net->features = cur;
  if ((rc = push_conv(cur, net->head_feat))) return rc;             // conv_cls.0 (+ fused head)
  if (net->has_sigma && (rc = push_conv(cur, net->sigma_feat))) return rc;   // conv_sigma.0 (+ fused head)
  return RCU_OK;
}

namespace rcu {

static int run_conv_halo(rcu_unet* net, const ConvLayer& L, int n_img, int cs, long long s0, long long n_slices, float* logits, int head_diff,
                         cudaStream_t st, long long* launches) {
  const HaloPack& hp = L.halo;
  const int ppl = hp.n_phases == 4 ? hp.phases_per_launch : 1;
  for (int ph = 0; ph < hp.n_phases; ph += ppl) {
    HaloParams prm;
    std::memset(&prm, 0, sizeof(prm));
    prm.n_img = n_img;
    prm.in_h = L.in_h; prm.in_w = L.in_w;
    prm.tiles_x = (L.in_w + kHaloTileW - 1) / kHaloTileW;
    prm.tiles_y = (L.in_h + kHaloTileH - 1) / kHaloTileH;
    prm.n_chunks = hp.n_chunks;
    prm.chunk_bytes = (uint32_t)(kHaloRows * kHaloPitch * (hp.pair ? 64 : 128));   // bytes TMA delivers (complete_tx counts data bytes)
    prm.chunk_stride = halo_chunk_stride(hp.pair);
    prm.n_stages = L.halo_stages;
    prm.tma_store = L.tma_store ? 1 : 0;
    HaloOutMaps maps;
    for (int i = 0; i < ppl; ++i) maps.m[i] = L.map_out[ph + i];
    prm.tiles_per_turn = 1;   // measured: two tiles per turn is no faster (the hand-off is not what is exposed)
    {
      static const int tt = [] { const char* e = std::getenv("RCU_HALO_TT"); return e ? std::atoi(e) : 0; }();
      if (tt == 2 && L.halo_stages >= 4 * hp.n_chunks) prm.tiles_per_turn = 2;
    }
    for (int i = 0; i < ppl && hp.n_phases == 4; ++i) {
      const int a = (ph + i) >> 1, b = (ph + i) & 1;
      prm.up_base16[i] = (uint32_t)((a * kHaloPitch + b) * 8);
      prm.up_dy[i] = a; prm.up_dx[i] = b;
    }
    prm.w_image = hp.d_wimg[ph];
    prm.w_bytes = hp.w_bytes * (uint32_t)ppl;
    prm.out_mul = L.out_mul;
    prm.out_h = L.in_h * L.out_mul; prm.out_w = L.in_w * L.out_mul;
    prm.out_c = L.dst.c_total;
    prm.out_img_stride = L.dst.img_stride;
    prm.out = L.dst.base;
    prm.pool_out = L.pool_dst.base;
    prm.pool_img_stride = L.pool_dst.img_stride;
    prm.coef = net->d_coef; prm.coef_stride = net->n_cols; prm.coef_off = L.coef_off;
    prm.relu = L.relu;
    prm.head = L.head ? L.d_head : nullptr;
    if (L.head) std::memcpy(prm.head_w, L.h_head, sizeof(prm.head_w));
    prm.logits = logits;
    prm.head_diff = head_diff;
    prm.chunk_slices = cs; prm.slice0 = s0; prm.n_slices_total = n_slices;
    int rc;
    if (hp.n_phases == 4) {
      if (ppl == 4) rc = L.c_out == 32 ? launch_conv_halo<32, HALO_UP64, 4>(L, prm, maps, st) : RCU_ENOTSUP;
      else if (ppl == 2) rc = L.c_out == 32 ? launch_conv_halo<32, HALO_UP64, 2>(L, prm, maps, st) : launch_conv_halo<64, HALO_UP64, 2>(L, prm, maps, st);
      else rc = L.c_out == 32 ? launch_conv_halo<32, HALO_UP64>(L, prm, maps, st) : launch_conv_halo<64, HALO_UP64>(L, prm, maps, st);
    } else if (hp.pair) rc = L.c_out == 32 ? launch_conv_halo<32, HALO_CONV32>(L, prm, maps, st) : launch_conv_halo<64, HALO_CONV32>(L, prm, maps, st);
    else rc = L.c_out == 32 ? launch_conv_halo<32, HALO_CONV64>(L, prm, maps, st) : launch_conv_halo<64, HALO_CONV64>(L, prm, maps, st);
    if (rc) return rc;
    ++*launches;
  }
  return RCU_OK;
}

static int run_conv_halo_pair(rcu_unet* net, const ConvLayer& L, int n_img, int cs, long long s0, long long n_slices, float* logits, int head_diff,
                              int dedup_mode, cudaStream_t st, long long* launches) {
  const HaloPack& hp = L.pairp;
  HaloParams prm;
  std::memset(&prm, 0, sizeof(prm));
  prm.n_img = n_img;
  prm.in_h = L.in_h; prm.in_w = L.in_w / 2;                      // GEMM pixel grid = pixel pairs
  prm.tiles_x = (prm.in_w + kHaloTileW - 1) / kHaloTileW;
  prm.tiles_y = (L.in_h + kHaloTileH - 1) / kHaloTileH;
  prm.n_chunks = hp.n_chunks;
  prm.chunk_bytes = (uint32_t)(kHaloRows * kHaloPitch * 128);
  prm.chunk_stride = halo_chunk_stride(false);
  prm.n_stages = L.pair_stages;
  prm.tma_store = L.pair_tma_store ? 1 : 0;
  prm.tiles_per_turn = 1;
  {
    static const int tt = [] { const char* e = std::getenv("RCU_HALO_TT"); return e ? std::atoi(e) : 0; }();
    if (tt == 2 && L.pair_stages >= 4 * hp.n_chunks) prm.tiles_per_turn = 2;   // A/B only
  }
  // the first-layer consumer (patch warps + fused pool) is the one pair layer that gains from two tiles per issue turn
  // (6.75 -> 6.56 ms per step; the plain 32->32 layer and the head lose 3 - 6 %): profiles/r02zz_ab_experiments.log
  if (dedup_mode != 0 && L.pair_stages >= 4 * hp.n_chunks) prm.tiles_per_turn = 2;
  HaloOutMaps maps;
  maps.m[0] = L.map_out_pair[0]; maps.m[1] = L.map_out_pair[1];
  prm.w_image = hp.d_wimg[0];
  prm.w_bytes = hp.w_bytes;
  prm.out_mul = 1;
  prm.out_h = L.in_h; prm.out_w = L.in_w;                        // output addressing is in pixels
  prm.out_c = L.dst.c_total;
  prm.out_img_stride = L.dst.img_stride;
  prm.out = L.dst.base;
  prm.pool_out = L.pool_dst.base;
  prm.pool_img_stride = L.pool_dst.img_stride;
  prm.coef = net->d_coef; prm.coef_stride = net->n_cols; prm.coef_off = L.coef_off;
  prm.relu = L.relu;
  prm.head = L.head ? L.d_head : nullptr;
  if (L.head) std::memcpy(prm.head_w, L.h_head, sizeof(prm.head_w));
  prm.logits = logits;
  prm.head_diff = head_diff;
  prm.chunk_slices = cs; prm.slice0 = s0; prm.n_slices_total = n_slices;
  prm.dedup_mode = dedup_mode;
  prm.patch_coef = net->d_coef; prm.patch_off = net->first_coef_off;
  ConvLayer view = L;
  view.map_halo = L.map_halo_pair;
  // the first-layer consumer runs the instantiation with patch warps (736 threads, 80 registers); every other 32-channel
  // pair layer the plain one (608 threads, 96 registers)
  int rc;
  if (hp.pair_mode != 1) {
    RCU_CHECK_ARG(prm.head == nullptr, "the fused head follows a 32-channel pair layer");
    rc = launch_conv_halo<64, HALO_PAIR64>(view, prm, maps, st);
  } else if (prm.head != nullptr) {
    rc = launch_conv_halo<64, HALO_PAIR32_HEAD>(view, prm, maps, st);
  } else if (dedup_mode != 0) {
    rc = launch_conv_halo<64, HALO_PAIR32_PATCH>(view, prm, maps, st);
  } else {
    rc = launch_conv_halo<64, HALO_PAIR32>(view, prm, maps, st);
  }
  if (rc) return rc;
  ++*launches;
  return RCU_OK;
}

}  // namespace rcu

extern "C" int rcu_unet_forward(rcu_unet* net, const float* images, int64_t n_slices, int n_samples, int dropout_mode,
                                int det_first, uint64_t seed, int64_t slice_index0, int sample0, const float* scale,
                                float* logits, void* stream) {
  rcu_unet_outputs out;
  std::memset(&out, 0, sizeof(out));
  out.logits = logits;
  return rcu_unet_forward_ex(net, images, n_slices, n_samples, dropout_mode, det_first, seed, slice_index0, sample0, scale, &out, stream);
}

extern "C" int rcu_unet_forward_ex(rcu_unet* net, const float* images, int64_t n_slices, int n_samples, int dropout_mode,
                                   int det_first, uint64_t seed, int64_t slice_index0, int sample0, const float* scale,
                                   const rcu_unet_outputs* outputs, void* stream) {
  RCU_CHECK_ARG(net != nullptr && images != nullptr && outputs != nullptr && (outputs->logits != nullptr || outputs->logit_diff != nullptr),
                "NULL argument");
  RCU_CHECK_ARG(outputs->logits == nullptr || outputs->logit_diff == nullptr, "logits and logit_diff are alternatives: the head writes one of them");
  float* const logits = outputs->logit_diff != nullptr ? outputs->logit_diff : outputs->logits;
  const int logits_are_diff = outputs->logit_diff != nullptr ? 1 : 0;
  float* const sigma = outputs->sigma;
  float* const features = outputs->features;
  const rcu_postnet* const post = outputs->postnet;
  RCU_CHECK_ARG(sigma == nullptr || net->has_sigma, "sigma output requested from a net without conv_sigma (sigma_out=False)");
  RCU_CHECK_ARG(post == nullptr || outputs->postnet_logits != nullptr, "postnet without postnet_logits");
  RCU_CHECK_ARG(post == nullptr || (post->device == net->device && net->start_filters == kPostC), "postnet does not match this net");
  RCU_CHECK_ARG(net->planned_bytes > 0, "rcu_unet_plan has not been called");
  RCU_CHECK_ARG(!net->ops.empty() && net->arena != nullptr, "no workspace bound (rcu_unet_bind_workspace)");
  RCU_CHECK_ARG(n_slices >= 0 && n_samples >= 1, "bad sizes: n_slices=%lld n_samples=%d", (long long)n_slices, n_samples);
  RCU_CHECK_ARG(dropout_mode >= 0 && dropout_mode <= 2, "dropout_mode must be 0, 1 or 2");
  RCU_CHECK_ARG(dropout_mode != 2 || scale != nullptr || net->total_dropout_channels == 0, "dropout_mode 2 needs a scale table");
  RCU_CHECK_ARG(n_samples <= net->max_images, "n_samples=%d exceeds the planned chunk of %d images", n_samples, net->max_images);
  cudaStream_t st = (cudaStream_t)stream;
  RCU_CUDA(cudaSetDevice(net->device));
  const int H = net->H, W = net->W, sf = net->start_filters;
  const int chunk_slices_max = net->max_images / n_samples;
  const uint32_t thr = dropout_threshold_u32(net->p_drop);
  const float inv_keep = 1.0f / (1.0f - net->p_drop);
  long long launches = 0;

  for (int64_t s0 = 0; s0 < n_slices; s0 += chunk_slices_max) {
    const int cs = (int)((n_slices - s0) < chunk_slices_max ? (n_slices - s0) : chunk_slices_max);
    const int n_img = cs * n_samples;
    net->last_n_img = n_img;
    {
      const long long total = (long long)n_img * net->n_cols;
      long long blocks = (total + 255) / 256;
      if (blocks > (long long)sm_count() * 8) blocks = (long long)sm_count() * 8;
      OpTimer timer(net, (int)net->ops.size(), st);
      coef_kernel<<<(unsigned)blocks, 256, 0, st>>>(net->cols, net->d_coef, n_img, cs, (long long)s0, (long long)n_slices, dropout_mode,
                                                    det_first, (uint32_t)seed, (uint32_t)(seed >> 32), thr, inv_keep,
                                                    (long long)slice_index0, sample0, scale, net->total_dropout_channels);
      RCU_LAUNCH_CHECK();
      ++launches;
    }
    // First-layer dedup: Dropout2d acts between the first conv + bias and its BatchNorm, so the unit's output for a
    // (sample, slice) is the "every channel kept" image of the slice with the dropped channels replaced by constants.  The
    // first conv then writes two images per slice (deterministic / all kept) instead of n_samples, and the pixel-pair kernel
    // that reads them patches the dropped channels into its landed tiles: 21x fewer bytes written and read at 240^2, and
    // bit-identical activations.  Needs the Philox / no-dropout modes (a caller's scale table may hold anything) and the
    // pixel-pair kernel over a 32-channel source as the only consumer.
    int dedup_mode = 0;
    if (net->first_dedup && net->conv_impl == 0 && net->pair_enabled && sf == 32 && dropout_mode != 2 && !net->residual && !net->convs.empty()) {
      const ConvLayer& L1 = net->convs[0];
      const bool reads_first = L1.src.base == net->first_out.base && L1.use_pair && L1.pairp.pair_mode == 1 && (net->halo_mask & 1ull);
      const int variants = (dropout_mode == 0 || !net->unit0_dropout) ? 1 : 2;
      if (reads_first && variants * cs <= n_img)
        dedup_mode = variants == 1 ? 1 : (det_first ? 2 : 3);
    }
    net->last_dedup_mode = dedup_mode;
    bool pool_done = false;   // the previous conv's epilogue already produced the pooled tensor
    for (size_t op_index = 0; op_index < net->ops.size(); ++op_index) {
      const Op& op = net->ops[op_index];
      if (op.kind == OP_POOL && pool_done) { pool_done = false; continue; }
      pool_done = false;
      OpTimer timer(net, (int)op_index, st);
      if (op.kind == OP_FIRST) {
        const int tiles = ((H + kFirstTileH - 1) / kFirstTileH) * ((W + kFirstTileW - 1) / kFirstTileW);
        const size_t smem = ((size_t)net->in_channels * 9 * sf + (size_t)net->in_channels * kFirstInPx) * sizeof(float);
        dim3 grid((unsigned)tiles, (unsigned)cs);
        if (dedup_mode != 0)   // two sample-independent variants per slice (deterministic, every channel kept) instead of n_samples images
          first_conv_kernel<32><<<grid, 256, smem, st>>>(images, net->in_channels, H, W, (long long)s0, cs, dedup_mode == 1 ? 1 : 2, net->d_first_w,
                                                         net->d_e0coef, sf, 0, 1, net->first_out.base, net->first_out.img_stride);
        else if (sf == 32)
          first_conv_kernel<32><<<grid, 256, smem, st>>>(images, net->in_channels, H, W, (long long)s0, cs, n_samples, net->d_first_w,
                                                         net->d_coef, net->n_cols, net->first_coef_off, 0, net->first_out.base,
                                                         net->first_out.img_stride);
        else
          first_conv_kernel<64><<<grid, 256, smem, st>>>(images, net->in_channels, H, W, (long long)s0, cs, n_samples, net->d_first_w,
                                                         net->d_coef, net->n_cols, net->first_coef_off, 0, net->first_out.base,
                                                         net->first_out.img_stride);
        RCU_LAUNCH_CHECK();
        ++launches;
      } else if (op.kind == OP_RES_FIRST) {
        const long long total = (long long)cs * H * W * (sf / 8);
        long long blocks = (total + 255) / 256;
        if (blocks > (long long)sm_count() * 16) blocks = (long long)sm_count() * 16;
        first_residual_kernel<<<(unsigned)blocks, 256, 0, st>>>(images, net->in_channels, H * W, (long long)s0, cs, n_samples, net->d_res0_w,
                                                                net->d_res0_b, sf, net->res0_dst.base, net->res0_dst.c_total,
                                                                net->res0_dst.img_stride);
        RCU_LAUNCH_CHECK();
        ++launches;
      } else if (op.kind == OP_POOL) {
        const long long total = (long long)n_img * op.h * op.w * (op.c / 8);
        long long blocks = (total + 255) / 256;
        if (blocks > (long long)sm_count() * 16) blocks = (long long)sm_count() * 16;
        maxpool2_kernel<<<(unsigned)blocks, 256, 0, st>>>(op.in.base, op.in.c_total, op.in.img_stride, op.out.base, op.out.img_stride,
                                                          n_img, op.in_h, op.in_w, op.c);
        RCU_LAUNCH_CHECK();
        ++launches;
      } else {
        const ConvLayer& L = net->convs[op.conv];
        float* const head_out = L.out_slot ? sigma : logits;
        const int head_diff = L.out_slot ? 0 : logits_are_diff;
        if (L.out_slot && sigma == nullptr) continue;   // nobody asked for the sigma branch
        if (net->conv_impl == 0 && L.use_pair && net->pair_enabled && ((net->halo_mask >> op.conv) & 1ull)) {
          int rc = run_conv_halo_pair(net, L, n_img, cs, (long long)s0, (long long)n_slices, head_out, head_diff, op.conv == 0 ? dedup_mode : 0, st, &launches);
          if (rc) return rc;
          pool_done = L.pool_dst.base != nullptr;
          continue;
        }
        if (net->conv_impl == 0 && L.halo.ok && ((net->halo_mask >> op.conv) & 1ull)) {
          int rc = run_conv_halo(net, L, n_img, cs, (long long)s0, (long long)n_slices, head_out, head_diff, st, &launches);
          if (rc) return rc;
          pool_done = L.pool_dst.base != nullptr;
          continue;
        }
        if (net->conv_impl == 0 && L.wide.ok && !L.head && ((net->halo_mask >> op.conv) & 1ull)) {
          WideParams wp;
          std::memset(&wp, 0, sizeof(wp));
          wp.n_img = n_img; wp.in_h = L.in_h; wp.in_w = L.in_w;
          wp.tiles_x = (L.in_w + kWideTile - 1) / kWideTile;
          wp.tiles_y = (L.in_h + kWideTile - 1) / kWideTile;
          wp.n_chunks = L.wide.n_chunks; wp.n_ntiles = L.wide.n_ntiles; wp.n_phases = L.wide.n_phases;
          wp.w_image = L.wide.d_wimg;
          wp.out_mul = L.out_mul;
          wp.out_h = L.in_h * L.out_mul; wp.out_w = L.in_w * L.out_mul;
          wp.out_c = L.dst.c_total; wp.out_img_stride = L.dst.img_stride; wp.out = L.dst.base;
          wp.coef = net->d_coef; wp.coef_stride = net->n_cols; wp.coef_off = L.coef_off;
          wp.relu = L.relu;
          {
            static const bool allow = [] { const char* e = std::getenv("RCU_WIDE_TMA_STORE"); return !(e && e[0] == '0'); }();
            wp.tma_store = allow ? 1 : 0;
          }
          HaloOutMaps maps;
          for (int i = 0; i < L.wide.n_phases; ++i) maps.m[i] = L.map_out[i];
          static const bool fuse_pool = [] { const char* e = std::getenv("RCU_WIDE_POOL"); return !(e && e[0] == '0'); }();
          if (fuse_pool && L.wide.n_phases == 1 && L.pool_dst.base != nullptr) {
            wp.pool_out = L.pool_dst.base; wp.pool_img_stride = L.pool_dst.img_stride; wp.pool_c = L.pool_dst.c_total;
          }
          int rc = L.wide.n_phases == 4 ? launch_conv_wide<4>(L, wp, maps, st) : launch_conv_wide<9>(L, wp, maps, st);
          if (rc) return rc;
          ++launches;
          pool_done = wp.pool_out != nullptr;
          continue;
        }
        ConvParams prm;
        std::memset(&prm, 0, sizeof(prm));
        prm.n_img = n_img;
        prm.in_h = L.in_h; prm.in_w = L.in_w;
        prm.tiles_x = (L.in_w + kTileW - 1) / kTileW;
        prm.tiles_y = (L.in_h + kTileH - 1) / kTileH;
        prm.n_tiles_n = L.c_out / L.block_n;
        prm.kc0 = L.c0 / L.kc; prm.kc1 = L.c1 / L.kc;
        prm.n_taps = L.n_taps; prm.n_phases = L.n_phases;
        std::memcpy(prm.dy, L.dy, sizeof(prm.dy));
        std::memcpy(prm.dx, L.dx, sizeof(prm.dx));
        prm.out_mul = L.out_mul;
        prm.out_h = L.in_h * L.out_mul; prm.out_w = L.in_w * L.out_mul;
        prm.out_c = L.dst.c_total;
        prm.out_img_stride = L.dst.img_stride;
        prm.out = L.dst.base;
        prm.coef = net->d_coef; prm.coef_stride = net->n_cols; prm.coef_off = L.coef_off;
        prm.relu = L.relu;
        prm.accumulate = L.accumulate ? 1 : 0;
        prm.head = nullptr; prm.logits = head_out; prm.head_diff = head_diff;
        prm.chunk_slices = cs; prm.slice0 = s0; prm.n_slices_total = n_slices;
        if (net->conv_impl != 1) {
          if (L.head) prm.head = L.d_head;
          int rc = dispatch_conv_tc(L, prm, st);
          if (rc) return rc;
          ++launches;
        } else {
          const long long total = (long long)n_img * L.n_phases * L.in_h * L.in_w * L.c_out;
          long long blocks = (total + 255) / 256;
          if (blocks > (long long)sm_count() * 32) blocks = (long long)sm_count() * 32;
          conv_check_kernel<<<(unsigned)blocks, 256, 0, st>>>(L.src.base, L.c0, L.src.c_total, L.src.img_stride, L.d_weights, L.c_out, prm);
          RCU_LAUNCH_CHECK();
          ++launches;
          if (L.head) {
            const long long px = (long long)n_img * H * W;
            long long hb = (px + 255) / 256;
            if (hb > (long long)sm_count() * 32) hb = (long long)sm_count() * 32;
            head_check_kernel<<<(unsigned)hb, 256, 0, st>>>(L.dst.base, L.dst.img_stride, L.d_head, head_out, n_img, H, W,
                                                           sf, cs, (long long)s0, (long long)n_slices, head_diff);
            RCU_LAUNCH_CHECK();
            ++launches;
          }
        }
      }
    }
    if (features != nullptr) {
      // `UNet.features` (unet.py:178-179): the decoder output that feeds conv_cls, as float32 NCHW per (sample, slice)
      const Act& f = net->features;
      const int runs = (H * W + kFeatRun - 1) / kFeatRun;
      dim3 grid((unsigned)runs, (unsigned)n_img);
      OpTimer timer(net, (int)net->ops.size(), st);
      if (sf == 32)
        features_export_kernel<32><<<grid, 256, 0, st>>>(f.base, f.c_total, f.img_stride, features, H * W, cs, (long long)s0, (long long)n_slices);
      else
        features_export_kernel<64><<<grid, 256, 0, st>>>(f.base, f.c_total, f.img_stride, features, H * W, cs, (long long)s0, (long long)n_slices);
      RCU_LAUNCH_CHECK();
      ++launches;
    }
    if (post != nullptr) {
      const Act& f = net->features;
      OpTimer timer(net, (int)net->ops.size(), st);
      int rc = launch_postnet(post, true, f.base, f.c_total, f.img_stride, H * W, n_img, cs, (long long)s0, (long long)n_slices,
                              outputs->postnet_logits, st);
      if (rc) return rc;
      ++launches;
    }
  }
  net->last_launches = launches;
  return RCU_OK;
}

extern "C" int rcu_postnet_create(const rcu_conv_unit* units, int n_units, const rcu_conv_unit* head, float bn_eps, int device, rcu_postnet** out) {
  RCU_CHECK_ARG(out != nullptr && head != nullptr && (units != nullptr || n_units == 0), "NULL argument");
  *out = nullptr;
  int rc = rcu_device_check(device);
  if (rc) return rc;
  if (n_units < 1 || n_units > kPostMaxConvs) { set_error("PostNet nb_convs=%d: supported 1..%d", n_units, kPostMaxConvs); return RCU_ENOTSUP; }
  rcu_postnet* post = new rcu_postnet();
  post->device = device;
  post->n_convs = n_units;
  std::memset(&post->w, 0, sizeof(post->w));
  for (int l = 0; l < n_units; ++l) {
    const rcu_conv_unit& u = units[l];
    if (u.c_in != kPostC || u.c_out != kPostC) { set_error("PostNet unit %d: %d -> %d channels, supported 32 -> 32", l, u.c_in, u.c_out); delete post; return RCU_ENOTSUP; }
    if (!(u.weight && u.bias && u.bn_weight && u.bn_bias && u.bn_mean && u.bn_var)) { set_error("PostNet unit %d: NULL weight/bn pointer", l); delete post; return RCU_EINVAL; }
    for (int co = 0; co < kPostC; ++co) {
      // y = relu(a * (W x + b - mean) + beta): fold a into the row and the bias
      const float a = u.bn_weight[co] / std::sqrt(u.bn_var[co] + bn_eps);
      for (int ci = 0; ci < kPostC; ++ci) post->w.w[l][co][ci] = a * u.weight[co * kPostC + ci];
      post->w.b[l][co] = a * (u.bias[co] - u.bn_mean[co]) + u.bn_bias[co];
    }
  }
  if (!(head->weight && head->bias) || head->c_in != kPostC || head->c_out != 2) { set_error("PostNet head must be a 1x1 conv 32 -> 2"); delete post; return RCU_EINVAL; }
  for (int k = 0; k < 2; ++k) {
    for (int c = 0; c < kPostC; ++c) post->w.hw[k][c] = head->weight[k * kPostC + c];
    post->w.hb[k] = head->bias[k];
  }
  *out = post;
  return RCU_OK;
}

extern "C" void rcu_postnet_destroy(rcu_postnet* post) { delete post; }

extern "C" int rcu_postnet_forward(const rcu_postnet* post, const float* features, int64_t n_images, int64_t hw, float* logits, void* stream) {
  RCU_CHECK_ARG(post != nullptr && features != nullptr && logits != nullptr, "NULL argument");
  RCU_CHECK_ARG(n_images >= 0 && hw >= 0 && hw < (1ll << 31) && n_images < (1ll << 31), "bad sizes");
  if (n_images == 0 || hw == 0) return RCU_OK;
  RCU_CUDA(cudaSetDevice(post->device));
  return launch_postnet(post, false, features, 0, 0, (int)hw, (int)n_images, (int)n_images, 0, n_images, logits, (cudaStream_t)stream);
}

extern "C" int rcu_unet_debug_activation(rcu_unet* net, int index, float* out, size_t out_elems, void* stream) {
  RCU_CHECK_ARG(net != nullptr && out != nullptr, "NULL argument");
  RCU_CHECK_ARG(index >= 0 && index < (int)net->ops.size(), "activation index %d out of range [0, %d)", index, (int)net->ops.size());
  const Op& op = net->ops[index];
  if (op.kind == OP_CONV && net->convs[op.conv].head && net->conv_impl != 1) {
    set_error("activation %d is fused away (conv_cls.0 feeds the head in registers on the tcgen05 path)", index);
    return RCU_ENOTSUP;
  }
  if (op.kind == OP_FIRST && net->last_dedup_mode != 0) {
    set_error("activation %d is stored once per slice (first-layer dedup): rcu_unet_set_first_layer_dedup(net, 0) materialises it per sample", index);
    return RCU_ENOTSUP;
  }
  const long long n = (long long)net->last_n_img * op.h * op.w * op.c;
  RCU_CHECK_ARG((long long)out_elems >= n, "output holds %zu elements, activation has %lld", out_elems, n);
  if (n == 0) return RCU_OK;
  long long blocks = (n + 255) / 256;
  if (blocks > 4096) blocks = 4096;
  bf16_to_f32_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(op.out.base, op.out.c_total, op.out.img_stride,
                                                                        (long long)op.h * op.w, op.c, out, n);
  RCU_LAUNCH_CHECK();
  return RCU_OK;
}

extern "C" int rcu_unet_enable_timing(rcu_unet* net, int enable) {
  RCU_CHECK_ARG(net != nullptr, "NULL handle");
  net->timing = enable != 0;
  net->ev_used = 0;
  net->ev_op.clear();
  return RCU_OK;
}

extern "C" int rcu_unet_num_ops(const rcu_unet* net) { return net ? (int)net->ops.size() + 1 : 0; }

extern "C" int rcu_unet_op_info(const rcu_unet* net, int op, int* kind, int64_t* macs_per_image, int* c_in, int* c_out, int* h, int* w) {
  RCU_CHECK_ARG(net != nullptr, "NULL handle");
  RCU_CHECK_ARG(op >= 0 && op <= (int)net->ops.size(), "op index %d out of range", op);
  int k = 3, ci = 0, co = 0, hh = 0, ww = 0;
  int64_t macs = 0;
  if (op < (int)net->ops.size()) {
    const Op& o = net->ops[op];
    hh = o.h; ww = o.w; co = o.c;
    if (o.kind == OP_FIRST) {
      k = 0; ci = net->in_channels;
      macs = (int64_t)net->H * net->W * 9 * net->in_channels * net->start_filters;
    } else if (o.kind == OP_POOL) {
      k = 2; ci = o.c;
    } else if (o.kind == OP_RES_FIRST) {
      k = 0; ci = net->in_channels;
      macs = (int64_t)o.h * o.w * ci * o.c;
    } else {
      const ConvLayer& L = net->convs[o.conv];
      k = 1; ci = L.c0 + L.c1;
      // algorithmic MACs of the reference layer: a 3x3 conv at the OUTPUT resolution (for the up-path conv that is
      // the conv after nearest-x2, common/model/unet.py:105; 1x1 for a residual branch), plus the fused 1x1 head
      macs = (int64_t)o.h * o.w * (L.accumulate ? 1 : 9) * ci * L.c_out + (L.head ? (int64_t)o.h * o.w * L.c_out * 2 : 0);
    }
  }
  if (kind) *kind = k;
  if (macs_per_image) *macs_per_image = macs;
  if (c_in) *c_in = ci;
  if (c_out) *c_out = co;
  if (h) *h = hh;
  if (w) *w = ww;
  return RCU_OK;
}

extern "C" int rcu_unet_op_executed_macs(const rcu_unet* net, int op, int64_t* macs_per_image) {
  RCU_CHECK_ARG(net != nullptr && macs_per_image != nullptr, "NULL argument");
  RCU_CHECK_ARG(op >= 0 && op <= (int)net->ops.size(), "op index %d out of range", op);
  int64_t macs = 0;
  if (op < (int)net->ops.size()) {
    const Op& o = net->ops[op];
    if (o.kind == OP_FIRST) {
      macs = (int64_t)net->H * net->W * 9 * net->in_channels * net->start_filters;
    } else if (o.kind == OP_RES_FIRST) {
      macs = (int64_t)net->H * net->W * net->in_channels * net->start_filters;
    } else if (o.kind == OP_CONV) {
      const ConvLayer& L = net->convs[o.conv];
      // the up-path convs run as four 2x2-tap phase convolutions on the low-resolution input: 4 taps per OUTPUT pixel
      const int taps = L.accumulate ? 1 : (L.n_phases == 4 ? 4 : 9);
      macs = (int64_t)o.h * o.w * taps * (L.c0 + L.c1) * L.c_out + (L.head ? (int64_t)o.h * o.w * L.c_out * 2 : 0);
    }
  }
  *macs_per_image = macs;
  return RCU_OK;
}

extern "C" int rcu_unet_read_timing(rcu_unet* net, float* ms, int64_t* launches, int n_ops) {
  RCU_CHECK_ARG(net != nullptr && ms != nullptr && launches != nullptr, "NULL argument");
  RCU_CHECK_ARG(n_ops == (int)net->ops.size() + 1, "expected %d ops", (int)net->ops.size() + 1);
  for (int i = 0; i < n_ops; ++i) { ms[i] = 0.f; launches[i] = 0; }
  for (size_t i = 0; i < net->ev_used; ++i) {
    RCU_CUDA(cudaEventSynchronize(net->ev_pool[i].second));
    float t = 0.f;
    RCU_CUDA(cudaEventElapsedTime(&t, net->ev_pool[i].first, net->ev_pool[i].second));
    ms[net->ev_op[i]] += t;
    launches[net->ev_op[i]] += 1;
  }
  net->ev_used = 0;
  net->ev_op.clear();
  return RCU_OK;
}
