// Non-GEMM kernels of the U-Net forward: first conv (tiny K, fp32 NCHW input), 2x2 max-pool, the per-(image,
// channel) epilogue-coefficient table (BN fold x Dropout2d keep-scale, Philox generated), and a plain CUDA-core
// convolution over the same bf16 data that exists only as an on-device cross-check of the tcgen05 path.
#pragma once
#include <cuda_bf16.h>

#include "common.cuh"
#include "conv_tc.cuh"
#include "philox.cuh"

namespace rcu {

// ---------------------------------------------------------------------------------------------------------
// Epilogue coefficient table.  For unit u, channel c, image i (= sample t, slice g):
//     s = 1 (Dropout2d in eval mode)  or  keep(t, g, site(u), c) / (1 - p)        (common/model/unet.py:14-15)
//     y = relu( (conv + bias) * s * a + d ),  a = gamma / sqrt(var + eps),  d = beta - a * mean   (unet.py:16-19)
//       = relu( conv * (a*s) + (bias*a*s + d) )            -> coef = (a*s, bias*a*s + d)
// Units without BN (up-path `upconv`, unet.py:105): coef = (1, bias).
// ---------------------------------------------------------------------------------------------------------
struct CoefColumns {            // device arrays, one entry per column of the table
  const float* fold_a;          // a        (1 for no-BN units)
  const float* fold_ba;         // bias * a (bias for no-BN units)
  const float* fold_d;          // d        (0 for no-BN units)
  const int* site;              // dropout site index or -1
  const int* ch_in_site;        // channel index inside the site
  const int* scale_col;         // column inside a caller-supplied scale row (mode 2)
  int n_cols;
};

__global__ void coef_kernel(CoefColumns cols, float2* __restrict__ coef, int n_img, int chunk_slices, long long slice0,
                            long long n_slices_total, int dropout_mode, int det_first, uint32_t seed_lo, uint32_t seed_hi,
                            uint32_t thr, float inv_keep, long long slice_index0, int sample0, const float* __restrict__ scale,
                            int scale_cols) {
  const long long total = (long long)n_img * cols.n_cols;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int col = (int)(i % cols.n_cols);
    const int img = (int)(i / cols.n_cols);
    const int t = img / chunk_slices, sl = img - t * chunk_slices;
    const long long g = slice0 + sl;  // slice index inside this forward call
    float s = 1.0f;
    const int site = cols.site[col];
    const bool stochastic = dropout_mode != 0 && site >= 0 && !(det_first && t == 0);
    if (stochastic) {
      const int ts = t - (det_first ? 1 : 0);
      if (dropout_mode == 1)
        s = dropout_scale(seed_lo, seed_hi, thr, inv_keep, (uint32_t)site, (uint32_t)(slice_index0 + g), (uint32_t)(sample0 + ts),
                          (uint32_t)cols.ch_in_site[col]);
      else
        s = scale[((long long)ts * n_slices_total + g) * scale_cols + cols.scale_col[col]];
    }
    const float as = cols.fold_a[col] * s;
    coef[i] = make_float2(as, fmaf(cols.fold_ba[col], s, cols.fold_d[col]));
  }
}

// ---------------------------------------------------------------------------------------------------------
// First conv: fp32 NCHW slices (C_in = 3 or 4) -> bf16 NHWC, 32*k output channels.  K = 9*C_in <= 36 is too
// thin for a tensor-core tile; the convolution itself does not depend on the MC sample (dropout acts after it),
// so each thread convolves its pixels ONCE and then only re-applies the per-sample coefficients: the layer is
// store-bound (64 B per pixel-sample).  Thread t owns 8 output channels (group t % CG) of CG pixels, so the
// lanes of a warp write 32 consecutive 16-byte pieces (8 x-adjacent pixels x 64 B): fully coalesced stores, and
// only 8 coefficient pairs per thread and sample.
// ---------------------------------------------------------------------------------------------------------
// 256 pixels per block of 256 threads as a wide, flat tile: every sample's store is kFirstTileH runs of
// kFirstTileW * 64 bytes (4 KB with 32 output channels), long enough to stream into HBM pages
#ifndef RCU_FIRST_TILE_W
#define RCU_FIRST_TILE_W 64
#endif
constexpr int kFirstTileW = RCU_FIRST_TILE_W, kFirstTileH = 256 / kFirstTileW;
constexpr int kFirstInW = kFirstTileW + 2, kFirstInH = kFirstTileH + 2, kFirstInPx = kFirstInW * kFirstInH;

template <int C_OUT>
__global__ void __launch_bounds__(256)
first_conv_kernel(const float* __restrict__ images, int c_in, int h, int w, long long slice0, int chunk_slices, int n_samples,
                  const float* __restrict__ weight /* [c_in*9][C_OUT] */, const float2* __restrict__ coef, long long coef_stride,
                  int coef_off, int coef_row_per_sample, __nv_bfloat16* __restrict__ out, long long out_img_stride) {
  constexpr int CG = C_OUT / 8;      // channel groups = pixels per thread
  constexpr int PG = 256 / CG;       // pixel groups
  extern __shared__ float s_first[];
  float* s_w = s_first;                                  // [c_in*9][C_OUT]
  float* s_in = s_first + c_in * 9 * C_OUT;              // [c_in][kFirstInH][kFirstInW]
  const int tiles_x = (w + kFirstTileW - 1) / kFirstTileW;
  const int tile = blockIdx.x;
  const int sl = blockIdx.y;
  const int ty0 = (tile / tiles_x) * kFirstTileH, tx0 = (tile % tiles_x) * kFirstTileW;
  const float* img = images + (slice0 + sl) * (long long)c_in * h * w;
  for (int i = threadIdx.x; i < c_in * 9 * C_OUT; i += 256) s_w[i] = weight[i];
  for (int i = threadIdx.x; i < c_in * kFirstInPx; i += 256) {
    const int c = i / kFirstInPx, r = i - c * kFirstInPx;
    const int yy = ty0 + r / kFirstInW - 1, xx = tx0 + r % kFirstInW - 1;
    s_in[i] = (yy >= 0 && yy < h && xx >= 0 && xx < w) ? img[((long long)c * h + yy) * w + xx] : 0.0f;
  }
  __syncthreads();
  const int cg = threadIdx.x % CG, pg = threadIdx.x / CG;
  float acc[CG][8];
#pragma unroll
  for (int j = 0; j < CG; ++j)
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[j][c] = 0.0f;
  for (int ci = 0; ci < c_in; ++ci)
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      const float4 w0 = *reinterpret_cast<const float4*>(s_w + (ci * 9 + k) * C_OUT + cg * 8);
      const float4 w1 = *reinterpret_cast<const float4*>(s_w + (ci * 9 + k) * C_OUT + cg * 8 + 4);
#pragma unroll
      for (int j = 0; j < CG; ++j) {
        const int p = pg + PG * j, ly = p / kFirstTileW, lx = p % kFirstTileW;
        const float v = s_in[ci * kFirstInPx + (ly + k / 3) * kFirstInW + lx + k % 3];
        acc[j][0] = fmaf(v, w0.x, acc[j][0]); acc[j][1] = fmaf(v, w0.y, acc[j][1]);
        acc[j][2] = fmaf(v, w0.z, acc[j][2]); acc[j][3] = fmaf(v, w0.w, acc[j][3]);
        acc[j][4] = fmaf(v, w1.x, acc[j][4]); acc[j][5] = fmaf(v, w1.y, acc[j][5]);
        acc[j][6] = fmaf(v, w1.z, acc[j][6]); acc[j][7] = fmaf(v, w1.w, acc[j][7]);
      }
    }
  for (int t = 0; t < n_samples; ++t) {
    const int im = t * chunk_slices + sl;
    const float2* cf = coef + (long long)(coef_row_per_sample ? t : im) * coef_stride + coef_off + cg * 8;   // dedup: one row per variant
    float2 c8[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) c8[c] = __ldg(cf + c);
#pragma unroll
    for (int j = 0; j < CG; ++j) {
      const int p = pg + PG * j, y = ty0 + p / kFirstTileW, x = tx0 + p % kFirstTileW;
      if (y >= h || x >= w) continue;
      uint32_t pk[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float a0 = fmaxf(fmaf(acc[j][2 * c], c8[2 * c].x, c8[2 * c].y), 0.0f);
        const float a1 = fmaxf(fmaf(acc[j][2 * c + 1], c8[2 * c + 1].x, c8[2 * c + 1].y), 0.0f);
        __nv_bfloat162 b = __floats2bfloat162_rn(a0, a1);
        pk[c] = *reinterpret_cast<uint32_t*>(&b);
      }
      *reinterpret_cast<uint4*>(out + (long long)im * out_img_stride + ((long long)y * w + x) * C_OUT + cg * 8) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// 2x2 max-pool, bf16 NHWC (nn.MaxPool2d(2), common/model/unet.py:90).  One thread = 8 channels of one output pixel.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 bf16x8_max(uint4 a, uint4 b) {
  uint4 r;
  const __nv_bfloat162* pa = reinterpret_cast<const __nv_bfloat162*>(&a);
  const __nv_bfloat162* pb = reinterpret_cast<const __nv_bfloat162*>(&b);
  __nv_bfloat162* pr = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) pr[i] = __hmax2(pa[i], pb[i]);
  return r;
}

// Residual branch of the FIRST block of a residual=True net (ConvResidualBlock.residual, common/model/unet.py:57-59): a 1x1
// convolution of the float32 input slices (c_in <= 8 channels) added onto the block output already in `out`.  The
// branch does not depend on the MC sample, so it is computed once per slice pixel and added to every sample's image.
// One thread = one pixel x 8 output channels.
__global__ void __launch_bounds__(256)
first_residual_kernel(const float* __restrict__ images, int c_in, int hw, long long slice0, int chunk_slices, int n_samples,
                      const float* __restrict__ weight /* [c_out][c_in] */, const float* __restrict__ bias, int c_out,
                      __nv_bfloat16* __restrict__ out, int out_c, long long out_img_stride) {
  const int groups = c_out / 8;
  const long long total = (long long)chunk_slices * hw * groups;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const int g = (int)(i % groups);
    const long long px_all = i / groups;
    const int sl = (int)(px_all / hw);
    const int px = (int)(px_all - (long long)sl * hw);
    const float* img = images + (slice0 + sl) * (long long)c_in * hw + px;
    float r[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) r[c] = __ldg(bias + g * 8 + c);
    for (int ci = 0; ci < c_in; ++ci) {
      const float v = __ldg(img + (long long)ci * hw);
#pragma unroll
      for (int c = 0; c < 8; ++c) r[c] = fmaf(v, __ldg(weight + (g * 8 + c) * c_in + ci), r[c]);
    }
    for (int t = 0; t < n_samples; ++t) {
      uint4* dst = reinterpret_cast<uint4*>(out + (long long)(t * chunk_slices + sl) * out_img_stride + (long long)px * out_c + g * 8);
      uint4 q = *dst;
      uint32_t w4[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float2 old = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w4[c]));
        __nv_bfloat162 b = __floats2bfloat162_rn(old.x + r[2 * c], old.y + r[2 * c + 1]);
        w4[c] = *reinterpret_cast<uint32_t*>(&b);
      }
      *dst = make_uint4(w4[0], w4[1], w4[2], w4[3]);
    }
  }
}


__global__ void __launch_bounds__(256)
maxpool2_kernel(const __nv_bfloat16* __restrict__ in, int in_px_stride, long long in_img_stride, __nv_bfloat16* __restrict__ out,
                long long out_img_stride, long long n_img, int h, int w, int c) {
  // `in` may be a channel slice of a wider tensor (the skip half of a concat buffer): pixel stride in_px_stride
  const int oh = h >> 1, ow = w >> 1, c8 = c >> 3;
  const long long total = n_img * oh * ow * c8;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const int cc = (int)(i % c8);
    long long r = i / c8;
    const int ox = (int)(r % ow); r /= ow;
    const int oy = (int)(r % oh);
    const long long im = r / oh;
    const __nv_bfloat16* p0 = in + im * in_img_stride + ((long long)(2 * oy) * w + 2 * ox) * in_px_stride + cc * 8;
    const long long row = (long long)w * in_px_stride;
    const uint4 m = bf16x8_max(bf16x8_max(__ldg(reinterpret_cast<const uint4*>(p0)), __ldg(reinterpret_cast<const uint4*>(p0 + in_px_stride))),
                               bf16x8_max(__ldg(reinterpret_cast<const uint4*>(p0 + row)), __ldg(reinterpret_cast<const uint4*>(p0 + row + in_px_stride))));
    *reinterpret_cast<uint4*>(out + im * out_img_stride + ((long long)oy * ow + ox) * c + cc * 8) = m;
  }
}

// ---------------------------------------------------------------------------------------------------------
// Cross-check convolution (CUDA cores, fp32 accumulate) with exactly the tcgen05 kernel's contract.
// One thread = one output pixel x one output channel.  Debug only: rcu_unet_set_conv_impl(net, 1).
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
conv_check_kernel(const __nv_bfloat16* __restrict__ src, int c_in, int src_px_stride, long long src_img_stride,
                  const __nv_bfloat16* __restrict__ weights, int c_out, const ConvParams prm) {
  const long long total = (long long)prm.n_img * prm.n_phases * prm.in_h * prm.in_w * c_out;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    long long r = i;
    const int co = (int)(r % c_out); r /= c_out;
    const int x = (int)(r % prm.in_w); r /= prm.in_w;
    const int y = (int)(r % prm.in_h); r /= prm.in_h;
    const int ph = (int)(r % prm.n_phases);
    const int img = (int)(r / prm.n_phases);
    float acc = 0.0f;
    for (int tap = 0; tap < prm.n_taps; ++tap) {
      const int yy = y + prm.dy[ph][tap], xx = x + prm.dx[ph][tap];
      if (yy < 0 || yy >= prm.in_h || xx < 0 || xx >= prm.in_w) continue;
      const __nv_bfloat16* wrow = weights + ((long long)(ph * prm.n_taps + tap) * c_out + co) * c_in;
      const __nv_bfloat16* a0 = src + (long long)img * src_img_stride + ((long long)yy * prm.in_w + xx) * src_px_stride;
      for (int k = 0; k < c_in; ++k) acc = fmaf(__bfloat162float(a0[k]), __bfloat162float(wrow[k]), acc);
    }
    const float2 cf = prm.coef[(long long)img * prm.coef_stride + prm.coef_off + co];
    float v = fmaf(acc, cf.x, cf.y);
    const int oy = prm.out_mul * y + (ph >> 1), ox = prm.out_mul * x + (ph & 1);
    __nv_bfloat16* o = prm.out + (long long)img * prm.out_img_stride + ((long long)oy * prm.out_w + ox) * prm.out_c + co;
    if (prm.accumulate) v += __bfloat162float(*o);
    if (prm.relu) v = fmaxf(v, 0.0f);
    *o = __float2bfloat16_rn(v);
  }
}

// 1x1 head for the cross-check path: bf16 features [img][h][w][32] -> logits (the tcgen05 path fuses this).
__global__ void __launch_bounds__(256)
head_check_kernel(const __nv_bfloat16* __restrict__ feat, long long feat_img_stride, const float* __restrict__ head,
                  float* __restrict__ logits, int n_img, int h, int w, int c, int chunk_slices, long long slice0, long long n_slices_total,
                  int head_diff) {
  const long long total = (long long)n_img * h * w;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const long long img = i / ((long long)h * w), px = i - img * h * w;
    float l0 = head[2 * c], l1 = head[2 * c + 1];
    for (int k = 0; k < c; ++k) {
      const float a = __bfloat162float(feat[img * feat_img_stride + px * c + k]);
      l0 = fmaf(a, head[k], l0);
      l1 = fmaf(a, head[c + k], l1);
    }
    const int t = (int)(img / chunk_slices), sl = (int)(img - (long long)t * chunk_slices);
    const long long gimg = (long long)t * n_slices_total + slice0 + sl;
    if (head_diff) logits[gimg * h * w + px] = l0 - l1;
    else reinterpret_cast<float2*>(logits)[gimg * h * w + px] = make_float2(l0, l1);
  }
}

// `UNet.features` export (common/model/unet.py:178-179): (strided) NHWC bf16 [image in chunk][hw][C] -> float32 NCHW
// [sample][slice][C][hw].  A block transposes kFeatRun pixels x C channels through shared memory: 16-byte coalesced
// loads, 256-byte coalesced store runs per channel.
constexpr int kFeatRun = 64;

template <int C>
__global__ void __launch_bounds__(256)
features_export_kernel(const __nv_bfloat16* __restrict__ feat, int px_stride, long long img_stride, float* __restrict__ out, int hw,
                       int chunk_slices, long long slice0, long long n_slices_total) {
  __shared__ float tile[C][kFeatRun + 1];
  constexpr int V = C / 8;
  const int img = blockIdx.y, px0 = blockIdx.x * kFeatRun;
  for (int i = threadIdx.x; i < kFeatRun * V; i += 256) {
    const int px = i / V, v = i % V;
    if (px0 + px < hw) {
      const uint4 q = __ldg(reinterpret_cast<const uint4*>(feat + (long long)img * img_stride + (long long)(px0 + px) * px_stride + v * 8));
      const __nv_bfloat162* b2 = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __bfloat1622float2(b2[j]);
        tile[v * 8 + 2 * j][px] = f.x;
        tile[v * 8 + 2 * j + 1][px] = f.y;
      }
    }
  }
  __syncthreads();
  const int t = img / chunk_slices, sl = img - t * chunk_slices;
  float* o = out + ((long long)t * n_slices_total + slice0 + sl) * (long long)C * hw;
  for (int i = threadIdx.x; i < C * kFeatRun; i += 256) {
    const int c = i / kFeatRun, px = i % kFeatRun;
    if (px0 + px < hw) o[(long long)c * hw + px0 + px] = tile[c][px];
  }
}

// (strided) NHWC bf16 -> dense fp32 copy used by rcu_unet_debug_activation.
__global__ void bf16_to_f32_kernel(const __nv_bfloat16* __restrict__ in, int px_stride, long long img_stride, long long hw, int c,
                                   float* __restrict__ out, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int ch = (int)(i % c);
    const long long px = (i / c) % hw, img = i / c / hw;
    out[i] = __bfloat162float(in[img * img_stride + px * px_stride + ch]);
  }
}

}  // namespace rcu
