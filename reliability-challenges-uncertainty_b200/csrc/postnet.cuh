// PostNet (common/model/postnet.py:6-17): nb_convs x [1x1 conv -> (Dropout2d, eval) -> BN(eval) -> ReLU] + a 1x1 logits
// conv, applied per pixel to the U-Net's `features` (bin-dl/brats_test_auxiliary_feat.py:61-80).  A per-pixel MLP of
// 32x32 layers: 3136 FMA per pixel against 64 B of input, i.e. FMA-bound on the CUDA cores (K = 32 is one quarter of a
// tcgen05 tile's depth and the chain would have to bounce through shared memory between layers).  One thread owns
// one pixel; the folded weights travel as a __grid_constant__ kernel parameter, so every FFMA takes its weight
// straight from the constant bank (warp-uniform, no shared-memory or register staging).
#pragma once
#include <cuda_bf16.h>

#include "common.cuh"

namespace rcu {

constexpr int kPostC = 32;
constexpr int kPostMaxConvs = 4;

struct PostNetWeights {
  float w[kPostMaxConvs][kPostC][kPostC];   // [layer][c_out][c_in], BN scale folded in
  float b[kPostMaxConvs][kPostC];           // a * bias + d
  float hw[2][kPostC];
  float hb[2];
};

// SRC_BF16: features are the bf16 NHWC tensor the U-Net forward left in its workspace (image order [sample][slice in
// chunk], pixel stride `px_stride` elements); otherwise float32 NCHW [image][32][hw] as `UNet.features` holds them.
template <int NC, bool SRC_BF16>
__global__ void __launch_bounds__(128, SRC_BF16 ? 4 : 5)   // measured: 128 registers for the bf16 source, ~100 for the NCHW one (uncapped: 186 / 255 and up to 2x slower)
postnet_kernel(const __grid_constant__ PostNetWeights wt, const void* __restrict__ src, int px_stride, long long img_stride, int hw,
               int n_img, int chunk_slices, long long slice0, long long n_slices_total, float* __restrict__ logits) {
  const long long total = (long long)n_img * hw;
  for (long long i = (long long)blockIdx.x * 128 + threadIdx.x; i < total; i += (long long)gridDim.x * 128) {
    const int img = (int)(i / hw), px = (int)(i - (long long)img * hw);
    float x[kPostC];
    if (SRC_BF16) {
      const uint4* p = reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(src) + (long long)img * img_stride + (long long)px * px_stride);
#pragma unroll
      for (int v = 0; v < kPostC / 8; ++v) {
        const uint4 q = __ldg(p + v);
        const __nv_bfloat162* b2 = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = __bfloat1622float2(b2[j]);
          x[v * 8 + 2 * j] = f.x;
          x[v * 8 + 2 * j + 1] = f.y;
        }
      }
    } else {
      const float* p = static_cast<const float*>(src) + (long long)img * kPostC * hw + px;
#pragma unroll
      for (int c = 0; c < kPostC; ++c) x[c] = __ldg(p + (long long)c * hw);
    }
#pragma unroll
    for (int l = 0; l < NC; ++l) {
      float y[kPostC];
#pragma unroll
      for (int co = 0; co < kPostC; ++co) {
        float acc = wt.b[l][co];
#pragma unroll
        for (int ci = 0; ci < kPostC; ++ci) acc = fmaf(wt.w[l][co][ci], x[ci], acc);
        y[co] = fmaxf(acc, 0.0f);
      }
#pragma unroll
      for (int c = 0; c < kPostC; ++c) x[c] = y[c];
    }
    float l0 = wt.hb[0], l1 = wt.hb[1];
#pragma unroll
    for (int c = 0; c < kPostC; ++c) {
      l0 = fmaf(wt.hw[0][c], x[c], l0);
      l1 = fmaf(wt.hw[1][c], x[c], l1);
    }
    const int t = img / chunk_slices, sl = img - t * chunk_slices;
    const long long gimg = (long long)t * n_slices_total + slice0 + sl;
    reinterpret_cast<float2*>(logits)[gimg * hw + px] = make_float2(l0, l1);
  }
}

}  // namespace rcu
