// Evaluation-side preparations that sit between the saved maps and the metric kernels:
//   rcu_border_mask              common/utils/labelhelper.py:12-20 (boarder_mask, used by rechun/eval/analysis.py:109-116)
//   rcu_minmax                   entry_np.min() / .max() of RescaleSubjectMinMax (rechun/eval/analysis.py:169-178)
//   rcu_confidence_to_foreground rechun/eval/helper.py:7-22 (rescale_uncertainties + uncertainty_to_foreground_probabilities)
// All three are one pass over byte / float maps: HBM-bound, grid-stride, 16-byte accesses where alignment allows.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstring>

#include "../../include/rcu_b200.h"
#include "common.cuh"

namespace rcu {

// mask = (dist_in <= d_in) * (dist_out <= d_out) with dist_in / dist_out the Euclidean distance transforms of the map
// and of its complement (unit spacing).  For a foreground voxel dist_out = 0 and dist_in <= d_in  <=>  some background
// voxel lies within the ball of radius d_in; symmetrically for a background voxel.  With the radii the reference
// uses (1, 1: analysis.py:113) that is the 6-neighbourhood.  Voxels outside the volume are neither class (scipy's
// transform only sees the array).  The ball offsets come sorted by distance as a kernel parameter (nearest first:
// interior voxels of a border-free region still visit all of them, border voxels stop at the first hit).
constexpr int kMaxBallOffsets = 124;   // radius 3: 123 offsets

struct BallOffsets {
  int n_in, n_out;
  char4 in[kMaxBallOffsets];    // (dx, dy, dz, -)
  char4 out[kMaxBallOffsets];
};

__global__ void __launch_bounds__(256)
border_mask_kernel(const uint8_t* __restrict__ label, int d0, int d1, int d2, const __grid_constant__ BallOffsets ball,
                   uint8_t* __restrict__ mask) {
  // one block row per (z, y) line: the index arithmetic is per line, the neighbour loads of a line hit L1
  for (long long row = blockIdx.y; row < (long long)d0 * d1; row += gridDim.y) {
    const int z = (int)(row / d1), y = (int)(row - (long long)z * d1);
    for (int x = blockIdx.x * 256 + threadIdx.x; x < d2; x += gridDim.x * 256) {
      const long long i = row * d2 + x;
      const bool fg = __ldg(label + i) != 0;
      const int cnt = fg ? ball.n_in : ball.n_out;
      const char4* offs = fg ? ball.in : ball.out;
      bool hit = false;
      for (int k = 0; k < cnt && !hit; ++k) {
        const char4 o = offs[k];
        const int xx = x + o.x, yy = y + o.y, zz = z + o.z;
        if (xx < 0 || xx >= d2 || yy < 0 || yy >= d1 || zz < 0 || zz >= d0) continue;
        hit = (__ldg(label + i + ((long long)o.z * d1 + o.y) * d2 + o.x) != 0) != fg;
      }
      mask[i] = hit ? 1 : 0;
    }
  }
}

static int fill_ball(int r, char4* out) {   // offsets of the ball of radius r without the centre, nearest first
  int n = 0;
  for (int d2 = 1; d2 <= r * r; ++d2)
    for (int dz = -r; dz <= r; ++dz)
      for (int dy = -r; dy <= r; ++dy)
        for (int dx = -r; dx <= r; ++dx)
          if (dz * dz + dy * dy + dx * dx == d2) out[n++] = make_char4((signed char)dx, (signed char)dy, (signed char)dz, 0);
  return n;
}

__device__ __forceinline__ uint32_t ordered_key(float v) {
  const uint32_t b = __float_as_uint(v);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

// out[0] = min key, out[1] = max key (ordered_key space; initialised by the launcher), out[2] = NaN count
__global__ void __launch_bounds__(256) minmax_kernel(const float* __restrict__ v, long long n, uint32_t* __restrict__ out) {
  uint32_t lo = 0xffffffffu, hi = 0u, nan = 0u;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
    const float f = __ldg(v + i);
    if (f != f) { ++nan; continue; }
    const uint32_t k = ordered_key(f);
    lo = min(lo, k);
    hi = max(hi, k);
  }
  for (int o = 16; o > 0; o >>= 1) {
    lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    nan += __shfl_xor_sync(0xffffffffu, nan, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMin(out, lo);
    atomicMax(out + 1, hi);
    if (nan) atomicAdd(out + 2, nan);
  }
}

// p = ((u - lo) / range) * scale + eps  (every operation rounded to float32 on its own, as numpy does), then the
// foreground pseudo-probability q = p * 0.5, flipped to 1 - q where the prediction is 1.  invalid[0] counts values
// that leave [0, 1] (check_min_max raises for them), invalid[1] predictions > 1.
__global__ void __launch_bounds__(256)
confidence_kernel(const float* __restrict__ u, const uint8_t* __restrict__ pred, long long n, int rescale, float lo, float range,
                  float scale, float eps, float* __restrict__ p_out, unsigned long long* __restrict__ invalid) {
  unsigned bad = 0, bad_pred = 0;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
    float p = __ldg(u + i);
    if (rescale) p = __fadd_rn(__fmul_rn(__fdiv_rn(__fsub_rn(p, lo), range), scale), eps);
    if (!(p >= 0.0f && p <= 1.0f)) ++bad;
    const uint8_t c = pred[i];
    if (c > 1) ++bad_pred;
    const float q = __fmul_rn(p, 0.5f);
    p_out[i] = c == 1 ? __fsub_rn(1.0f, q) : q;
  }
  bad = __reduce_add_sync(0xffffffffu, bad);
  bad_pred = __reduce_add_sync(0xffffffffu, bad_pred);
  if ((threadIdx.x & 31) == 0) {
    if (bad) atomicAdd(invalid, (unsigned long long)bad);
    if (bad_pred) atomicAdd(invalid + 1, (unsigned long long)bad_pred);
  }
}

static unsigned grid_for(long long n, int per_sm) {
  long long blocks = (n + 255) / 256;
  const long long cap = (long long)sm_count() * per_sm;
  if (blocks > cap) blocks = cap;
  return (unsigned)(blocks < 1 ? 1 : blocks);
}

}  // namespace rcu

using namespace rcu;

extern "C" int rcu_border_mask(const uint8_t* label, int d0, int d1, int d2, int distance_in, int distance_out, uint8_t* mask, void* stream) {
  RCU_CHECK_ARG(label != nullptr && mask != nullptr, "NULL argument");
  RCU_CHECK_ARG(d0 >= 1 && d1 >= 1 && d2 >= 1, "bad shape %d x %d x %d", d0, d1, d2);
  RCU_CHECK_ARG(distance_in >= 0 && distance_out >= 0, "distances must be >= 0");
  if (distance_in > 3 || distance_out > 3) { set_error("border distances above 3 voxels are not supported"); return RCU_ENOTSUP; }
  BallOffsets ball;
  ball.n_in = fill_ball(distance_in, ball.in);
  ball.n_out = fill_ball(distance_out, ball.out);
  const long long rows = (long long)d0 * d1;
  dim3 grid((unsigned)((d2 + 255) / 256 > 64 ? 64 : (d2 + 255) / 256), (unsigned)(rows > 65535 ? 65535 : rows));
  border_mask_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(label, d0, d1, d2, ball, mask);
  RCU_LAUNCH_CHECK();
  return RCU_OK;
}

extern "C" int rcu_minmax(const float* values, int64_t n, uint32_t* out3, void* stream) {
  RCU_CHECK_ARG(values != nullptr && out3 != nullptr && n >= 1, "NULL argument or empty input");
  cudaStream_t st = (cudaStream_t)stream;
  static const uint32_t init[3] = {0xffffffffu, 0u, 0u};
  RCU_CUDA(cudaMemcpyAsync(out3, init, sizeof(init), cudaMemcpyHostToDevice, st));
  minmax_kernel<<<grid_for(n, 8), 256, 0, st>>>(values, n, out3);
  RCU_LAUNCH_CHECK();
  return RCU_OK;
}

extern "C" float rcu_minmax_decode(uint32_t key) {
  const uint32_t b = (key & 0x80000000u) ? (key & 0x7fffffffu) : ~key;
  float f;
  memcpy(&f, &b, 4);
  return f;
}

extern "C" int rcu_confidence_to_foreground(const float* uncertainty, const uint8_t* prediction, int64_t n, int rescale, float lo,
                                            float range, float scale, float epsilon, float* foreground, uint64_t* invalid2,
                                            void* stream) {
  RCU_CHECK_ARG(uncertainty != nullptr && prediction != nullptr && foreground != nullptr && invalid2 != nullptr, "NULL argument");
  RCU_CHECK_ARG(n >= 0, "negative length");
  cudaStream_t st = (cudaStream_t)stream;
  RCU_CUDA(cudaMemsetAsync(invalid2, 0, 2 * sizeof(uint64_t), st));
  if (n == 0) return RCU_OK;
  confidence_kernel<<<grid_for(n, 16), 256, 0, st>>>(uncertainty, prediction, n, rescale, lo, range, scale, epsilon, foreground,
                                                     reinterpret_cast<unsigned long long*>(invalid2));
  RCU_LAUNCH_CHECK();
  return RCU_OK;
}
