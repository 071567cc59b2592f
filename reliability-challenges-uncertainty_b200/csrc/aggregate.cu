// Per-pixel aggregation over MC samples / ensemble members: softmax, running mean, predictive entropy,
// mutual information, variance, argmax — one pass over the per-sample logits (HBM-bound: 8*T+12 B/voxel).
//
// Reference arithmetic (citations relative to the reference root):
//   rechun/dl/customsteps.py:31-36   probs_t = F.softmax(logits_t, 1); torch.stack
//   rechun/dl/customsteps.py:57-71   mean(dim=0); entropy; mutual_info = H(mean) - mean_t H(p_t);
//                                    variance = var(dim=0, unbiased).mean(dim=1)
//   common/utils/torchhelper.py:53-54 entropy = -sum_c where(p > 0, p * log p, 0)
//   bin-dl/brats_test_default.py:97  prediction = np.argmax(probabilities, -1)   (ties -> class 0)
#include "common.cuh"

#ifndef RCU_AGG_TU
#define RCU_AGG_TU 2
#endif
#ifndef RCU_AGG_P
#define RCU_AGG_P 4
#endif
#ifndef RCU_AGG_MINB
#define RCU_AGG_MINB 4
#endif
#ifndef RCU_AGG_FAST_MATH
#define RCU_AGG_FAST_MATH 1
#endif
#ifndef RCU_AGG_DIFF_P        // pairs per thread on logit-difference input (a multiple of two: 16-byte loads cover two pairs)
#define RCU_AGG_DIFF_P 4
#endif
#ifndef RCU_AGG_DIFF_MINB
#define RCU_AGG_DIFF_MINB 4
#endif

namespace rcu {

constexpr int kAggThreads = 256;
constexpr int kAggSampleUnroll = RCU_AGG_TU;

__device__ __forceinline__ void softmax2(float l0, float l1, float& p0, float& p1) {
  // torch's formulation, exp(l - max) / sum, for two classes: the larger logit contributes exp(0) = 1 exactly, the
  // other e = exp(-|l0 - l1|); p_max = 1 / (1 + e), p_min = e * p_max.  One exponential and one reciprocal per
  // pixel-sample instead of two each.  RCU_AGG_FAST_MATH (default): ex2.approx / rcp.approx (2 + 1 ulp: below 2e-7
  // absolute on a probability, against the 1e-6 the parity tests allow) — with logit differences as input the pass is
  // bound by instruction issue, not by HBM, and the IEEE-rounded expf / reciprocal are two thirds of its instructions.
  const float d = l0 - l1;
#if RCU_AGG_FAST_MATH
  const float e = __expf(-fabsf(d));
  const float r = __fdividef(1.0f, 1.0f + e);
#else
  const float e = expf(-fabsf(d));
  const float r = __frcp_rn(1.0f + e);
#endif
  const float q = e * r;
  const bool first = d >= 0.0f;      // NaN logits give NaN probabilities either way
  p0 = first ? r : q;
  p1 = first ? q : r;
}

__device__ __forceinline__ float plogp(float p) { return p > 0.0f ? p * logf(p) : 0.0f; }

struct AggAcc {
  float s0, s1;      // sum_t p
  float h;           // sum_t H(p_t)
  float m0, m1;      // running mean (Welford) for the unbiased variance
  float q0, q1;      // sum of squared deviations
};

// One thread = P pixel pairs (a float4 of interleaved logits, or two float2 of planar values, per pair and sample).
// Pair (thread, k) = warp_base + k * 32 + lane: every load instruction of a warp reads 512 contiguous bytes of one
// sample, and with P = 4 the P loads of a sample are issued back to back, so a warp walks 2 KB runs of each of the
// n_samples streams (long DRAM bursts; deeper unrolling over samples instead opens more streams at once and is slower).
// KIND 0: interleaved logits [t][n][hw][2];  1: planar probabilities [t][n][2][hw];  2: planar logits;
// 3: logit differences l0 - l1 [t][n][hw] (what the fused head writes in rcu_unet_outputs.logit_diff mode: softmax2 only ever
//    uses the difference, so the outputs are bit-identical to KIND 0 at half the bytes).  A lane then owns two ADJACENT
//    pixel pairs per 16-byte load: pair (thread, k) = warp_base + (k / 2) * 64 + lane * 2 + (k & 1).
template <int KIND, bool MI, bool VAR, bool PARTIAL, int P>
__global__ void __launch_bounds__(kAggThreads, KIND == 3 && P > 1 ? RCU_AGG_DIFF_MINB : (P > 1 ? RCU_AGG_MINB : 1))
aggregate_kernel(const float* __restrict__ in, int n_samples, long long n_images, long long hw, float inv_or_scale,
                 float* __restrict__ mean, float* __restrict__ entropy, float* __restrict__ mutual_info,
                 float* __restrict__ variance, unsigned char* __restrict__ prediction, float* __restrict__ foreground,
                 float* __restrict__ multi_out, float* __restrict__ sums, const float* __restrict__ ws_in,
                 float* __restrict__ ws_out) {
  const long long pairs_per_image = hw >> 1;  // hw is even (checked on the host)
  const long long total_pairs = n_images * pairs_per_image;
  const long long sample_stride = n_images * hw * (KIND == 3 ? 1 : 2);
  static_assert(KIND != 3 || P % 2 == 0 || P == 1, "difference input reads two adjacent pairs per 16-byte load");
  const int lane = threadIdx.x & 31;
  const long long warp_id = ((long long)blockIdx.x * kAggThreads + threadIdx.x) >> 5;
  const long long n_warps = ((long long)gridDim.x * kAggThreads) >> 5;
  for (long long wbase = warp_id * (32 * P); wbase < total_pairs; wbase += n_warps * (32 * P)) {
    AggAcc a[P][2] = {};
    const float* src[P];
    long long img_k[P], px_k[P];
    bool on[P];
#pragma unroll
    for (int k = 0; k < P; ++k) {
      const long long pair = (KIND == 3 && P > 1) ? wbase + (k >> 1) * 64 + lane * 2 + (k & 1) : wbase + k * 32 + lane;
      on[k] = pair < total_pairs;
      const long long pr = on[k] ? pair : 0;
      img_k[k] = pr / pairs_per_image;
      px_k[k] = (pr - img_k[k] * pairs_per_image) * 2;
      src[k] = KIND == 3 ? in + img_k[k] * hw + px_k[k] : in + img_k[k] * hw * 2 + (KIND == 0 ? px_k[k] * 2 : px_k[k]);
    }
    if (ws_in != nullptr) {
      // the deterministic weight-scaling pass of McPredictStep (rechun/dl/customsteps.py:23-25) rides in the same launch:
      // one more interleaved-logits sample in, its planar softmax out
      float4 q[P];
#pragma unroll
      for (int k = 0; k < P; ++k) {
        if (KIND == 3) {   // the weight-scaling sample as differences as well: (d, 0) is the same pair of logits to softmax2
          const float2 d2 = on[k] ? ld_stream_f2(ws_in + img_k[k] * hw + px_k[k]) : make_float2(0.f, 0.f);
          q[k] = make_float4(d2.x, 0.f, d2.y, 0.f);
        } else {
          q[k] = on[k] ? ld_stream_f4(ws_in + img_k[k] * hw * 2 + px_k[k] * 2) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
#pragma unroll
      for (int k = 0; k < P; ++k) {
        float a0, a1, b0, b1;
        softmax2(q[k].x, q[k].y, a0, a1);
        softmax2(q[k].z, q[k].w, b0, b1);
        if (on[k]) {
          float* dst = ws_out + img_k[k] * hw * 2 + px_k[k];
          *reinterpret_cast<float2*>(dst) = make_float2(a0, b0);
          *reinterpret_cast<float2*>(dst + hw) = make_float2(a1, b1);
        }
      }
    }
#pragma unroll kAggSampleUnroll
    for (int t = 0; t < n_samples; ++t) {
      float v[P][4];  // [pair][pixel * 2 + class]
#pragma unroll
      for (int k = 0; k < P; ++k) {
        if (!on[k]) { v[k][0] = v[k][1] = v[k][2] = v[k][3] = 0.0f; continue; }
        if (KIND == 0) {
          const float4 q = ld_stream_f4(src[k] + (long long)t * sample_stride);
          v[k][0] = q.x; v[k][1] = q.y; v[k][2] = q.z; v[k][3] = q.w;
        } else if (KIND == 3) {
          if (P > 1) {
            if ((k & 1) == 0) {   // one 16-byte load serves this pair and the next (both on or both off: pairs_per_image is even there)
              const float4 q = ld_stream_f4(src[k] + (long long)t * sample_stride);
              v[k][0] = q.x; v[k][1] = 0.f; v[k][2] = q.y; v[k][3] = 0.f;
              if (k + 1 < P) { v[k + 1][0] = q.z; v[k + 1][1] = 0.f; v[k + 1][2] = q.w; v[k + 1][3] = 0.f; }
            }
          } else {
            const float2 q = ld_stream_f2(src[k] + (long long)t * sample_stride);
            v[k][0] = q.x; v[k][1] = 0.f; v[k][2] = q.y; v[k][3] = 0.f;
          }
        } else {
          const float2 c0 = ld_stream_f2(src[k] + (long long)t * sample_stride);
          const float2 c1 = ld_stream_f2(src[k] + (long long)t * sample_stride + hw);
          v[k][0] = c0.x; v[k][2] = c0.y; v[k][1] = c1.x; v[k][3] = c1.y;
        }
      }
#pragma unroll
      for (int k = 0; k < P; ++k) {
        float p[2][2];
        if (KIND == 1) {
          p[0][0] = v[k][0]; p[0][1] = v[k][1]; p[1][0] = v[k][2]; p[1][1] = v[k][3];
        } else {
          softmax2(v[k][0], v[k][1], p[0][0], p[0][1]);
          softmax2(v[k][2], v[k][3], p[1][0], p[1][1]);
        }
        if (multi_out != nullptr && on[k]) {
          float* dst = multi_out + (long long)t * n_images * hw * 2 + img_k[k] * hw * 2 + px_k[k];
          *reinterpret_cast<float2*>(dst) = make_float2(p[0][0], p[1][0]);
          *reinterpret_cast<float2*>(dst + hw) = make_float2(p[0][1], p[1][1]);
        }
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          AggAcc& acc = a[k][j];
          acc.s0 += p[j][0];
          acc.s1 += p[j][1];
          if (MI) acc.h += -(plogp(p[j][0]) + plogp(p[j][1]));
          if (VAR) {
            if (PARTIAL) {  // raw second moments, combined across ranks later
              acc.q0 += p[j][0] * p[j][0];
              acc.q1 += p[j][1] * p[j][1];
            } else {
              const float inv = 1.0f / (float)(t + 1);
              const float d0 = p[j][0] - acc.m0, d1 = p[j][1] - acc.m1;
              acc.m0 += d0 * inv;
              acc.m1 += d1 * inv;
              acc.q0 += d0 * (p[j][0] - acc.m0);
              acc.q1 += d1 * (p[j][1] - acc.m1);
            }
          }
        }
      }
    }
#pragma unroll
    for (int k = 0; k < P; ++k) {
      if (!on[k]) continue;
      const long long img = img_k[k], px = px_k[k];
      if (PARTIAL) {
        const int planes = 2 + (MI ? 1 : 0) + (VAR ? 2 : 0);
        float* dst = sums + img * planes * hw + px;
        *reinterpret_cast<float2*>(dst) = make_float2(a[k][0].s0, a[k][1].s0);
        *reinterpret_cast<float2*>(dst + hw) = make_float2(a[k][0].s1, a[k][1].s1);
        int pl = 2;
        if (MI) { *reinterpret_cast<float2*>(dst + pl * hw) = make_float2(a[k][0].h, a[k][1].h); ++pl; }
        if (VAR) {
          *reinterpret_cast<float2*>(dst + pl * hw) = make_float2(a[k][0].q0, a[k][1].q0);
          *reinterpret_cast<float2*>(dst + (pl + 1) * hw) = make_float2(a[k][0].q1, a[k][1].q1);
        }
      } else {
        float m0[2], m1[2], ent[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          // torch: sum then divide by T
          m0[j] = __fdiv_rn(a[k][j].s0, inv_or_scale);
          m1[j] = __fdiv_rn(a[k][j].s1, inv_or_scale);
          ent[j] = -(plogp(m0[j]) + plogp(m1[j]));
        }
        float* mdst = mean + img * hw * 2 + px;
        *reinterpret_cast<float2*>(mdst) = make_float2(m0[0], m0[1]);
        *reinterpret_cast<float2*>(mdst + hw) = make_float2(m1[0], m1[1]);
        const long long o = img * hw + px;
        if (entropy) *reinterpret_cast<float2*>(entropy + o) = make_float2(ent[0], ent[1]);
        if (MI) *reinterpret_cast<float2*>(mutual_info + o) =
            make_float2(ent[0] - __fdiv_rn(a[k][0].h, inv_or_scale), ent[1] - __fdiv_rn(a[k][1].h, inv_or_scale));
        if (VAR) {
          const float dn = (float)(n_samples - 1);
          *reinterpret_cast<float2*>(variance + o) =
              make_float2(0.5f * (a[k][0].q0 / dn + a[k][0].q1 / dn), 0.5f * (a[k][1].q0 / dn + a[k][1].q1 / dn));
        }
        if (prediction) {
          uchar2 pr;
          pr.x = m1[0] > m0[0] ? 1 : 0;
          pr.y = m1[1] > m0[1] ? 1 : 0;
          *reinterpret_cast<uchar2*>(prediction + o) = pr;
        }
        if (foreground) *reinterpret_cast<float2*>(foreground + o) = make_float2(m1[0], m1[1]);
      }
    }
  }
}

// sums [n][K][hw] (allreduced across ranks) -> same outputs as the one-pass kernel
__global__ void __launch_bounds__(kAggThreads)
aggregate_finish_kernel(const float* __restrict__ sums, float total_samples, long long n_images, long long hw, int has_mi,
                        int has_var, float* __restrict__ mean, float* __restrict__ entropy, float* __restrict__ mutual_info,
                        float* __restrict__ variance, unsigned char* __restrict__ prediction, float* __restrict__ foreground) {
  const int planes = 2 + (has_mi ? 1 : 0) + (has_var ? 2 : 0);
  const long long total = n_images * hw;
  for (long long i = (long long)blockIdx.x * kAggThreads + threadIdx.x; i < total; i += (long long)gridDim.x * kAggThreads) {
    const long long img = i / hw, px = i - img * hw;
    const float* s = sums + img * planes * hw + px;
    const float s0 = s[0], s1 = s[hw];
    const float m0 = __fdiv_rn(s0, total_samples), m1 = __fdiv_rn(s1, total_samples);
    const float ent = -(plogp(m0) + plogp(m1));
    mean[img * 2 * hw + px] = m0;
    mean[img * 2 * hw + hw + px] = m1;
    if (entropy) entropy[i] = ent;
    int pl = 2;
    if (has_mi) {
      if (mutual_info) mutual_info[i] = ent - __fdiv_rn(s[pl * hw], total_samples);
      ++pl;
    }
    if (has_var && variance) {
      // unbiased variance from raw moments: (sum p^2 - (sum p)^2 / T) / (T - 1), averaged over the two classes.  The
      // difference cancels badly in float for confident pixels (p ~ 1: sum p^2 ~ T), so it is taken in double and clamped
      // at 0 like torch.var, which never goes negative
      const double T = (double)total_samples, dn = T - 1.0;
      const double q0 = s[pl * hw], q1 = s[(pl + 1) * hw], d0 = s0, d1 = s1;
      variance[i] = (float)(0.5 * ((fmax(q0 - d0 * d0 / T, 0.0) + fmax(q1 - d1 * d1 / T, 0.0)) / dn));
    }
    if (prediction) prediction[i] = m1 > m0 ? 1 : 0;
    if (foreground) foreground[i] = m1;
  }
}

template <int KIND, bool PARTIAL>
static int launch_aggregate(bool mi, bool var, const float* in, int n_samples, int64_t n_images, int64_t hw, float denom,
                            float* mean, float* entropy, float* mutual_info, float* variance, uint8_t* prediction,
                            float* foreground, float* multi_out, float* sums, cudaStream_t st, const float* ws_in = nullptr,
                            float* ws_out = nullptr) {
  const long long total_pairs = n_images * (hw / 2);
  if (total_pairs == 0) return RCU_OK;
  // the plain summary (no MI / variance) runs four pairs per thread; the variants with more accumulators keep one
  const bool wide = !mi && !var && total_pairs >= (long long)sm_count() * kAggThreads * 8 && (KIND != 3 || hw % 4 == 0);
  // (difference input: eight pairs, so that a warp still walks 2 KB runs of every sample stream with 16-byte loads)
  constexpr int kWideP = KIND == 3 ? RCU_AGG_DIFF_P : RCU_AGG_P;
  const int per_thread = wide ? kWideP : 1;
  long long blocks = (total_pairs + (long long)kAggThreads * per_thread - 1) / ((long long)kAggThreads * per_thread);
  const long long cap = (long long)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  if (wide) {
    aggregate_kernel<KIND, false, false, PARTIAL, kWideP><<<(unsigned)blocks, kAggThreads, 0, st>>>(
        in, n_samples, (long long)n_images, (long long)hw, denom, mean, entropy, mutual_info, variance, prediction, foreground,
        multi_out, sums, ws_in, ws_out);
    RCU_LAUNCH_CHECK();
    return RCU_OK;
  }
#define RCU_AGG_LAUNCH(MI_, VAR_)                                                                                    \
  aggregate_kernel<KIND, MI_, VAR_, PARTIAL, 1><<<(unsigned)blocks, kAggThreads, 0, st>>>(                            \
      in, n_samples, (long long)n_images, (long long)hw, denom, mean, entropy, mutual_info, variance, prediction,     \
      foreground, multi_out, sums, ws_in, ws_out)
  if (mi && var) RCU_AGG_LAUNCH(true, true);
  else if (mi) RCU_AGG_LAUNCH(true, false);
  else if (var) RCU_AGG_LAUNCH(false, true);
  else RCU_AGG_LAUNCH(false, false);
#undef RCU_AGG_LAUNCH
  RCU_LAUNCH_CHECK();
  return RCU_OK;
}

static int check_agg_common(const float* input, int input_kind, int n_samples, int64_t n_images, int64_t hw) {
  RCU_CHECK_ARG(input != nullptr, "input is NULL");
  RCU_CHECK_ARG(input_kind >= 0 && input_kind <= 3, "input_kind must be 0, 1, 2 or 3");
  RCU_CHECK_ARG(n_samples >= 1 && n_images >= 0 && hw >= 0, "bad sizes: n_samples=%d n_images=%lld hw=%lld", n_samples,
                (long long)n_images, (long long)hw);
  RCU_CHECK_ARG(hw % 2 == 0, "hw=%lld must be even (pixel pairs are processed together)", (long long)hw);
  RCU_CHECK_ARG(reinterpret_cast<uintptr_t>(input) % 16 == 0, "input must be 16-byte aligned");
  return RCU_OK;
}

}  // namespace rcu

using namespace rcu;

extern "C" int rcu_aggregate(const float* input, int input_kind, int n_samples, int64_t n_images, int64_t hw, float* mean,
                             float* entropy, float* mutual_info, float* variance, uint8_t* prediction, float* foreground,
                             float* multi_out, void* stream) {
  int rc = check_agg_common(input, input_kind, n_samples, n_images, hw);
  if (rc) return rc;
  RCU_CHECK_ARG(mean != nullptr, "mean output is NULL");
  RCU_CHECK_ARG(variance == nullptr || n_samples >= 2, "variance needs at least two samples");
  const bool mi = mutual_info != nullptr, var = variance != nullptr;
  cudaStream_t st = (cudaStream_t)stream;
  const float denom = (float)n_samples;
  switch (input_kind) {
    case 0: return launch_aggregate<0, false>(mi, var, input, n_samples, n_images, hw, denom, mean, entropy, mutual_info, variance, prediction, foreground, multi_out, nullptr, st);
    case 1: return launch_aggregate<1, false>(mi, var, input, n_samples, n_images, hw, denom, mean, entropy, mutual_info, variance, prediction, foreground, multi_out, nullptr, st);
    case 3: return launch_aggregate<3, false>(mi, var, input, n_samples, n_images, hw, denom, mean, entropy, mutual_info, variance, prediction, foreground, multi_out, nullptr, st);
    default: return launch_aggregate<2, false>(mi, var, input, n_samples, n_images, hw, denom, mean, entropy, mutual_info, variance, prediction, foreground, multi_out, nullptr, st);
  }
}

extern "C" int rcu_aggregate_ws(const float* logits, int n_samples, int64_t n_images, int64_t hw, const float* ws_logits,
                                float* ws_probabilities, float* mean, float* entropy, float* mutual_info, float* variance,
                                uint8_t* prediction, float* foreground, void* stream) {
  int rc = check_agg_common(logits, 0, n_samples, n_images, hw);
  if (rc) return rc;
  RCU_CHECK_ARG(mean != nullptr && ws_logits != nullptr && ws_probabilities != nullptr, "NULL pointer argument");
  RCU_CHECK_ARG(reinterpret_cast<uintptr_t>(ws_logits) % 16 == 0, "ws_logits must be 16-byte aligned");
  RCU_CHECK_ARG(variance == nullptr || n_samples >= 2, "variance needs at least two samples");
  return launch_aggregate<0, false>(mutual_info != nullptr, variance != nullptr, logits, n_samples, n_images, hw, (float)n_samples, mean,
                                    entropy, mutual_info, variance, prediction, foreground, nullptr, nullptr, (cudaStream_t)stream,
                                    ws_logits, ws_probabilities);
}

extern "C" int rcu_aggregate_ws_diff(const float* logit_diff, int n_samples, int64_t n_images, int64_t hw, const float* ws_logit_diff,
                                     float* ws_probabilities, float* mean, float* entropy, float* mutual_info, float* variance,
                                     uint8_t* prediction, float* foreground, void* stream) {
  int rc = check_agg_common(logit_diff, 3, n_samples, n_images, hw);
  if (rc) return rc;
  RCU_CHECK_ARG(mean != nullptr && ws_logit_diff != nullptr && ws_probabilities != nullptr, "NULL pointer argument");
  RCU_CHECK_ARG(reinterpret_cast<uintptr_t>(ws_logit_diff) % 8 == 0, "ws_logit_diff must be 8-byte aligned");
  RCU_CHECK_ARG(variance == nullptr || n_samples >= 2, "variance needs at least two samples");
  return launch_aggregate<3, false>(mutual_info != nullptr, variance != nullptr, logit_diff, n_samples, n_images, hw, (float)n_samples, mean,
                                    entropy, mutual_info, variance, prediction, foreground, nullptr, nullptr, (cudaStream_t)stream,
                                    ws_logit_diff, ws_probabilities);
}

extern "C" int rcu_aggregate_partial(const float* input, int input_kind, int n_samples, int64_t n_images, int64_t hw,
                                     int want_mi, int want_var, float* sums, void* stream) {
  int rc = check_agg_common(input, input_kind, n_samples, n_images, hw);
  if (rc) return rc;
  RCU_CHECK_ARG(sums != nullptr, "sums output is NULL");
  cudaStream_t st = (cudaStream_t)stream;
  switch (input_kind) {
    case 0: return launch_aggregate<0, true>(want_mi != 0, want_var != 0, input, n_samples, n_images, hw, 1.f, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, sums, st);
    case 1: return launch_aggregate<1, true>(want_mi != 0, want_var != 0, input, n_samples, n_images, hw, 1.f, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, sums, st);
    case 3: return launch_aggregate<3, true>(want_mi != 0, want_var != 0, input, n_samples, n_images, hw, 1.f, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, sums, st);
    default: return launch_aggregate<2, true>(want_mi != 0, want_var != 0, input, n_samples, n_images, hw, 1.f, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, sums, st);
  }
}

extern "C" int rcu_aggregate_finish(const float* sums, int total_samples, int64_t n_images, int64_t hw, int has_mi, int has_var,
                                    float* mean, float* entropy, float* mutual_info, float* variance, uint8_t* prediction,
                                    float* foreground, void* stream) {
  RCU_CHECK_ARG(sums != nullptr && mean != nullptr, "NULL pointer argument");
  RCU_CHECK_ARG(total_samples >= 1 && n_images >= 0 && hw >= 0, "bad sizes");
  RCU_CHECK_ARG(!(has_var && variance) || total_samples >= 2, "variance needs at least two samples");
  const long long total = n_images * hw;
  if (total == 0) return RCU_OK;
  long long blocks = (total + kAggThreads - 1) / kAggThreads;
  const long long cap = (long long)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  aggregate_finish_kernel<<<(unsigned)blocks, kAggThreads, 0, (cudaStream_t)stream>>>(
      sums, (float)total_samples, (long long)n_images, (long long)hw, has_mi, has_var, mean, entropy, mutual_info, variance,
      prediction, foreground);
  RCU_LAUNCH_CHECK();
  return RCU_OK;
}
