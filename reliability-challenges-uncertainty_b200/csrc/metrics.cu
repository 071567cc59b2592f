// Calibration (ECE reliability bins) and uncertainty-error joint histograms — HBM-bound streaming kernels.
//
// Reference arithmetic restated on device (citations relative to the reference root):
//   common/evalutation/numpyfunctions.py:51-69   _binary_calibration: np.digitize against float64
//                                                np.linspace(0, 1+1e-8, n+1) edges + three np.bincount
//   common/evalutation/numpyfunctions.py:86-107  uncertainty(): tp/tn/fp/fn and their intersection with (u > th)
//   bin-eval/eval_uncertainty.py:195-202,239     the 11-threshold sweep (11 passes in the reference, one here)
//
// Design: one pass, 6-7 bytes per voxel (float4 + packed-byte loads, L1 no-allocate).  Every thread owns a
// private column of shared-memory counters (no atomics, no bank conflicts: bank == lane), blocks reduce their
// columns in a fixed order and the last block of each subject (ticket) folds the per-block partials in a fixed
// order, so the float64 confidence sums are deterministic run to run.  Integer tables are exact.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace rcu {

constexpr int kHistThreads = 256;
constexpr int kMaxBlocksPerSubject = 1024;
constexpr int kFoldGroup = 16;                                        // blocks per first-level fold
constexpr int kTicketsPerSubject = 1 + kMaxBlocksPerSubject / kFoldGroup;   // subject ticket + one per group
constexpr int kPartialSlots = 3 * (RCU_MAX_BINS + 1) + 4 * RCU_MAX_UE_CLASSES + 1;
constexpr int kMaxVoxelsPerThread = 60000;  // 16-bit private counters must not wrap
constexpr int kBreakPad = 128;              // break table padded with +inf to a power of two

struct CalibParams {
  float edges[RCU_MAX_BINS + 1];
  int n_bins;
  float range_lo, range_hi;
  int has_range;
};

struct UeParams {
  float breaks32[RCU_MAX_BREAKS];
  double breaks64[RCU_MAX_UE_CLASSES];
  unsigned char seg_class[RCU_MAX_BREAKS + 1];
  int n_breaks;
  int n_classes;
  int search_top;   // first step of the branch-free search (power of two)
  // fused kernel (calibration + U-E on the same float32 p): ONE search over the merged, sorted edge / break list;
  // segment s = #{merged entries <= p}; joint_kj[s] = calibration bin | (U-E class << 8) of that segment
  float joint[kBreakPad];
  unsigned short joint_kj[kBreakPad + 1];
  int joint_n, joint_top;
  int check_range;  // count values outside [0, 1] (value_kind 0: p must be a probability)
};

struct HistOut {
  unsigned long long* count;
  unsigned long long* positives;
  double* conf_sum;
  unsigned long long* ue_counts;
  unsigned long long* invalid;
};

__device__ __forceinline__ int calib_bin(float p, const float* s_edges, int n_bins, float n_bins_f) {
  // k = floor(p * n) is right up to +-1 (fp32 rounding of the product, and the 1e-8 stretch of the edges);
  // the comparison against the exact float32-rounded-up edges settles it.  NaN / negative / >= last edge -> n_bins.
  if (!(p >= 0.0f) || !(p < s_edges[n_bins])) return n_bins;
  int k = min((int)(p * n_bins_f), n_bins - 1);
  k += (p >= s_edges[k + 1]) ? 1 : 0;
  k -= (p < s_edges[k]) ? 1 : 0;
  return k;
}

template <typename T, bool STRICT>
__device__ __forceinline__ int count_breaks(T x, const T* s_breaks, int top) {
  // number of (sorted, +inf padded) breaks b with b <= x (or b < x when STRICT); NaN compares false -> 0
  int lo = 0;
#pragma unroll
  for (int step = kBreakPad / 2; step >= 1; step >>= 1) {
    if (step <= top) {   // warp-uniform
      const T b = s_breaks[lo + step - 1];
      const bool take = STRICT ? (b < x) : (b <= x);
      lo += take ? step : 0;
    }
  }
  return lo;
}

__device__ __forceinline__ void hist_global_fold(int nb1, int ncls, int subject, int blocks_per_subject, HistOut out,
                                                 unsigned int* __restrict__ tickets, unsigned long long* __restrict__ partials);

// Block reduction of the per-thread private columns (fixed order), per-block partials to global memory, and the fold of
// all partials by the subject's last block (ticket) — shared by the generic and the LUT kernels.
__device__ __forceinline__ void hist_block_finish(unsigned char* smem_raw, double* s_conf, unsigned int* s_cntpos, unsigned short* s_ue,
                                                  int nb1, int ncls, unsigned int n_invalid, int subject, int blocks_per_subject,
                                                  HistOut out, unsigned int* __restrict__ tickets,
                                                  unsigned long long* __restrict__ partials) {
  const int tid = threadIdx.x;
  // ---- block reduction of the private columns, fixed order ----
  const int warp = tid >> 5, lane = tid & 31;
  const int n_slots = 3 * nb1 + 4 * ncls + 1;
  unsigned long long* my_partial = partials + ((long long)subject * blocks_per_subject + blockIdx.x) * kPartialSlots;
  for (int s = warp; s < n_slots - 1; s += kHistThreads / 32) {
    if (s >= 2 * nb1 && s < 3 * nb1) {
      const double* col = s_conf + (s - 2 * nb1) * kHistThreads;
      double acc = 0.0;
#pragma unroll
      for (int i = 0; i < kHistThreads / 32; ++i) acc += col[lane + 32 * i];
#pragma unroll
      for (int o = 16; o >= 1; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
      if (lane == 0) my_partial[s] = (unsigned long long)__double_as_longlong(acc);
    } else {
      unsigned long long acc = 0;
      if (s < 2 * nb1) {
        const unsigned int* col = s_cntpos + (s < nb1 ? s : s - nb1) * kHistThreads;
        const int sh = s < nb1 ? 0 : 16;
#pragma unroll
        for (int i = 0; i < kHistThreads / 32; ++i) acc += (col[lane + 32 * i] >> sh) & 0xffffu;
      } else {
        const unsigned short* col = s_ue + (s - 3 * nb1) * kHistThreads;
#pragma unroll
        for (int i = 0; i < kHistThreads / 32; ++i) acc += col[lane + 32 * i];
      }
#pragma unroll
      for (int o = 16; o >= 1; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
      if (lane == 0) my_partial[s] = acc;
    }
  }
  // invalid-key count lives in registers: block-wide sum through a scratch word per warp (the columns are dead now)
  __syncthreads();
  unsigned int* s_scratch = reinterpret_cast<unsigned int*>(smem_raw);
  {
    unsigned int v = n_invalid;
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (lane == 0) s_scratch[warp] = v;
  }
  __syncthreads();
  if (tid == 0) {
    unsigned long long tot = 0;
    for (int w = 0; w < kHistThreads / 32; ++w) tot += s_scratch[w];
    my_partial[n_slots - 1] = tot;
  }

  hist_global_fold(nb1, ncls, subject, blocks_per_subject, out, tickets, partials);
}

// Two-level fold of the per-block partial tables (written to `partials` by every block of the subject before the call).
__device__ __forceinline__ void hist_global_fold(int nb1, int ncls, int subject, int blocks_per_subject, HistOut out,
                                                 unsigned int* __restrict__ tickets, unsigned long long* __restrict__ partials) {
  const int tid = threadIdx.x;
  const int n_slots = 3 * nb1 + 4 * ncls + 1;
  __shared__ int s_is_last;
  // ---- two-level fold in fixed order: the last block of every group of kFoldGroup blocks sums the group (block order),
  // the last group to finish sums the group results (group order).  One thread per slot, all loads of a level in
  // flight together: a level costs about one L2 round trip.  (A single block folding ~300 partials, one slot per warp,
  // used to take as long as the streaming pass itself: 35 us of a 79 us single-subject launch.) ----
  __threadfence();
  __syncthreads();
  const int n_groups = (blocks_per_subject + kFoldGroup - 1) / kFoldGroup;
  const int group = blockIdx.x / kFoldGroup;
  const int g_first = group * kFoldGroup;
  const int g_size = min(kFoldGroup, blocks_per_subject - g_first);
  unsigned int* tk = tickets + (long long)subject * kTicketsPerSubject;
  if (tid == 0) {
    const unsigned int t = atomicAdd(&tk[1 + group], 1u);
    s_is_last = (t == (unsigned int)g_size - 1u);
  }
  __syncthreads();
  if (!s_is_last) return;
  __threadfence();
  unsigned long long* sp = partials + (long long)subject * blocks_per_subject * kPartialSlots;
  const bool is_conf = (tid >= 2 * nb1 && tid < 3 * nb1);
  if (tid < n_slots) {
    unsigned long long v[kFoldGroup];
#pragma unroll
    for (int u = 0; u < kFoldGroup; ++u) v[u] = u < g_size ? __ldcg(sp + (long long)(g_first + u) * kPartialSlots + tid) : 0ull;
    unsigned long long iacc = 0;
    double dacc = 0.0;
#pragma unroll
    for (int u = 0; u < kFoldGroup; ++u) {
      if (is_conf) dacc += __longlong_as_double((long long)v[u]);      // +0.0 for the padding: neutral
      else iacc += v[u];
    }
    // the group's result replaces the partial of its first block (only this block touches the group's partials now)
    __stcg(sp + (long long)g_first * kPartialSlots + tid, is_conf ? (unsigned long long)__double_as_longlong(dacc) : iacc);
  }
  if (tid == 0) tk[1 + group] = 0u;
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    const unsigned int t = atomicAdd(&tk[0], 1u);
    s_is_last = (t == (unsigned int)n_groups - 1u);
  }
  __syncthreads();
  if (!s_is_last) return;
  __threadfence();
  if (tid < n_slots) {
    unsigned long long iacc = 0;
    double dacc = 0.0;
    for (int g0 = 0; g0 < n_groups; g0 += kFoldGroup) {
      unsigned long long v[kFoldGroup];
#pragma unroll
      for (int u = 0; u < kFoldGroup; ++u)
        v[u] = g0 + u < n_groups ? __ldcg(sp + (long long)(g0 + u) * kFoldGroup * kPartialSlots + tid) : 0ull;
#pragma unroll
      for (int u = 0; u < kFoldGroup; ++u) {
        if (is_conf) dacc += __longlong_as_double((long long)v[u]);
        else iacc += v[u];
      }
    }
    const int s = tid;
    if (s < nb1) out.count[(long long)subject * nb1 + s] = iacc;
    else if (s < 2 * nb1) out.positives[(long long)subject * nb1 + (s - nb1)] = iacc;
    else if (s < 3 * nb1) out.conf_sum[(long long)subject * nb1 + (s - 2 * nb1)] = dacc;
    else if (s < n_slots - 1) out.ue_counts[(long long)subject * 4 * ncls + (s - 3 * nb1)] = iacc;
    else if (out.invalid) out.invalid[subject] = iacc;
  }
  if (tid == 0) tk[0] = 0u;  // workspace is reusable by the next stream-ordered call
}

// VK: 0 = float32 p (or float32 uncertainty, same search), 2 = float64 uncertainty; -1 = no U-E part.
template <bool CALIB, int VK, bool HAS_MASK>
__global__ void __launch_bounds__(kHistThreads)
eval_hist_kernel(const float* __restrict__ p, const double* __restrict__ u64v, const unsigned char* __restrict__ pred,
                 const unsigned char* __restrict__ target, const unsigned char* __restrict__ mask,
                 long long voxels_per_subject, int blocks_per_subject, const __grid_constant__ CalibParams cp,
                 const __grid_constant__ UeParams up, HistOut out, unsigned int* __restrict__ tickets,
                 unsigned long long* __restrict__ partials, int vec_ok) {
  constexpr bool UE = VK >= 0;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x;
  const int nb1 = CALIB ? cp.n_bins + 1 : 0;
  const int ncls = UE ? up.n_classes : 0;

  // shared layout: conf (double) | cntpos (u32) | ue (u16) | edges (float) | breaks (float/double) | seg (u8)
  // (offsets are computed as integers and added to smem_raw: a pointer -> integer -> pointer round trip would turn
  // every access below into a generic load)
  const int conf_bytes = nb1 * kHistThreads * (int)sizeof(double);
  const int cntpos_bytes = nb1 * kHistThreads * (int)sizeof(unsigned int);
  const int ue_bytes = 4 * ncls * kHistThreads * (int)sizeof(unsigned short);
  const int tail_off = (conf_bytes + cntpos_bytes + ue_bytes + 15) & ~15;
  const int breaks_d_bytes = (VK == 2 ? RCU_MAX_UE_CLASSES : 0) * (int)sizeof(double);
  const int breaks_f_bytes = (VK == 0 ? kBreakPad : 0) * (int)sizeof(float);
  const int edges_bytes = (CALIB ? RCU_MAX_BINS + 1 : 0) * (int)sizeof(float);
  double* s_conf = reinterpret_cast<double*>(smem_raw);
  unsigned int* s_cntpos = reinterpret_cast<unsigned int*>(smem_raw + conf_bytes);
  unsigned short* s_ue = reinterpret_cast<unsigned short*>(smem_raw + conf_bytes + cntpos_bytes);
  double* s_breaks_d = reinterpret_cast<double*>(smem_raw + tail_off);
  float* s_breaks_f = reinterpret_cast<float*>(smem_raw + tail_off + breaks_d_bytes);
  float* s_edges = reinterpret_cast<float*>(smem_raw + tail_off + breaks_d_bytes + breaks_f_bytes);
  unsigned char* s_seg = smem_raw + tail_off + breaks_d_bytes + breaks_f_bytes + edges_bytes;

  for (int i = tid; i < nb1 * kHistThreads; i += kHistThreads) {
    s_conf[i] = 0.0;
    s_cntpos[i] = 0u;
  }
  for (int i = tid; i < 4 * ncls * kHistThreads; i += kHistThreads) s_ue[i] = 0;
  if (CALIB)
    for (int i = tid; i <= cp.n_bins; i += kHistThreads) s_edges[i] = cp.edges[i];
  constexpr bool JOINT = CALIB && VK == 0;   // one merged search for the calibration bin and the U-E class
  unsigned short* s_kj = reinterpret_cast<unsigned short*>(smem_raw + tail_off + breaks_d_bytes + breaks_f_bytes + edges_bytes + 128);
  if (JOINT) {
    for (int i = tid; i < kBreakPad; i += kHistThreads) s_breaks_f[i] = i < up.joint_n ? up.joint[i] : __int_as_float(0x7f800000);
    for (int i = tid; i <= up.joint_n; i += kHistThreads) s_kj[i] = up.joint_kj[i];
  } else if (VK == 0)
    for (int i = tid; i < kBreakPad; i += kHistThreads) s_breaks_f[i] = i < up.n_breaks ? up.breaks32[i] : __int_as_float(0x7f800000);
  if (VK == 2)
    for (int i = tid; i < RCU_MAX_UE_CLASSES; i += kHistThreads) s_breaks_d[i] = i < up.n_breaks ? up.breaks64[i] : __longlong_as_double(0x7ff0000000000000LL);
  if (UE)
    for (int i = tid; i <= up.n_breaks; i += kHistThreads) s_seg[i] = up.seg_class[i];
  __syncthreads();

  const int subject = blockIdx.y;
  const long long base = (long long)subject * voxels_per_subject;
  // contiguous chunk of this block, in units of 4 voxels
  const long long groups = (voxels_per_subject + 3) >> 2;
  const long long gpb = (groups + blocks_per_subject - 1) / blocks_per_subject;
  const long long g0 = (long long)blockIdx.x * gpb;
  const long long g1 = min(groups, g0 + gpb);
  const float n_bins_f = (float)cp.n_bins;
  unsigned int n_invalid = 0;

  // One voxel, branch-free: every voxel performs the same read-modify-writes on this thread's private counters, with a
  // zero increment when it is masked out (the private columns are bank == lane, so the RMWs are conflict-free; a
  // predicated-off voxel is cheaper than a divergent branch).  A macro, not a lambda: the counters must be addressed
  // as shared memory (LDS/STS), which a by-reference capture degrades to generic loads/stores.
  const float e_last = CALIB ? s_edges[cp.n_bins] : 0.0f;
  const int nbm1 = cp.n_bins - 1;
  const bool has_range = CALIB && cp.has_range;
#define RCU_HIST_ONE(PV, UV, T, D, M)                                                                                   \
  do {                                                                                                                  \
    const float pv_ = (PV);                                                                                             \
    const unsigned int t_ = (T), d_ = (D), m_ = (M);                                                                    \
    if (JOINT) {                                                                                                        \
      const unsigned int kj_ = s_kj[count_breaks<float, false>(pv_, s_breaks_f, up.joint_top)];                         \
      bool use_ = HAS_MASK ? (m_ != 0u) : true;                                                                         \
      if (has_range) use_ = use_ && (pv_ < cp.range_hi) && (pv_ > cp.range_lo);                                         \
      const bool inr_ = (pv_ >= 0.0f) && (pv_ < e_last);                                                                \
      const int ci_ = (inr_ ? (int)(kj_ & 0xffu) : cp.n_bins) * kHistThreads + tid;                                     \
      s_cntpos[ci_] += use_ ? (1u + ((t_ != 0u) ? 65536u : 0u)) : 0u;                                                   \
      s_conf[ci_] += use_ ? (double)pv_ : 0.0;                                                                          \
      n_invalid += !(pv_ >= 0.0f && pv_ <= 1.0f) ? 1u : 0u;                                                             \
      const int row_ = (t_ != 0u) ? ((d_ != 0u) ? 0 : 3) : ((d_ != 0u) ? 2 : 1);                                        \
      s_ue[(row_ * ncls + (int)(kj_ >> 8)) * kHistThreads + tid] += 1;                                                  \
    } else if (CALIB) {                                                                                                 \
      bool use_ = HAS_MASK ? (m_ != 0u) : true;                                                                         \
      if (has_range) use_ = use_ && (pv_ < cp.range_hi) && (pv_ > cp.range_lo);                                         \
      const bool inr_ = (pv_ >= 0.0f) && (pv_ < e_last);          /* NaN / negative / >= last edge -> slot n_bins */     \
      int k_ = max(0, min(__float2int_rz(pv_ * n_bins_f), nbm1)); /* floor(p * n) is right up to +-1 ...           */     \
      k_ += (pv_ >= s_edges[k_ + 1]) ? 1 : 0;                     /* ... the float32-rounded-up edges settle it     */     \
      k_ -= (pv_ < s_edges[min(k_, nbm1 + 1)]) ? 1 : 0;                                                                 \
      k_ = inr_ ? k_ : cp.n_bins;                                                                                       \
      const int ci_ = k_ * kHistThreads + tid;                                                                          \
      s_cntpos[ci_] += use_ ? (1u + ((t_ != 0u) ? 65536u : 0u)) : 0u;                                                   \
      s_conf[ci_] += use_ ? (double)pv_ : 0.0;                                                                          \
    }                                                                                                                   \
    if (UE && !JOINT) {                                                                                                 \
      /* U-E tables are unmasked in the fused kernel (the mask belongs to the calibration part); in the U-E-only */      \
      /* kernel HAS_MASK applies to them (UncertaintyErrorDiceNumpy(with_mask=True)).                              */      \
      const bool useu_ = (HAS_MASK && !CALIB) ? (m_ != 0u) : true;                                                      \
      int idx_;                                                                                                         \
      if (VK == 2) {                                                                                                    \
        idx_ = count_breaks<double, true>((UV), s_breaks_d, up.search_top);                                             \
      } else {                                                                                                          \
        idx_ = count_breaks<float, false>(pv_, s_breaks_f, up.search_top);                                              \
        const bool bad_ = up.check_range ? !(pv_ >= 0.0f && pv_ <= 1.0f) : !(pv_ == pv_);                               \
        n_invalid += (useu_ && bad_) ? 1u : 0u;                                                                         \
      }                                                                                                                 \
      const int j_ = s_seg[idx_];                                                                                       \
      const int row_ = (t_ != 0u) ? ((d_ != 0u) ? 0 : 3) : ((d_ != 0u) ? 2 : 1); /* tp, tn, fp, fn */                    \
      s_ue[(row_ * ncls + j_) * kHistThreads + tid] += useu_ ? 1 : 0;                                                   \
    }                                                                                                                   \
  } while (0)

  if (vec_ok) {
    constexpr int U = 4;
    for (long long g = g0 + tid; g < g1; g += (long long)kHistThreads * U) {
      float4 pv[U];
      double2 uv[U][2];
      unsigned int tv[U], dv[U], mv[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long gg = g + (long long)u * kHistThreads;
        const bool in = gg < g1 && (gg * 4 + 3 < voxels_per_subject);
        pv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        tv[u] = dv[u] = 0u;
        mv[u] = 0u;
        if (in) {
          const long long v = base + gg * 4;
          if (VK != 2) pv[u] = ld_stream_f4(p + v);
          if (VK == 2) {
            uv[u][0] = *reinterpret_cast<const double2*>(u64v + v);
            uv[u][1] = *reinterpret_cast<const double2*>(u64v + v + 2);
          }
          tv[u] = ld_stream_u32(target + v);
          if (UE) dv[u] = ld_stream_u32(pred + v);
          if (HAS_MASK) mv[u] = ld_stream_u32(mask + v);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long gg = g + (long long)u * kHistThreads;
        if (gg >= g1) break;
        if (gg * 4 + 3 < voxels_per_subject) {
          const float pe[4] = {pv[u].x, pv[u].y, pv[u].z, pv[u].w};
          double ue[4] = {0., 0., 0., 0.};
          if (VK == 2) { ue[0] = uv[u][0].x; ue[1] = uv[u][0].y; ue[2] = uv[u][1].x; ue[3] = uv[u][1].y; }
#pragma unroll
          for (int e = 0; e < 4; ++e)
            RCU_HIST_ONE(pe[e], ue[e], (tv[u] >> (8 * e)) & 0xffu, (dv[u] >> (8 * e)) & 0xffu, (mv[u] >> (8 * e)) & 0xffu);
        } else {  // ragged last group of the subject
          for (long long v = gg * 4; v < voxels_per_subject; ++v) {
            const long long a = base + v;
            RCU_HIST_ONE(VK != 2 ? p[a] : 0.f, VK == 2 ? u64v[a] : 0., target[a], UE ? pred[a] : 0, HAS_MASK ? mask[a] : 1);
          }
        }
      }
    }
  } else {
    for (long long g = g0 + tid; g < g1; g += kHistThreads) {
      const long long vend = min(voxels_per_subject, g * 4 + 4);
      for (long long v = g * 4; v < vend; ++v) {
        const long long a = base + v;
        RCU_HIST_ONE(VK != 2 ? p[a] : 0.f, VK == 2 ? u64v[a] : 0., target[a], UE ? pred[a] : 0, HAS_MASK ? mask[a] : 1);
      }
    }
  }
#undef RCU_HIST_ONE
  __syncthreads();

  hist_block_finish(smem_raw, s_conf, s_cntpos, s_ue, nb1, ncls, n_invalid, subject, blocks_per_subject, out, tickets, partials);
}


// Fused calibration + U-E pass, lean variant for the common case (float32 p, 16-byte aligned inputs, no threshold_range,
// at most one merged edge / break point per 1/NB-wide bucket of p).  The generic kernel is instruction bound at ~105
// instructions per voxel (two searches, index arithmetic); here the joint segment of p comes from a bucket table —
// s = base[bucket] + (p >= break[bucket]), exact because p * NB is exact for a power-of-two NB — and the per-segment
// table holds the counter offsets pre-multiplied, which leaves ~45 instructions per voxel.  Same private-column counters,
// same reduction, bit-identical tables.
// LUT4: the bucket entry carries its (at most three) in-bucket break points itself — {base, b0, b1, b2}, +inf padded —
// so the segment is ONE 16-byte table read and three compares: no second dependent lookup and no data-dependent loop,
// which lets the compiler interleave the 16 voxels of an iteration (the scan variant serialises them on its branch).
template <bool HAS_MASK, bool LUT4>
__global__ void __launch_bounds__(kHistThreads)
eval_fused_lut_kernel(const float* __restrict__ p, const unsigned char* __restrict__ pred, const unsigned char* __restrict__ target,
                      const unsigned char* __restrict__ mask, long long voxels_per_subject, int blocks_per_subject,
                      const __grid_constant__ CalibParams cp, const __grid_constant__ UeParams up, int n_buckets, HistOut out,
                      unsigned int* __restrict__ tickets, unsigned long long* __restrict__ partials) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x;
  const int nb1 = cp.n_bins + 1;
  const int ncls = up.n_classes;
  const int conf_bytes = nb1 * kHistThreads * (int)sizeof(double);
  const int cntpos_bytes = nb1 * kHistThreads * (int)sizeof(unsigned int);
  const int ue_bytes = 4 * ncls * kHistThreads * (int)sizeof(unsigned short);
  const int tail_off = (conf_bytes + cntpos_bytes + ue_bytes + 15) & ~15;
  double* s_conf = reinterpret_cast<double*>(smem_raw);
  unsigned int* s_cntpos = reinterpret_cast<unsigned int*>(smem_raw + conf_bytes);
  unsigned short* s_ue = reinterpret_cast<unsigned short*>(smem_raw + conf_bytes + cntpos_bytes);
  float* s_joint = reinterpret_cast<float*>(smem_raw + tail_off);                           // [kBreakPad] merged list, +inf padded
  unsigned int* s_seg2 = reinterpret_cast<unsigned int*>(smem_raw + tail_off + kBreakPad * 4);   // [kBreakPad + 1] k*T | (j*T) << 16
  unsigned int* s_lut = reinterpret_cast<unsigned int*>(smem_raw + tail_off + kBreakPad * 4 + (kBreakPad + 8) * 4);   // [n_buckets] base | n_inside << 8
  uint4* s_lut4 = reinterpret_cast<uint4*>(s_lut);                                                                   // LUT4: [n_buckets] {base, b0, b1, b2}

  for (int i = tid; i < nb1 * kHistThreads; i += kHistThreads) {
    s_conf[i] = 0.0;
    s_cntpos[i] = 0u;
  }
  for (int i = tid; i < 4 * ncls * kHistThreads; i += kHistThreads) s_ue[i] = 0;
  for (int i = tid; i < kBreakPad; i += kHistThreads) s_joint[i] = i < up.joint_n ? up.joint[i] : __int_as_float(0x7f800000);
  for (int i = tid; i <= up.joint_n; i += kHistThreads) {
    const unsigned int kj = up.joint_kj[i];
    s_seg2[i] = ((kj & 0xffu) * kHistThreads) | (((kj >> 8) * kHistThreads) << 16);
  }
  __syncthreads();
  const float nb_f = (float)n_buckets, inv_nb = 1.0f / nb_f;
  for (int b = tid; b < n_buckets; b += kHistThreads) {
    const float lo = (float)b * inv_nb;                                            // exact (power-of-two bucket count)
    const float hi = b + 1 < n_buckets ? (float)(b + 1) * inv_nb : __int_as_float(0x7f800000);
    const int base = count_breaks<float, false>(lo, s_joint, up.joint_top);        // entries <= lo
    // entries strictly inside (lo, hi): the break points of the float arithmetic cluster a few ulps apart around each
    // threshold, so a bucket may hold several; they are resolved by a short scan (most buckets hold none)
    const int upto = b + 1 < n_buckets ? count_breaks<float, false>(__uint_as_float(__float_as_uint(hi) - 1u), s_joint, up.joint_top) : up.joint_n;
    if (LUT4) {
      uint4 e;
      e.x = (unsigned int)base;
      e.y = __float_as_uint(base + 0 < upto ? s_joint[base + 0] : __int_as_float(0x7f800000));
      e.z = __float_as_uint(base + 1 < upto ? s_joint[base + 1] : __int_as_float(0x7f800000));
      e.w = __float_as_uint(base + 2 < upto ? s_joint[base + 2] : __int_as_float(0x7f800000));
      s_lut4[b] = e;
    } else {
      s_lut[b] = (unsigned int)base | ((unsigned int)(upto - base) << 8);
    }
  }
  __syncthreads();

  const int subject = blockIdx.y;
  const long long base_v = (long long)subject * voxels_per_subject;
  const long long groups = (voxels_per_subject + 3) >> 2;
  const long long gpb = (groups + blocks_per_subject - 1) / blocks_per_subject;
  const long long g0 = (long long)blockIdx.x * gpb;
  const long long g1 = min(groups, g0 + gpb);
  const float e_last = cp.edges[cp.n_bins];
  const unsigned int nb_off = (unsigned int)cp.n_bins * kHistThreads;
  const int row_stride = ncls * kHistThreads;
  unsigned int* cnt_col = s_cntpos + tid;
  double* conf_col = s_conf + tid;
  unsigned short* ue_col = s_ue + tid;
  unsigned int n_invalid = 0;

#define RCU_LUT_ONE(PV, T, D, M)                                                                                        \
  do {                                                                                                                  \
    const float pv_ = (PV);                                                                                             \
    const unsigned int t_ = (T), d_ = (D), m_ = (M);                                                                    \
    const int b_ = max(0, min(__float2int_rz(pv_ * nb_f), n_buckets - 1));                                              \
    const bool nonneg_ = pv_ >= 0.0f;                                 /* false for NaN */                               \
    int s_;                                                                                                             \
    if (LUT4) {                                                                                                         \
      const uint4 le_ = s_lut4[b_];                                                                                     \
      s_ = (int)le_.x + ((pv_ >= __uint_as_float(le_.y)) ? 1 : 0) + ((pv_ >= __uint_as_float(le_.z)) ? 1 : 0) +         \
           ((pv_ >= __uint_as_float(le_.w)) ? 1 : 0);                                                                   \
    } else {                                                                                                            \
      const unsigned int le_ = s_lut[b_];                                                                               \
      s_ = (int)(le_ & 0xffu);                                                                                          \
      for (int i_ = (int)(le_ & 0xffu), e_ = i_ + (int)(le_ >> 8); i_ < e_; ++i_) s_ += (pv_ >= s_joint[i_]) ? 1 : 0;  \
    }                                                                                                                   \
    s_ = nonneg_ ? s_ : 0;                                            /* the search counts 0 entries for p < 0 / NaN */ \
    const unsigned int sg_ = s_seg2[s_];                                                                                \
    const bool use_ = HAS_MASK ? (m_ != 0u) : true;                                                                     \
    const unsigned int ko_ = (nonneg_ && pv_ < e_last) ? (sg_ & 0xffffu) : nb_off;                                      \
    cnt_col[ko_] += use_ ? (1u + ((t_ != 0u) ? 65536u : 0u)) : 0u;                                                      \
    conf_col[ko_] += use_ ? (double)pv_ : 0.0;                                                                          \
    n_invalid += !(nonneg_ && pv_ <= 1.0f) ? 1u : 0u;                                                                   \
    const int row_ = (t_ != 0u) ? ((d_ != 0u) ? 0 : 3) : ((d_ != 0u) ? 2 : 1);     /* tp, tn, fp, fn */                 \
    ue_col[row_ * row_stride + (int)(sg_ >> 16)] += 1;                                                                  \
  } while (0)

  constexpr int U = 4;
  for (long long g = g0 + tid; g < g1; g += (long long)kHistThreads * U) {
    float4 pv[U];
    unsigned int tv[U], dv[U], mv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long gg = g + (long long)u * kHistThreads;
      const bool in = gg < g1 && (gg * 4 + 3 < voxels_per_subject);
      pv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      tv[u] = dv[u] = mv[u] = 0u;
      if (in) {
        const long long v = base_v + gg * 4;
        pv[u] = ld_stream_f4(p + v);
        tv[u] = ld_stream_u32(target + v);
        dv[u] = ld_stream_u32(pred + v);
        if (HAS_MASK) mv[u] = ld_stream_u32(mask + v);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long gg = g + (long long)u * kHistThreads;
      if (gg >= g1) break;
      if (gg * 4 + 3 < voxels_per_subject) {
        const float pe[4] = {pv[u].x, pv[u].y, pv[u].z, pv[u].w};
#pragma unroll
        for (int e = 0; e < 4; ++e)
          RCU_LUT_ONE(pe[e], (tv[u] >> (8 * e)) & 0xffu, (dv[u] >> (8 * e)) & 0xffu, (mv[u] >> (8 * e)) & 0xffu);
      } else {  // ragged last group of the subject
        for (long long v = gg * 4; v < voxels_per_subject; ++v) {
          const long long a = base_v + v;
          RCU_LUT_ONE(p[a], target[a], pred[a], HAS_MASK ? mask[a] : 1);
        }
      }
    }
  }
#undef RCU_LUT_ONE
  __syncthreads();
  hist_block_finish(smem_raw, s_conf, s_cntpos, s_ue, nb1, ncls, n_invalid, subject, blocks_per_subject, out, tickets, partials);
}


__device__ __forceinline__ unsigned int swar_nonzero(unsigned int w) {   // per byte: 1 if the byte is non-zero
  return ((w | ((w & 0x7f7f7f7fu) + 0x7f7f7f7fu)) >> 7) & 0x01010101u;
}

#include "eval_atom.cuh"

// Confusion matrix with pymia 0.2.1 semantics (ConfusionMatrix: prediction == 1 / == 0 against label == 1 / == 0),
// reached from np_fn.dice / confusion_matrx / accuracy (common/evalutation/numpyfunctions.py:128-151).  2 B/voxel.
__global__ void __launch_bounds__(256)
confusion_kernel(const unsigned char* __restrict__ pred, const unsigned char* __restrict__ target, long long vps,
                 unsigned long long* __restrict__ out /* [S][4] tp tn fp fn */, int vec_ok) {
  const int subject = blockIdx.y;
  const long long base = (long long)subject * vps;
  unsigned int c[4] = {0u, 0u, 0u, 0u};
  auto one = [&](unsigned int d, unsigned int t) {
    c[0] += (d == 1u && t == 1u);
    c[1] += (d == 0u && t == 0u);
    c[2] += (d == 1u && t == 0u);
    c[3] += (d == 0u && t == 1u);
  };
  const long long groups = (vps + 15) >> 4;
  for (long long g = (long long)blockIdx.x * 256 + threadIdx.x; g < groups; g += (long long)gridDim.x * 256) {
    const long long v = g * 16;
    if (vec_ok && v + 15 < vps) {
      const uint4 dv = ld_stream_u4(pred + base + v), tv = ld_stream_u4(target + base + v);
      const unsigned int dw[4] = {dv.x, dv.y, dv.z, dv.w}, tw[4] = {tv.x, tv.y, tv.z, tv.w};
#pragma unroll
      for (int w = 0; w < 4; ++w)
#pragma unroll
        for (int e = 0; e < 4; ++e) one((dw[w] >> (8 * e)) & 0xffu, (tw[w] >> (8 * e)) & 0xffu);
    } else {
      for (long long a = v; a < min(vps, v + 16); ++a) one(pred[base + a], target[base + a]);
    }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    unsigned int v = c[k];
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(out + subject * 4 + k, (unsigned long long)v);
  }
}

static size_t tickets_bytes(int n_subjects) { return ((size_t)n_subjects * kTicketsPerSubject * sizeof(unsigned int) + 255) & ~size_t(255); }
static long long partial_blocks_cap(int n_subjects) { return n_subjects > 2048 ? n_subjects : 2048; }
// Workspace layout: per-block partial tables grow from the front, the tickets sit at the very end.  Calls with different
// n_subjects on the same workspace (same workspace_bytes) then never see each other's partials in their ticket words,
// which must read zero at entry (every launch leaves the tickets it used at zero).
static unsigned int* tickets_of(void* workspace, size_t workspace_bytes, int n_subjects) {
  return reinterpret_cast<unsigned int*>(reinterpret_cast<unsigned char*>(workspace) + (workspace_bytes & ~size_t(255)) - tickets_bytes(n_subjects));
}
// Per-subject accumulator tables of the shared-atomic kernel (64-bit global reductions), right below the tickets and, like
// them, zero at rest.
static size_t accs_bytes(int n_subjects) { return ((size_t)n_subjects * kPartialSlots * sizeof(unsigned long long) + 255) & ~size_t(255); }
static unsigned long long* accs_of(void* workspace, size_t workspace_bytes, int n_subjects) {
  return reinterpret_cast<unsigned long long*>(reinterpret_cast<unsigned char*>(tickets_of(workspace, workspace_bytes, n_subjects)) - accs_bytes(n_subjects));
}

static size_t hist_smem_bytes(bool calib, int vk, int n_bins, int n_classes) {
  const int nb1 = calib ? n_bins + 1 : 0;
  const int ncls = vk >= 0 ? n_classes : 0;
  size_t b = (size_t)nb1 * kHistThreads * (sizeof(double) + sizeof(unsigned int)) + (size_t)4 * ncls * kHistThreads * sizeof(unsigned short);
  b = (b + 15) & ~size_t(15);
  b += (vk == 2 ? RCU_MAX_UE_CLASSES * sizeof(double) : 0) + (vk == 0 ? kBreakPad * sizeof(float) : 0) +
       (calib ? (RCU_MAX_BINS + 1) * sizeof(float) : 0) + 128 + (kBreakPad + 1) * sizeof(unsigned short) + 64;
  return b < 64 ? 64 : b;
}

template <bool CALIB, int VK, bool HAS_MASK>
static int launch_hist(const float* p, const double* u64v, const uint8_t* pred, const uint8_t* target, const uint8_t* mask,
                       int64_t vps, int n_subjects, const CalibParams& cp, const UeParams& up, HistOut out, void* workspace,
                       size_t workspace_bytes, cudaStream_t stream) {
  RCU_CHECK_ARG(workspace_bytes >= rcu_metrics_workspace_bytes(n_subjects), "metrics workspace too small: %zu < %zu",
                workspace_bytes, rcu_metrics_workspace_bytes(n_subjects));
  const int sms = sm_count();
  // one wave of fat blocks: the per-block fixed cost (zeroing ~60 KB of private counters, the column reduction) and the
  // last block's fold over all partials are what a small launch pays, the streaming part is short
  static const int blocks_per_sm = [] { const char* e = std::getenv("RCU_HIST_BPSM"); return e ? std::atoi(e) : 3; }();
  long long bps = ((long long)sms * blocks_per_sm + n_subjects - 1) / n_subjects;
  const long long groups = (vps + 3) / 4;
  const long long min_groups_per_block = 256;  // do not shred small subjects into blocks with < 1 group / thread
  if (bps * min_groups_per_block > groups) bps = groups / min_groups_per_block;
  if (bps < 1) bps = 1;
  if (bps > kMaxBlocksPerSubject) bps = kMaxBlocksPerSubject;
  if (bps * n_subjects > partial_blocks_cap(n_subjects)) bps = partial_blocks_cap(n_subjects) / n_subjects;
  if (bps < 1) bps = 1;
  const long long per_thread = ((groups + bps - 1) / bps + kHistThreads - 1) / kHistThreads * 4;
  RCU_CHECK_ARG(per_thread <= kMaxVoxelsPerThread, "subject of %lld voxels is too large for one launch (split it)", (long long)vps);
  RCU_CHECK_ARG(n_subjects <= 65535, "n_subjects %d exceeds grid.y limit", n_subjects);

  const bool aligned = ((VK == 2 ? (reinterpret_cast<uintptr_t>(u64v) % 16 == 0) : (reinterpret_cast<uintptr_t>(p) % 16 == 0)) &&
                        reinterpret_cast<uintptr_t>(target) % 4 == 0 && (pred == nullptr || reinterpret_cast<uintptr_t>(pred) % 4 == 0) &&
                        (mask == nullptr || reinterpret_cast<uintptr_t>(mask) % 4 == 0));
  const int vec_ok = aligned && (n_subjects == 1 || vps % 4 == 0);

  auto kern = eval_hist_kernel<CALIB, VK, HAS_MASK>;
  const size_t smem = hist_smem_bytes(CALIB, VK, cp.n_bins, up.n_classes);
  RCU_CHECK_ARG(smem <= 227 * 1024, "bin/class configuration needs %zu bytes of shared memory", smem);
  static size_t configured[64] = {0};
  int dev = 0;
  RCU_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || smem > configured[dev]) {
    RCU_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (dev >= 0 && dev < 64) configured[dev] = smem;
  }
  unsigned int* tickets = tickets_of(workspace, workspace_bytes, n_subjects);
  unsigned long long* partials = reinterpret_cast<unsigned long long*>(workspace);
  dim3 grid((unsigned)bps, (unsigned)n_subjects);
  kern<<<grid, kHistThreads, smem, stream>>>(p, u64v, pred, target, mask, (long long)vps, (int)bps, cp, up, out, tickets, partials, vec_ok);
  RCU_LAUNCH_CHECK();
  return RCU_OK;
}

static int fill_calib(CalibParams& cp, const float* edges, int n_bins, float lo, float hi) {
  RCU_CHECK_ARG(n_bins >= 1 && n_bins <= RCU_MAX_BINS, "n_bins must be in [1, %d], got %d", RCU_MAX_BINS, n_bins);
  RCU_CHECK_ARG(edges != nullptr, "edges_f32 is NULL");
  for (int i = 0; i <= n_bins; ++i) {
    cp.edges[i] = edges[i];
    RCU_CHECK_ARG(i == 0 || edges[i] > edges[i - 1], "edges must be strictly increasing");
  }
  cp.n_bins = n_bins;
  cp.has_range = !(lo != lo) && !(hi != hi);
  cp.range_lo = lo;
  cp.range_hi = hi;
  return RCU_OK;
}

static int fill_ue(UeParams& up, int value_kind, const float* b32, const double* b64, int n_breaks, const uint8_t* seg, int n_classes) {
  RCU_CHECK_ARG(n_classes >= 1 && n_classes <= RCU_MAX_UE_CLASSES, "n_classes must be in [1, %d], got %d", RCU_MAX_UE_CLASSES, n_classes);
  RCU_CHECK_ARG(seg != nullptr, "seg_class is NULL");
  if (value_kind == 2) {
    RCU_CHECK_ARG(n_breaks >= 0 && n_breaks < RCU_MAX_UE_CLASSES, "float64 mode takes at most %d thresholds", RCU_MAX_UE_CLASSES - 1);
    RCU_CHECK_ARG(b64 != nullptr || n_breaks == 0, "breaks_f64 is NULL");
    for (int i = 0; i < n_breaks; ++i) {
      up.breaks64[i] = b64[i];
      RCU_CHECK_ARG(i == 0 || b64[i] >= b64[i - 1], "thresholds must be sorted ascending");
    }
  } else {
    RCU_CHECK_ARG(n_breaks >= 0 && n_breaks <= RCU_MAX_BREAKS, "at most %d break points", RCU_MAX_BREAKS);
    RCU_CHECK_ARG(b32 != nullptr || n_breaks == 0, "breaks_f32 is NULL");
    for (int i = 0; i < n_breaks; ++i) {
      up.breaks32[i] = b32[i];
      RCU_CHECK_ARG(i == 0 || b32[i] >= b32[i - 1], "break points must be sorted ascending");
    }
  }
  for (int i = 0; i <= n_breaks; ++i) {
    RCU_CHECK_ARG(seg[i] < n_classes, "seg_class[%d]=%d is not below n_classes=%d", i, (int)seg[i], n_classes);
    up.seg_class[i] = seg[i];
  }
  up.n_breaks = n_breaks;
  up.n_classes = n_classes;
  // branch-free search: steps top, top/2, ..., 1 reach counts up to 2*top-1 and read indices <= 2*top-2,
  // which stay inside the +inf padded tables (128 float / 32 double entries).
  int top = 1;
  while (2 * top - 1 < n_breaks) top *= 2;
  up.search_top = top;
  up.check_range = value_kind == 0 ? 1 : 0;
  return RCU_OK;
}

}  // namespace rcu

using namespace rcu;

extern "C" size_t rcu_metrics_workspace_bytes(int n_subjects) {
  if (n_subjects < 1) n_subjects = 1;
  const size_t partial_bytes = (size_t)partial_blocks_cap(n_subjects) * kPartialSlots * sizeof(unsigned long long);
  return ((partial_bytes + 255) & ~size_t(255)) + accs_bytes(n_subjects) + tickets_bytes(n_subjects) + 256;   // + slack for the round-down in tickets_of
}

extern "C" int rcu_metrics_workspace_init(void* workspace, size_t workspace_bytes, void* stream) {
  RCU_CHECK_ARG(workspace != nullptr, "workspace is NULL");
  RCU_CHECK_ARG(workspace_bytes >= rcu_metrics_workspace_bytes(1), "workspace too small");
  RCU_CUDA(cudaMemsetAsync(workspace, 0, workspace_bytes, (cudaStream_t)stream));
  return RCU_OK;
}

extern "C" int rcu_calib_hist(const float* p, const uint8_t* target, const uint8_t* mask, int64_t vps, int n_subjects,
                              const float* edges_f32, int n_bins, float range_lo, float range_hi, uint64_t* count,
                              uint64_t* positives, double* conf_sum, void* workspace, size_t workspace_bytes, void* stream) {
  RCU_CHECK_ARG(p && target && count && positives && conf_sum && workspace, "NULL pointer argument");
  RCU_CHECK_ARG(vps >= 0 && n_subjects >= 1, "bad sizes: voxels_per_subject=%lld n_subjects=%d", (long long)vps, n_subjects);
  CalibParams cp;
  UeParams up = {};
  int rc = fill_calib(cp, edges_f32, n_bins, range_lo, range_hi);
  if (rc) return rc;
  HistOut out = {reinterpret_cast<unsigned long long*>(count), reinterpret_cast<unsigned long long*>(positives), conf_sum, nullptr, nullptr};
  if (mask) return launch_hist<true, -1, true>(p, nullptr, nullptr, target, mask, vps, n_subjects, cp, up, out, workspace, workspace_bytes, (cudaStream_t)stream);
  return launch_hist<true, -1, false>(p, nullptr, nullptr, target, nullptr, vps, n_subjects, cp, up, out, workspace, workspace_bytes, (cudaStream_t)stream);
}

extern "C" int rcu_ue_hist(const void* values, int value_kind, const uint8_t* prediction, const uint8_t* target,
                           const uint8_t* mask, int64_t vps, int n_subjects, const float* breaks_f32,
                           const double* breaks_f64, int n_breaks, const uint8_t* seg_class, int n_classes,
                           uint64_t* ue_counts, uint64_t* invalid, void* workspace, size_t workspace_bytes, void* stream) {
  RCU_CHECK_ARG(values && prediction && target && ue_counts && workspace, "NULL pointer argument");
  RCU_CHECK_ARG(vps >= 0 && n_subjects >= 1, "bad sizes: voxels_per_subject=%lld n_subjects=%d", (long long)vps, n_subjects);
  RCU_CHECK_ARG(value_kind >= 0 && value_kind <= 2, "value_kind must be 0, 1 or 2");
  CalibParams cp = {};
  UeParams up;
  int rc = fill_ue(up, value_kind, breaks_f32, breaks_f64, n_breaks, seg_class, n_classes);
  if (rc) return rc;
  HistOut out = {nullptr, nullptr, nullptr, reinterpret_cast<unsigned long long*>(ue_counts), reinterpret_cast<unsigned long long*>(invalid)};
  cudaStream_t st = (cudaStream_t)stream;
  if (value_kind == 2) {
    const double* u = reinterpret_cast<const double*>(values);
    if (mask) return launch_hist<false, 2, true>(nullptr, u, prediction, target, mask, vps, n_subjects, cp, up, out, workspace, workspace_bytes, st);
    return launch_hist<false, 2, false>(nullptr, u, prediction, target, nullptr, vps, n_subjects, cp, up, out, workspace, workspace_bytes, st);
  }
  const float* pf = reinterpret_cast<const float*>(values);
  if (mask) return launch_hist<false, 0, true>(pf, nullptr, prediction, target, mask, vps, n_subjects, cp, up, out, workspace, workspace_bytes, st);
  return launch_hist<false, 0, false>(pf, nullptr, prediction, target, nullptr, vps, n_subjects, cp, up, out, workspace, workspace_bytes, st);
}

extern "C" int rcu_eval_fused(const float* p, const uint8_t* prediction, const uint8_t* target, const uint8_t* mask,
                              int64_t vps, int n_subjects, const float* edges_f32, int n_bins, const float* breaks_f32,
                              int n_breaks, const uint8_t* seg_class, int n_classes, uint64_t* count, uint64_t* positives,
                              double* conf_sum, uint64_t* ue_counts, uint64_t* invalid, void* workspace, size_t workspace_bytes,
                              void* stream) {
  RCU_CHECK_ARG(p && prediction && target && count && positives && conf_sum && ue_counts && workspace, "NULL pointer argument");
  RCU_CHECK_ARG(vps >= 0 && n_subjects >= 1, "bad sizes: voxels_per_subject=%lld n_subjects=%d", (long long)vps, n_subjects);
  CalibParams cp;
  UeParams up;
  int rc = fill_calib(cp, edges_f32, n_bins, nanf(""), nanf(""));
  if (rc) return rc;
  rc = fill_ue(up, 0, breaks_f32, nullptr, n_breaks, seg_class, n_classes);
  if (rc) return rc;
  {
    // merged list of the inner calibration edges (bin k = #{edges[1..n_bins-1] <= p} for p inside [0, last edge)) and the
    // U-E breaks (class = seg_class[#{breaks <= p}]): both are "count of entries <= p", so one search serves both
    RCU_CHECK_ARG((n_bins - 1) + n_breaks <= kBreakPad - 1, "too many bins + break points for the fused kernel");
    int ie = 1, ib = 0, n = 0;
    up.joint_kj[0] = (unsigned short)(0 | (up.seg_class[0] << 8));
    while (ie < n_bins || ib < n_breaks) {
      const bool take_edge = ib >= n_breaks || (ie < n_bins && cp.edges[ie] <= up.breaks32[ib]);
      up.joint[n] = take_edge ? cp.edges[ie] : up.breaks32[ib];
      if (take_edge) ++ie; else ++ib;
      ++n;
      up.joint_kj[n] = (unsigned short)((ie - 1) | (up.seg_class[ib] << 8));
    }
    up.joint_n = n;
    int top = 1;
    while (2 * top - 1 < n) top *= 2;
    up.joint_top = top;
  }
  HistOut out = {reinterpret_cast<unsigned long long*>(count), reinterpret_cast<unsigned long long*>(positives), conf_sum,
                 reinterpret_cast<unsigned long long*>(ue_counts), reinterpret_cast<unsigned long long*>(invalid)};
  cudaStream_t st = (cudaStream_t)stream;
  {
    // shared-atomic kernel (eval_atom.cuh): the default for aligned float32 inputs.  RCU_HIST_ATOM=0 falls through to the
    // private-column kernel below (A/B and cross-checks).  (Round 2 also tried private 16-bit counters per joint cell: same
    // instruction count as the bucket-table kernel, one block of 8 warps per SM, slower — removed.)
    static const bool allow_atom = [] { const char* e = std::getenv("RCU_HIST_ATOM"); return !(e && e[0] == '0'); }();
    static const int atom_bits = [] { const char* e = std::getenv("RCU_HIST_ATOM_BITS"); return e ? std::atoi(e) : 10; }();
    constexpr int atom_threads = kAtomThreads;
    const bool aligned_a = reinterpret_cast<uintptr_t>(p) % 16 == 0 && reinterpret_cast<uintptr_t>(target) % 4 == 0 &&
                           reinterpret_cast<uintptr_t>(prediction) % 4 == 0 && (mask == nullptr || reinterpret_cast<uintptr_t>(mask) % 4 == 0);
    if (allow_atom && aligned_a && (n_subjects == 1 || vps % 4 == 0) && atom_bits >= 4 && atom_bits <= kAtomMaxLutBits) {
      AtomParams ap;
      float vals[kBreakPad + RCU_MAX_BINS + 4];
      int nv = 0;
      vals[nv++] = 0.0f;
      for (int i = 1; i <= n_bins; ++i) vals[nv++] = cp.edges[i];
      const float oneplus = nextafterf(1.0f, 2.0f);
      vals[nv++] = oneplus;
      for (int i = 0; i < n_breaks; ++i) vals[nv++] = up.breaks32[i];
      bool ok = true;
      for (int i = 0; i < nv; ++i) ok = ok && vals[i] >= 0.0f && vals[i] < INFINITY;   // false for NaN
      std::sort(vals, vals + nv);
      nv = (int)(std::unique(vals, vals + nv) - vals);
      ok = ok && nv <= kBreakPad - 1;
      // integer confidence sums need edge[1] >= 2^-K with K <= 8 (q = p * 2^(23+K) < 2^32 up to the last edge)
      int K = 0;
      while (K <= 8 && std::ldexp(1.0f, -K) > cp.edges[1]) ++K;
      ok = ok && (n_bins == 1 || K <= 8) && cp.edges[n_bins] <= 1.5f;
      const int n_seg = nv + 1;
      const size_t smem_a = (size_t)n_seg * kAtomSegBytes + ((size_t)(1u << atom_bits) + 1) * 8;
      ok = ok && n_seg <= kAtomMaxSegs && smem_a + 6 * 1024 <= 227 * 1024;
      // blocks: every block at most kAtomMaxVoxelsPerBlock voxels; whole waves of the resident slots (one block per SM)
      const int sms = sm_count();
      const long long groups = (vps + 3) / 4;
      // (the partial region only holds one float64 per block here)
      long long max_bps = std::min<long long>(kMaxBlocksPerSubject, partial_blocks_cap(n_subjects) * (long long)kPartialSlots / n_subjects);
      if (max_bps * atom_threads > groups) max_bps = groups / atom_threads;   // at least one group per thread
      if (max_bps < 1) max_bps = 1;
      const long long min_bps = (vps + kAtomMaxVoxelsPerBlock - 1) / kAtomMaxVoxelsPerBlock;
      ok = ok && min_bps <= max_bps && n_subjects <= 65535;
      if (ok) {
        const long long slots = sms;
        // model: waves x (fixed cost per block + streaming time of a block's share); a block streams ~5.5 voxels / ns
        // (measured), so the grid is sized to fill whole waves of the 148 slots as long as the waves are few
        static const int bps_env = [] { const char* e = std::getenv("RCU_HIST_ATOM_BPS"); return e ? std::atoi(e) : 0; }();
        long long bps = min_bps > 1 ? min_bps : 1;
        double best = 1e300;
        for (long long b = bps; b <= max_bps; ++b) {
          const long long waves = (b * n_subjects + slots - 1) / slots;
          const double cost = (double)waves * (4000.0 + (double)vps / (double)b / 5.5);
          if (cost < best) { best = cost; bps = b; }
          if (waves > 16) break;
        }
        if (bps_env > 0) bps = std::max<long long>(min_bps, std::min<long long>(bps_env, max_bps));
        for (int i = 0; i < kBreakPad; ++i) ap.list[i] = i < nv ? vals[i] : INFINITY;
        ap.n_list = nv;
        int top = 1;
        while (2 * top - 1 < nv) top *= 2;
        ap.top = top;
        ap.attr[0] = (unsigned int)n_bins | ((unsigned int)up.seg_class[0] << 8) | (1u << 16);   // negative / NaN
        ap.seg_sum_lo = ap.seg_sum_hi = -1;
        for (int sg = 1; sg <= nv; ++sg) {
          const float r = vals[sg - 1];   // the smallest value of segment sg
          int k = 0, nb = 0;
          for (int i = 1; i <= n_bins; ++i) k += cp.edges[i] <= r ? 1 : 0;
          for (int i = 0; i < n_breaks; ++i) nb += up.breaks32[i] <= r ? 1 : 0;
          ap.attr[sg] = (unsigned int)k | ((unsigned int)up.seg_class[nb] << 8) | ((oneplus <= r ? 1u : 0u) << 16);
          if (k >= 1 && ap.seg_sum_lo < 0) ap.seg_sum_lo = sg;
          if (k >= n_bins && ap.seg_sum_hi < 0) ap.seg_sum_hi = sg;
        }
        if (ap.seg_sum_hi < 0) ap.seg_sum_hi = nv + 1;
        if (ap.seg_sum_lo < 0) ap.seg_sum_lo = ap.seg_sum_hi;
        for (int k = 0; k <= n_bins + 1 && k < RCU_MAX_BINS + 2; ++k) ap.bin_seg[k] = (unsigned char)ap.seg_sum_hi;
        for (int sg = nv; sg >= 1; --sg) {
          const int k = (int)(ap.attr[sg] & 0xffu);
          if (k <= n_bins) ap.bin_seg[k] = (unsigned char)sg;   // first segment of bin k
        }
        for (int k = n_bins - 1; k >= 0; --k)
          if (ap.bin_seg[k] > ap.bin_seg[k + 1]) ap.bin_seg[k] = ap.bin_seg[k + 1];   // empty bins (cannot happen for increasing edges)
        ap.q_scale = std::ldexp(1.0f, 23 + K);
        ap.q_inv = std::ldexp(1.0, -(23 + K));
        ap.edge1 = cp.edges[1];
        ap.lut_bits = atom_bits;
        {
          // the device's bucket map, on the host (same IEEE arithmetic): widest span of a bucket that holds several entries
          auto bucket_of = [&](float v) {
            const float sat = v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v);
            const float t = sat + 1.0f;
            uint32_t bits;
            std::memcpy(&bits, &t, 4);
            return (bits - 0x3f800000u) >> (23 - atom_bits);
          };
          auto bits_of = [](float v) { uint32_t b; std::memcpy(&b, &v, 4); return b; };
          uint32_t span = 0;
          for (int i = 0; i < nv;) {
            int j = i;
            while (j + 1 < nv && bucket_of(vals[j + 1]) == bucket_of(vals[i])) ++j;
            if (j > i) span = std::max(span, bits_of(vals[j]) - bits_of(vals[i]));
            i = j + 1;
          }
          ap.span_ulps = span + 1u;
        }
        RCU_CHECK_ARG(workspace_bytes >= rcu_metrics_workspace_bytes(n_subjects), "metrics workspace too small");
        unsigned int* tickets = tickets_of(workspace, workspace_bytes, n_subjects);
        unsigned long long* partials = reinterpret_cast<unsigned long long*>(workspace);
        dim3 grid((unsigned)bps, (unsigned)n_subjects);
        static size_t configured[2][64] = {{0}};
        int dev = 0;
        RCU_CUDA(cudaGetDevice(&dev));
        auto kern = mask ? eval_fused_atom_kernel<true> : eval_fused_atom_kernel<false>;
        const int mi = mask ? 1 : 0;
        if (dev < 0 || dev >= 64 || smem_a > configured[mi][dev]) {
          RCU_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_a));
          if (dev >= 0 && dev < 64) configured[mi][dev] = smem_a;
        }
        kern<<<grid, atom_threads, smem_a, st>>>(p, prediction, target, mask, (long long)vps, (int)bps, n_bins, n_classes, ap, out, tickets, partials,
                                                 accs_of(workspace, workspace_bytes, n_subjects));
        RCU_LAUNCH_CHECK();
        return RCU_OK;
      }
    }
  }
  {
    // lean bucket-table kernel for aligned inputs (the generic kernel handles the rest)
    static const bool allow_lut = [] { const char* e = std::getenv("RCU_HIST_LUT"); return !(e && e[0] == '0'); }();
    const bool aligned = reinterpret_cast<uintptr_t>(p) % 16 == 0 && reinterpret_cast<uintptr_t>(target) % 4 == 0 &&
                         reinterpret_cast<uintptr_t>(prediction) % 4 == 0 && (mask == nullptr || reinterpret_cast<uintptr_t>(mask) % 4 == 0);
    // measured (gpurun sweep, r02): 3 blocks / SM with a 512-entry table for a few subjects per launch (48 us per 8.9 M-voxel
    // subject), 6 blocks / SM with a 256-entry table when many subjects share the launch (31 us per subject at 50)
    static const int lut_buckets_env = [] { const char* e = std::getenv("RCU_HIST_BUCKETS"); return e ? std::atoi(e) : 0; }();
    const int lut_buckets = lut_buckets_env > 0 ? lut_buckets_env : (n_subjects >= 8 ? 256 : 512);
    const int n_buckets = (allow_lut && aligned && (n_subjects == 1 || vps % 4 == 0)) ? lut_buckets : 0;
    if (n_buckets > 0) {
      RCU_CHECK_ARG(workspace_bytes >= rcu_metrics_workspace_bytes(n_subjects), "metrics workspace too small");
      const int sms = sm_count();
      static const int lut_bpsm_env = [] { const char* e = std::getenv("RCU_HIST_BPSM"); return e ? std::atoi(e) : 0; }();
      const int lut_bpsm = lut_bpsm_env > 0 ? lut_bpsm_env : (n_subjects >= 8 ? 6 : 3);
      long long bps = ((long long)sms * lut_bpsm + n_subjects - 1) / n_subjects;
      const long long groups = (vps + 3) / 4;
      if (bps * 256 > groups) bps = groups / 256;
      if (bps < 1) bps = 1;
      if (bps > kMaxBlocksPerSubject) bps = kMaxBlocksPerSubject;
      if (bps * n_subjects > partial_blocks_cap(n_subjects)) bps = partial_blocks_cap(n_subjects) / n_subjects;
      if (bps < 1) bps = 1;
      const long long per_thread = ((groups + bps - 1) / bps + kHistThreads - 1) / kHistThreads * 4;
      RCU_CHECK_ARG(per_thread <= kMaxVoxelsPerThread, "subject of %lld voxels is too large for one launch (split it)", (long long)vps);
      RCU_CHECK_ARG(n_subjects <= 65535, "n_subjects %d exceeds grid.y limit", n_subjects);
      const int nb1 = n_bins + 1;
      // in-bucket break points: the 16-byte table entry holds three; a denser cluster takes the scanning variant
      int max_inside = 0;
      {
        int run = 0, run_bucket = -1;
        for (int i = 0; i < up.joint_n; ++i) {
          const float fb = up.joint[i] * (float)n_buckets;          // exact: power-of-two bucket count
          int b = fb >= (float)n_buckets ? n_buckets - 1 : (int)fb;
          if (b < 0) b = 0;
          const bool on_edge = fb < (float)n_buckets && fb == (float)b;   // equal to the bucket's lower bound: part of `base`
          if (on_edge) continue;
          run = (b == run_bucket) ? run + 1 : 1;
          run_bucket = b;
          if (run > max_inside) max_inside = run;
        }
      }
      static const bool allow_lut4 = [] { const char* e = std::getenv("RCU_HIST_LUT4"); return !(e && e[0] == '0'); }();
      const bool lut4 = allow_lut4 && max_inside <= 3 && (n_buckets & (n_buckets - 1)) == 0;
      size_t smem = (size_t)nb1 * kHistThreads * 12 + (size_t)4 * n_classes * kHistThreads * 2;
      smem = ((smem + 15) & ~size_t(15)) + kBreakPad * 4 + (kBreakPad + 8) * 4 + (size_t)n_buckets * (lut4 ? 16 : 4) + 64;
      unsigned int* tickets = tickets_of(workspace, workspace_bytes, n_subjects);
      unsigned long long* partials = reinterpret_cast<unsigned long long*>(workspace);
      dim3 grid((unsigned)bps, (unsigned)n_subjects);
      static size_t configured[4][64] = {{0}};
      int dev = 0;
      RCU_CUDA(cudaGetDevice(&dev));
      const int mi = (mask ? 1 : 0) + (lut4 ? 2 : 0);
      auto kern = mask ? (lut4 ? eval_fused_lut_kernel<true, true> : eval_fused_lut_kernel<true, false>)
                       : (lut4 ? eval_fused_lut_kernel<false, true> : eval_fused_lut_kernel<false, false>);
      if (dev < 0 || dev >= 64 || smem > configured[mi][dev]) {
        RCU_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (dev >= 0 && dev < 64) configured[mi][dev] = smem;
      }
      kern<<<grid, kHistThreads, smem, st>>>(p, prediction, target, mask, (long long)vps, (int)bps, cp, up, n_buckets, out, tickets, partials);
      RCU_LAUNCH_CHECK();
      return RCU_OK;
    }
  }
  if (mask) return launch_hist<true, 0, true>(p, nullptr, prediction, target, mask, vps, n_subjects, cp, up, out, workspace, workspace_bytes, st);
  return launch_hist<true, 0, false>(p, nullptr, prediction, target, nullptr, vps, n_subjects, cp, up, out, workspace, workspace_bytes, st);
}

#ifdef RCU_ATOM_TRACE
extern "C" int rcu_debug_atom_trace(unsigned long long* host_out, int n_words) {
  RCU_CUDA(cudaDeviceSynchronize());
  RCU_CUDA(cudaMemcpyFromSymbol(host_out, rcu::g_atom_trace, sizeof(unsigned long long) * (size_t)n_words));
  return RCU_OK;
}
#endif

extern "C" int rcu_confusion(const uint8_t* prediction, const uint8_t* target, int64_t vps, int n_subjects, uint64_t* counts,
                             void* stream) {
  RCU_CHECK_ARG(prediction && target && counts, "NULL pointer argument");
  RCU_CHECK_ARG(vps >= 0 && n_subjects >= 1 && n_subjects <= 65535, "bad sizes: voxels_per_subject=%lld n_subjects=%d", (long long)vps, n_subjects);
  cudaStream_t st = (cudaStream_t)stream;
  RCU_CUDA(cudaMemsetAsync(counts, 0, sizeof(uint64_t) * 4 * n_subjects, st));
  if (vps == 0) return RCU_OK;
  // a thread's 32-bit partial counters see at most 16 * ceil(groups / (grid*256)) voxels
  const long long groups = (vps + 15) / 16;
  long long bx = ((long long)sm_count() * 8 + n_subjects - 1) / n_subjects;
  if (bx > (groups + 255) / 256) bx = (groups + 255) / 256;
  if (bx < 1) bx = 1;
  const int vec_ok = reinterpret_cast<uintptr_t>(prediction) % 16 == 0 && reinterpret_cast<uintptr_t>(target) % 16 == 0 &&
                     (n_subjects == 1 || vps % 16 == 0);
  dim3 grid((unsigned)bx, (unsigned)n_subjects);
  confusion_kernel<<<grid, 256, 0, st>>>(prediction, target, (long long)vps, reinterpret_cast<unsigned long long*>(counts), vec_ok);
  RCU_LAUNCH_CHECK();
  return RCU_OK;
}
