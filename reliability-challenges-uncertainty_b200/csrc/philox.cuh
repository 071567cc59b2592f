// Philox4x32-10 counter-based generator for Dropout2d keep decisions (host + device).
//
// Stream definition (restated on the host in oracle/restate.py::philox_keep_masks and in
// rcu_philox_masks_host): dropout site s, run-global slice index g, MC sample t, channel c
//     r = philox4x32_10(counter = (c / 4, s, g, t), key = (seed_lo, seed_hi));   keep <=> r[c % 4] >= thr
//     thr = min(2^32 - 1, ceil(p * 2^32))
// The stream depends on neither batching nor which GPU a slice lands on.
// It replaces torch's global-generator stream behind nn.Dropout2d (common/model/unet.py:14-15).
#pragma once
#include <cstdint>
#include <cmath>

namespace rcu {

struct Philox4 { uint32_t v[4]; };

#if defined(__CUDACC__)
#define RCU_HD __host__ __device__ __forceinline__
#else
#define RCU_HD inline
#endif

RCU_HD void philox_mulhilo(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo) {
#if defined(__CUDA_ARCH__)
  lo = a * b;
  hi = __umulhi(a, b);
#else
  const uint64_t p = (uint64_t)a * (uint64_t)b;
  lo = (uint32_t)p;
  hi = (uint32_t)(p >> 32);
#endif
}

RCU_HD Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0, lo0, hi1, lo1;
    philox_mulhilo(0xD2511F53u, c0, hi0, lo0);
    philox_mulhilo(0xCD9E8D57u, c2, hi1, lo1);
    const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  Philox4 out;
  out.v[0] = c0; out.v[1] = c1; out.v[2] = c2; out.v[3] = c3;
  return out;
}

inline uint32_t dropout_threshold_u32(float p) {
  const double t = std::ceil((double)p * 4294967296.0);
  return t >= 4294967295.0 ? 0xFFFFFFFFu : (t <= 0.0 ? 0u : (uint32_t)t);
}

// Keep-scale of one (site, slice, sample, channel): 0 or inv_keep.
RCU_HD float dropout_scale(uint32_t seed_lo, uint32_t seed_hi, uint32_t thr, float inv_keep, uint32_t site,
                           uint32_t slice, uint32_t sample, uint32_t channel) {
  const Philox4 r = philox4x32_10(channel >> 2, site, slice, sample, seed_lo, seed_hi);
  return r.v[channel & 3u] >= thr ? inv_keep : 0.0f;
}

}  // namespace rcu
