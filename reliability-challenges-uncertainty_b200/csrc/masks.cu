// Materialised Philox keep-scale tables (device kernel + bit-identical host function).
// See philox.cuh for the stream definition and include/rcu_b200.h for the table layout.
#include <vector>

#include "common.cuh"
#include "philox.cuh"

namespace rcu {

__global__ void philox_masks_kernel(uint32_t seed_lo, uint32_t seed_hi, uint32_t thr, float inv_keep,
                                    const int* __restrict__ site_of_channel, const int* __restrict__ channel_in_site,
                                    int total_channels, long long slice_index0, long long n_slices, int sample0,
                                    int n_samples, float* __restrict__ scale) {
  const long long total = (long long)n_samples * n_slices * total_channels;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int col = (int)(i % total_channels);
    const long long row = i / total_channels;
    const long long slice = row % n_slices;
    const int sample = (int)(row / n_slices);
    scale[i] = dropout_scale(seed_lo, seed_hi, thr, inv_keep, (uint32_t)site_of_channel[col],
                             (uint32_t)(slice_index0 + slice), (uint32_t)(sample0 + sample),
                             (uint32_t)channel_in_site[col]);
  }
}

static int check_mask_args(float p_drop, const int* site_channels, int n_sites, int64_t n_slices, int n_samples, const void* out) {
  RCU_CHECK_ARG(p_drop >= 0.0f && p_drop < 1.0f, "dropout p must be in [0, 1), got %f", (double)p_drop);
  RCU_CHECK_ARG(site_channels != nullptr && n_sites >= 1, "site_channels is NULL or empty");
  RCU_CHECK_ARG(n_slices >= 0 && n_samples >= 0 && out != nullptr, "bad sizes or NULL output");
  for (int s = 0; s < n_sites; ++s) RCU_CHECK_ARG(site_channels[s] >= 1, "site %d has %d channels", s, site_channels[s]);
  return RCU_OK;
}

}  // namespace rcu

using namespace rcu;

extern "C" int rcu_philox_masks_host(uint64_t seed, float p_drop, const int* site_channels, int n_sites, int64_t slice_index0,
                                     int64_t n_slices, int sample0, int n_samples, float* scale_host) {
  int rc = check_mask_args(p_drop, site_channels, n_sites, n_slices, n_samples, scale_host);
  if (rc) return rc;
  const uint32_t thr = dropout_threshold_u32(p_drop);
  const float inv_keep = 1.0f / (1.0f - p_drop);
  long long total_channels = 0;
  for (int s = 0; s < n_sites; ++s) total_channels += site_channels[s];
  for (int t = 0; t < n_samples; ++t)
    for (int64_t g = 0; g < n_slices; ++g) {
      float* row = scale_host + ((long long)t * n_slices + g) * total_channels;
      long long off = 0;
      for (int s = 0; s < n_sites; ++s) {
        for (int c = 0; c < site_channels[s]; ++c)
          row[off + c] = dropout_scale((uint32_t)seed, (uint32_t)(seed >> 32), thr, inv_keep, (uint32_t)s,
                                       (uint32_t)(slice_index0 + g), (uint32_t)(sample0 + t), (uint32_t)c);
        off += site_channels[s];
      }
    }
  return RCU_OK;
}

extern "C" int rcu_philox_masks(uint64_t seed, float p_drop, const int* site_channels, int n_sites, int64_t slice_index0,
                                int64_t n_slices, int sample0, int n_samples, float* scale, void* stream) {
  int rc = check_mask_args(p_drop, site_channels, n_sites, n_slices, n_samples, scale);
  if (rc) return rc;
  std::vector<int> site_of, ch_in;
  for (int s = 0; s < n_sites; ++s)
    for (int c = 0; c < site_channels[s]; ++c) {
      site_of.push_back(s);
      ch_in.push_back(c);
    }
  const int total_channels = (int)site_of.size();
  const long long total = (long long)n_samples * n_slices * total_channels;
  if (total == 0) return RCU_OK;
  cudaStream_t st = (cudaStream_t)stream;
  int* d_tab = nullptr;
  RCU_CUDA(cudaMallocAsync(&d_tab, sizeof(int) * 2 * total_channels, st));
  RCU_CUDA(cudaMemcpyAsync(d_tab, site_of.data(), sizeof(int) * total_channels, cudaMemcpyHostToDevice, st));
  RCU_CUDA(cudaMemcpyAsync(d_tab + total_channels, ch_in.data(), sizeof(int) * total_channels, cudaMemcpyHostToDevice, st));
  RCU_CUDA(cudaStreamSynchronize(st));  // the pageable staging vectors die at return
  long long blocks = (total + 255) / 256;
  if (blocks > (long long)sm_count() * 8) blocks = (long long)sm_count() * 8;
  philox_masks_kernel<<<(unsigned)blocks, 256, 0, st>>>((uint32_t)seed, (uint32_t)(seed >> 32), dropout_threshold_u32(p_drop),
                                                        1.0f / (1.0f - p_drop), d_tab, d_tab + total_channels, total_channels,
                                                        (long long)slice_index0, (long long)n_slices, sample0, n_samples, scale);
  cudaError_t e = cudaGetLastError();
  cudaFreeAsync(d_tab, st);
  if (e != cudaSuccess) {
    set_error("philox_masks_kernel launch failed: %s", cudaGetErrorString(e));
    return RCU_ECUDA;
  }
  return RCU_OK;
}
