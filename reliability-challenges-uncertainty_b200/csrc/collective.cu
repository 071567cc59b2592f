// The two exchange steps of the sharded hot path (the reference has no distributed code — its only
// multi-GPU mechanism is nn.DataParallel in training, common/trainloop/context.py:223-233):
//
//   (i)  MC samples / ensemble members split over the GPUs of one box: every rank holds fp32 partial sums
//        [n_images][K][hw] (rcu_aggregate_partial).  rcu_aggregate_finish_peer is the fused "reduce, then finish" step
//        over NVLink peer memory: rank r reads pixel slice r of EVERY rank's sums through mapped peer pointers, adds them
//        in rank order (deterministic, and bit-identical on all ranks because every pixel is reduced exactly once),
//        applies MultiPredictionSummary's arithmetic (rechun/dl/customsteps.py:57-71) and stores the outputs into every
//        rank's output region.  One kernel between two flag barriers; no NCCL call, no second pass over the sums.
//        rcu_allreduce_probsum is the NCCL route to the same result (all-reduce, then rcu_aggregate_finish on each rank).
//   (ii) the per-subject ECE / U-E tables of subjects whose slices span ranks: rcu_allreduce_counts, one grouped NCCL
//        all-reduce of the uint64 count tables and the float64 confidence sums (< 1 KB per subject, latency bound).
//
// NCCL is resolved at run time from the libnccl.so.2 already mapped into the process (PyTorch's), so the library has no
// link-time dependency on it and loads on boxes without NCCL (the collective entries then return RCU_ENCCL).
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>

#include "common.cuh"

namespace rcu {

struct NcclApi {
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};

static const NcclApi& nccl_api() {
  static NcclApi api = [] {
    NcclApi a;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL | RTLD_NOLOAD);   // the copy the process already uses (PyTorch's)
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return a;
#define RCU_NCCL_SYM(name) *reinterpret_cast<void**>(&a.name) = dlsym(h, "nccl" #name)
    RCU_NCCL_SYM(GetUniqueId); RCU_NCCL_SYM(CommInitRank); RCU_NCCL_SYM(CommDestroy); RCU_NCCL_SYM(AllReduce);
    RCU_NCCL_SYM(GroupStart); RCU_NCCL_SYM(GroupEnd); RCU_NCCL_SYM(GetErrorString);
#undef RCU_NCCL_SYM
    a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.AllReduce && a.GroupStart && a.GroupEnd && a.GetErrorString;
    return a;
  }();
  return api;
}

#define RCU_NCCL(call)                                                                                   \
  do {                                                                                                   \
    ncclResult_t r__ = (call);                                                                           \
    if (r__ != ncclSuccess) {                                                                            \
      ::rcu::set_error("%s failed: %s (%s:%d)", #call, nccl_api().GetErrorString(r__), __FILE__, __LINE__); \
      return RCU_ENCCL;                                                                                  \
    }                                                                                                    \
  } while (0)

static int need_nccl() {
  if (!nccl_api().ok) {
    set_error("libnccl.so.2 could not be loaded (%s)", dlerror() ? dlerror() : "symbols missing");
    return RCU_ENCCL;
  }
  return RCU_OK;
}

}  // namespace rcu

using namespace rcu;

struct rcu_comm {
  ncclComm_t comm = nullptr;
  int n_ranks = 0, rank = 0, device = 0;
};

static_assert(sizeof(ncclUniqueId) == RCU_COMM_ID_BYTES, "RCU_COMM_ID_BYTES must equal sizeof(ncclUniqueId)");

extern "C" int rcu_comm_unique_id(void* id_out) {
  RCU_CHECK_ARG(id_out != nullptr, "NULL argument");
  int rc = need_nccl();
  if (rc) return rc;
  ncclUniqueId id;
  RCU_NCCL(nccl_api().GetUniqueId(&id));
  std::memcpy(id_out, &id, sizeof(id));
  return RCU_OK;
}

extern "C" int rcu_comm_create(const void* id_in, int n_ranks, int rank, int device, rcu_comm** out) {
  RCU_CHECK_ARG(id_in != nullptr && out != nullptr, "NULL argument");
  RCU_CHECK_ARG(n_ranks >= 1 && rank >= 0 && rank < n_ranks, "bad rank %d / %d", rank, n_ranks);
  *out = nullptr;
  int rc = need_nccl();
  if (rc) return rc;
  RCU_CUDA(cudaSetDevice(device));
  ncclUniqueId id;
  std::memcpy(&id, id_in, sizeof(id));
  rcu_comm* c = new rcu_comm();
  c->n_ranks = n_ranks; c->rank = rank; c->device = device;
  ncclResult_t r = nccl_api().CommInitRank(&c->comm, n_ranks, id, rank);
  if (r != ncclSuccess) {
    set_error("ncclCommInitRank failed: %s", nccl_api().GetErrorString(r));
    delete c;
    return RCU_ENCCL;
  }
  *out = c;
  return RCU_OK;
}

extern "C" void rcu_comm_destroy(rcu_comm* comm) {
  if (!comm) return;
  if (comm->comm && nccl_api().ok) nccl_api().CommDestroy(comm->comm);
  delete comm;
}

extern "C" int rcu_allreduce_probsum(rcu_comm* comm, float* sums, int64_t n, void* stream) {
  RCU_CHECK_ARG(comm != nullptr && comm->comm != nullptr, "NULL communicator");
  RCU_CHECK_ARG(sums != nullptr || n == 0, "NULL buffer");
  RCU_CHECK_ARG(n >= 0, "negative element count");
  if (n == 0 || comm->n_ranks == 1) return RCU_OK;
  RCU_NCCL(nccl_api().AllReduce(sums, sums, (size_t)n, ncclFloat32, ncclSum, comm->comm, (cudaStream_t)stream));
  return RCU_OK;
}

extern "C" int rcu_allreduce_counts(rcu_comm* comm, uint64_t* counts, int64_t n_counts, double* conf_sums, int64_t n_conf, void* stream) {
  RCU_CHECK_ARG(comm != nullptr && comm->comm != nullptr, "NULL communicator");
  RCU_CHECK_ARG(n_counts >= 0 && n_conf >= 0 && (counts != nullptr || n_counts == 0) && (conf_sums != nullptr || n_conf == 0), "bad buffers");
  if (comm->n_ranks == 1) return RCU_OK;
  RCU_NCCL(nccl_api().GroupStart());
  ncclResult_t r1 = ncclSuccess, r2 = ncclSuccess;
  if (n_counts) r1 = nccl_api().AllReduce(counts, counts, (size_t)n_counts, ncclUint64, ncclSum, comm->comm, (cudaStream_t)stream);
  if (n_conf) r2 = nccl_api().AllReduce(conf_sums, conf_sums, (size_t)n_conf, ncclFloat64, ncclSum, comm->comm, (cudaStream_t)stream);
  ncclResult_t r3 = nccl_api().GroupEnd();
  RCU_NCCL(r1);
  RCU_NCCL(r2);
  RCU_NCCL(r3);
  return RCU_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// Peer-memory exchange regions (CUDA IPC; one process per GPU on one box)
// ---------------------------------------------------------------------------------------------------------------------
extern "C" int rcu_peer_region_alloc(size_t bytes, int device, void** ptr, void* ipc_handle_out) {
  RCU_CHECK_ARG(ptr != nullptr && ipc_handle_out != nullptr && bytes > 0, "bad arguments");
  static_assert(sizeof(cudaIpcMemHandle_t) == RCU_IPC_HANDLE_BYTES, "RCU_IPC_HANDLE_BYTES must equal sizeof(cudaIpcMemHandle_t)");
  RCU_CUDA(cudaSetDevice(device));
  void* p = nullptr;
  RCU_CUDA(cudaMalloc(&p, bytes));
  RCU_CUDA(cudaMemset(p, 0, bytes));            // the signal words must start at zero
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    set_error("cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
    cudaFree(p);
    return RCU_ECUDA;
  }
  std::memcpy(ipc_handle_out, &h, sizeof(h));
  *ptr = p;
  return RCU_OK;
}

extern "C" int rcu_peer_region_open(const void* ipc_handle, int device, void** ptr) {
  RCU_CHECK_ARG(ptr != nullptr && ipc_handle != nullptr, "NULL argument");
  RCU_CUDA(cudaSetDevice(device));
  cudaIpcMemHandle_t h;
  std::memcpy(&h, ipc_handle, sizeof(h));
  RCU_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return RCU_OK;
}

extern "C" int rcu_peer_region_close(void* ptr) {
  if (ptr) RCU_CUDA(cudaIpcCloseMemHandle(ptr));
  return RCU_OK;
}

extern "C" int rcu_peer_region_free(void* ptr) {
  if (ptr) RCU_CUDA(cudaFree(ptr));
  return RCU_OK;
}

namespace rcu {

constexpr int kPeerMaxRanks = 16;
constexpr int kPeerThreads = 256;

struct PeerPtrs { unsigned char* base[kPeerMaxRanks]; };   // region base of every rank as mapped into THIS process

// Flag barrier over the regions' signal words: rank r writes `epoch` into word [r] of every rank's signal row `row`, then
// waits until all n words of its own row carry it.  Ordered with the data by system-scope release / acquire.
__global__ void peer_barrier_kernel(PeerPtrs regions, int n_ranks, int rank, int row, unsigned int epoch) {
  const int i = threadIdx.x;
  if (i < n_ranks) {
    __threadfence_system();
    unsigned int* theirs = reinterpret_cast<unsigned int*>(regions.base[i]) + row * kPeerMaxRanks + rank;
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(theirs), "r"(epoch) : "memory");
    const unsigned int* mine = reinterpret_cast<const unsigned int*>(regions.base[rank]) + row * kPeerMaxRanks + i;
    unsigned int v;
    long long spins = 0;
    do {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
      if (++spins > (1ll << 31)) {   // a lost peer must not hang the GPU
        printf("rcu peer barrier: rank %d timed out waiting for rank %d (row %d epoch %u, saw %u)\n", rank, i, row, epoch, v);
        __trap();
      }
    } while ((int)(v - epoch) < 0);
    __threadfence_system();
  }
}

__device__ __forceinline__ float plogp_c(float p) { return p > 0.0f ? p * logf(p) : 0.0f; }

// Pixel slice [lo, hi) of the (n_images * hw) pixels: sum the K planes over all ranks (rank order), finish, store everywhere.
// Pixels are processed in pairs (hw is even): 8-byte loads / stores per plane.
__global__ void __launch_bounds__(kPeerThreads)
aggregate_finish_peer_kernel(PeerPtrs regions, int n_ranks, long long lo_pair, long long hi_pair, long long hw, int planes, int has_mi,
                             int has_var, float total_samples, size_t off_sums, size_t off_mean, size_t off_entropy, size_t off_mi,
                             size_t off_var, size_t off_fg, size_t off_pred) {
  const long long pairs_per_image = hw >> 1;
  for (long long pair = lo_pair + (long long)blockIdx.x * kPeerThreads + threadIdx.x; pair < hi_pair; pair += (long long)gridDim.x * kPeerThreads) {
    const long long img = pair / pairs_per_image;
    const long long px = (pair - img * pairs_per_image) * 2;
    const size_t so = ((size_t)img * planes * hw + px) * sizeof(float);
    float2 acc[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) acc[k] = make_float2(0.f, 0.f);
    for (int r = 0; r < n_ranks; ++r) {
      const unsigned char* s = regions.base[r] + off_sums + so;
#pragma unroll
      for (int k = 0; k < 5; ++k)
        if (k < planes) {
          const float2 v = *reinterpret_cast<const float2*>(s + (size_t)k * hw * sizeof(float));
          acc[k].x += v.x;
          acc[k].y += v.y;
        }
    }
    const float m0x = __fdiv_rn(acc[0].x, total_samples), m0y = __fdiv_rn(acc[0].y, total_samples);
    const float m1x = __fdiv_rn(acc[1].x, total_samples), m1y = __fdiv_rn(acc[1].y, total_samples);
    const float ex = -(plogp_c(m0x) + plogp_c(m1x)), ey = -(plogp_c(m0y) + plogp_c(m1y));
    float2 mi = make_float2(0.f, 0.f), var = make_float2(0.f, 0.f);
    int pl = 2;
    if (has_mi) {
      mi = make_float2(ex - __fdiv_rn(acc[pl].x, total_samples), ey - __fdiv_rn(acc[pl].y, total_samples));
      ++pl;
    }
    if (has_var) {
      // unbiased variance from the raw second moments, in double: (sum p^2 - T m^2) cancels badly in float for confident
      // pixels; clamped at 0 like torch.var never goes below
      const double T = (double)total_samples, dn = T - 1.0;
      const double sx0 = acc[0].x, sx1 = acc[1].x, sy0 = acc[0].y, sy1 = acc[1].y;
      const double vx = 0.5 * ((fmax((double)acc[pl].x - sx0 * sx0 / T, 0.0) + fmax((double)acc[pl + 1].x - sx1 * sx1 / T, 0.0)) / dn);
      const double vy = 0.5 * ((fmax((double)acc[pl].y - sy0 * sy0 / T, 0.0) + fmax((double)acc[pl + 1].y - sy1 * sy1 / T, 0.0)) / dn);
      var = make_float2((float)vx, (float)vy);
    }
    const size_t o1 = ((size_t)img * hw + px) * sizeof(float);
    const size_t o2 = ((size_t)img * 2 * hw + px) * sizeof(float);
    uchar2 pr;
    pr.x = m1x > m0x ? 1 : 0;
    pr.y = m1y > m0y ? 1 : 0;
    for (int r = 0; r < n_ranks; ++r) {
      unsigned char* b = regions.base[r];
      *reinterpret_cast<float2*>(b + off_mean + o2) = make_float2(m0x, m0y);
      *reinterpret_cast<float2*>(b + off_mean + o2 + (size_t)hw * sizeof(float)) = make_float2(m1x, m1y);
      *reinterpret_cast<float2*>(b + off_entropy + o1) = make_float2(ex, ey);
      if (has_mi) *reinterpret_cast<float2*>(b + off_mi + o1) = mi;
      if (has_var) *reinterpret_cast<float2*>(b + off_var + o1) = var;
      *reinterpret_cast<float2*>(b + off_fg + o1) = make_float2(m1x, m1y);
      *reinterpret_cast<uchar2*>(b + off_pred + (size_t)img * hw + px) = pr;
    }
  }
}

}  // namespace rcu

extern "C" int rcu_peer_layout(int64_t n_images, int64_t hw, int has_mi, int has_var, rcu_peer_layout_t* out) {
  RCU_CHECK_ARG(out != nullptr && n_images >= 0 && hw >= 0 && hw % 2 == 0, "bad arguments (hw must be even)");
  const size_t px = (size_t)n_images * (size_t)hw;
  const int planes = 2 + (has_mi ? 1 : 0) + (has_var ? 2 : 0);
  auto up = [](size_t v) { return (v + 255) & ~size_t(255); };
  size_t off = 4096;                       // signal rows first: [row][rank] 32-bit words
  out->planes = planes;
  out->off_sums = off;      off += up(px * planes * sizeof(float));
  out->off_mean = off;      off += up(px * 2 * sizeof(float));
  out->off_entropy = off;   off += up(px * sizeof(float));
  out->off_mi = off;        off += has_mi ? up(px * sizeof(float)) : 0;
  out->off_var = off;       off += has_var ? up(px * sizeof(float)) : 0;
  out->off_foreground = off; off += up(px * sizeof(float));
  out->off_prediction = off; off += up(px);
  out->bytes = off;
  return RCU_OK;
}

extern "C" int rcu_aggregate_finish_peer(void* const* region_ptrs, int n_ranks, int rank, uint32_t epoch, int total_samples,
                                         int64_t n_images, int64_t hw, int has_mi, int has_var, void* stream) {
  RCU_CHECK_ARG(region_ptrs != nullptr, "NULL argument");
  RCU_CHECK_ARG(n_ranks >= 1 && n_ranks <= kPeerMaxRanks && rank >= 0 && rank < n_ranks, "bad rank %d / %d (at most %d ranks)", rank, n_ranks, kPeerMaxRanks);
  RCU_CHECK_ARG(total_samples >= 1 && (!has_var || total_samples >= 2), "bad total_samples");
  RCU_CHECK_ARG(epoch != 0, "epoch must be non-zero and increase from call to call");
  rcu_peer_layout_t lay;
  int rc = rcu_peer_layout(n_images, hw, has_mi, has_var, &lay);
  if (rc) return rc;
  PeerPtrs regions;
  for (int r = 0; r < kPeerMaxRanks; ++r) regions.base[r] = r < n_ranks ? reinterpret_cast<unsigned char*>(region_ptrs[r]) : nullptr;
  for (int r = 0; r < n_ranks; ++r) RCU_CHECK_ARG(regions.base[r] != nullptr, "region %d is NULL", r);
  cudaStream_t st = (cudaStream_t)stream;
  const long long total_pairs = (long long)n_images * (hw / 2);
  const long long base = total_pairs / n_ranks, extra = total_pairs % n_ranks;
  const long long lo = rank * base + (rank < extra ? rank : extra);
  const long long hi = lo + base + (rank < extra ? 1 : 0);
  // every rank's partial sums are complete and visible before anybody reads them
  peer_barrier_kernel<<<1, 32, 0, st>>>(regions, n_ranks, rank, 0, epoch);
  RCU_LAUNCH_CHECK();
  if (hi > lo) {
    long long blocks = (hi - lo + kPeerThreads - 1) / kPeerThreads;
    const long long cap = (long long)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    aggregate_finish_peer_kernel<<<(unsigned)blocks, kPeerThreads, 0, st>>>(regions, n_ranks, lo, hi, (long long)hw, lay.planes, has_mi, has_var,
                                                                            (float)total_samples, lay.off_sums, lay.off_mean, lay.off_entropy,
                                                                            lay.off_mi, lay.off_var, lay.off_foreground, lay.off_prediction);
    RCU_LAUNCH_CHECK();
  }
  // every rank's slice of the outputs has landed in this rank's region (and nobody still reads this rank's sums)
  peer_barrier_kernel<<<1, 32, 0, st>>>(regions, n_ranks, rank, 1, epoch);
  RCU_LAUNCH_CHECK();
  return RCU_OK;
}
