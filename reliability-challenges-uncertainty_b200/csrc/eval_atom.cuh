// Fused calibration + U-E histogram pass, shared-atomic variant (round 2) — included by metrics.cu inside namespace rcu.
//
// Reference arithmetic (relative to the reference root): common/evalutation/numpyfunctions.py:51-69 (_binary_calibration:
// np.digitize against the linspace edges + three np.bincount) and :86-107 (uncertainty(): tp/tn/fp/fn and their intersection
// with u > th), swept over the thresholds of bin-eval/eval_uncertainty.py:195-202.
//
// The private-column kernels (eval_fused_lut_kernel / eval_fused_cell_kernel) are bound by instruction issue and by
// dependent shared-memory read-modify-write chains (76 instructions per voxel-warp, counters that scale with the thread
// count and cap the occupancy at 16 warps per SM).  Here
//   * every table entry is a function of ONE joint cell (segment of p) x (mask, target, prediction), and a cell is three
//     32-bit words updated by fire-and-forget shared-memory atomics (ATOMS, no return value) in a [cell][lane] table:
//     bank == lane, so a warp's 32 updates never conflict, there are no read-modify-write chains to wait for, and the
//     table does not grow with the block size (one block of 32 warps per SM);
//   * the segment of p comes from ONE 8-byte table entry {cell base, break point} indexed by the top mantissa bits of
//     saturate(p) + 1.0f (a monotone map, so "entries in lower buckets" / "entries in this bucket" split the sorted list
//     exactly; negative values and NaN land in bucket 0 and fail its p >= 0 compare, values above 1 land in the top
//     bucket) and one exact compare against the original p; buckets holding two or more list entries (the break points of
//     the float32 u(p) arithmetic cluster a few ulps apart) are flagged and scan their (two or three) entries;
//   * the float64 confidence sums of bins 1.. are EXACT integers: p >= edge[1] >= 2^-K makes q = p * 2^(23+K) an integer
//     below 2^32; the cell's second and third word accumulate sum(q) mod 2^32 and sum(q >> 16), which determine sum(q)
//     exactly while a column sees fewer than 65536 values; integers below 2^53 add without rounding in float64, so those
//     sums do not depend on the launch shape at all.  Every voxel adds into its cell unconditionally — no predicates on
//     mask or bin — and the fold reads only the cells it needs.  Bin 0 (arbitrarily small p) keeps a float64 register
//     accumulator per thread, reduced in a fixed order;
//   * count / positives / U-E rows / invalid are folded out of the cell totals at block end (integer, exact) and added to a
//     per-subject accumulator table in the workspace with 64-bit global reductions (integers: order-free); only the bin-0
//     float64 sum goes through per-block partials that the subject's last block (one ticket) adds in block order.  The
//     accumulators are zero at rest: the last block copies them out and clears them.
//   * two groups of four voxels per thread are in flight in registers; each set is reloaded as soon as its voxels are
//     counted, and the first loads leave before the tables are built.  (Tried and dropped, measured on B200: 768 / 512
//     threads with three or four sets in flight — fewer warps cost more than deeper prefetch gains; staging the inputs in
//     shared memory with 1-D bulk copies three tiles deep — 19 % more instructions for the barriers and staging reads,
//     0.55 instead of 0.6 of the HBM rate.)
// conf_sum of the overflow slot (values outside every bin) is 0 here: a sum over NaN / out-of-range values means nothing.
#pragma once

constexpr int kAtomThreads = 1024;                        // one block of 32 warps per SM
constexpr int kAtomMaxLutBits = 10;                      // at most 1025 buckets: two per thread
constexpr int kAtomSegBytes = 3072;                       // [word: count, W, H][code][lane] x 4 B
constexpr int kAtomMaxSegs = 68;
constexpr long long kAtomMaxVoxelsPerBlock = 1ll << 20;   // < 65536 values per (cell, lane) column with margin

struct AtomParams {
  float list[kBreakPad];             // 0, inner edges, last edge, nextafter(1), U-E breaks: strictly increasing, +inf padded
  unsigned int attr[kBreakPad + 1];  // per segment s = #{list <= p} (s = 0: negative / NaN): bin | class << 8 | invalid << 16
  int n_list, top;
  int seg_sum_lo, seg_sum_hi;        // segments of calibration bins 1 .. n_bins-1 (integer sums); [1, seg_sum_lo) is bin 0
  unsigned char bin_seg[RCU_MAX_BINS + 2];   // first segment of bin k (bins are runs of segments); [n_bins] = seg_sum_hi
  float q_scale;                     // 2^(23 + K)
  double q_inv;                      // 2^-(23 + K)
  float edge1;                       // upper edge of bin 0
  unsigned int span_ulps;            // widest (last - first entry, in float32 ulps) + 1 over the buckets holding several entries
  int lut_bits;
};

#ifdef RCU_ATOM_TRACE   // debug builds only (build.build_variant): per-block timeline of the kernel's phases, read by tools/hist_trace.py
__device__ unsigned long long g_atom_trace[2048 * 8];
__device__ __forceinline__ unsigned long long atom_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define ATOM_TRACE(I) do { if (threadIdx.x == 0 && blockIdx.y == 0 && blockIdx.x < 2040) g_atom_trace[blockIdx.x * 8 + (I)] = atom_now(); } while (0)
#define ATOM_CLK(I) do { if (threadIdx.x == 0 && blockIdx.y == 0 && blockIdx.x == 0) g_atom_trace[2040 * 8 + (I)] = (unsigned long long)clock64(); } while (0)
#else
#define ATOM_TRACE(I) do { } while (0)
#define ATOM_CLK(I) do { } while (0)
#endif

__device__ __forceinline__ unsigned int atom_bucket(float p_sat, int sh) {
  // top mantissa bits of saturate(p) + 1.0f: 0 .. 2^bits
  return (__float_as_uint(p_sat + 1.0f) - 0x3f800000u) >> sh;
}

template <bool HAS_MASK>
__global__ void __launch_bounds__(kAtomThreads, 1)
eval_fused_atom_kernel(const float* __restrict__ p, const unsigned char* __restrict__ pred, const unsigned char* __restrict__ target,
                       const unsigned char* __restrict__ mask, long long voxels_per_subject, int blocks_per_subject, int n_bins,
                       int n_classes, const __grid_constant__ AtomParams ap, HistOut out, unsigned int* __restrict__ tickets,
                       unsigned long long* __restrict__ partials, unsigned long long* __restrict__ accs) {
  extern __shared__ __align__(16) unsigned char smem_atom[];   // cells [seg][count, W, H][code][lane] u32 | lut
  __shared__ float s_list[kBreakPad + 4];
  __shared__ unsigned int s_bidx[kBreakPad];
  __shared__ unsigned int s_attr[kBreakPad + 1];
  __shared__ unsigned int s_slot[kPartialSlots];                 // this block's integer table slots
  __shared__ unsigned long long s_segq[kAtomMaxSegs];            // this block's sum(q) per segment (masked-in codes)
  __shared__ double s_red[32];
  __shared__ int s_binseg[RCU_MAX_BINS + 2];
  __shared__ int s_is_last;
  const int tid = threadIdx.x;
  ATOM_TRACE(0);
  ATOM_CLK(0);
  const int warp = tid >> 5, lane = tid & 31;
  const int nb1 = n_bins + 1;
  const int ncls = n_classes;
  const int n_seg = ap.n_list + 1;
  const int sh = 23 - ap.lut_bits;
  const unsigned int n_buckets = (1u << ap.lut_bits) + 1u;     // saturate(p) = 1 is a bucket of its own
  unsigned int* s_cells = reinterpret_cast<unsigned int*>(smem_atom);
  uint2* s_lut = reinterpret_cast<uint2*>(smem_atom + (size_t)n_seg * kAtomSegBytes);

  const int subject = blockIdx.y;
  const long long groups = (voxels_per_subject + 3) >> 2;
  const long long gpb = (groups + blocks_per_subject - 1) / blocks_per_subject;
  const long long g0 = min(groups, (long long)blockIdx.x * gpb);
  const long long g1 = min(groups, g0 + gpb);
  const long long gend = min(g1, voxels_per_subject >> 2);     // whole groups of four voxels
  // this block's inputs, per thread: group g0 + tid + i * 1024 for i = 0 .. n_mine-1 (32-bit loop state from here on)
  const int n_block = (int)(gend - g0);
  const int n_mine = n_block > tid ? (n_block - tid + kAtomThreads - 1) / kAtomThreads : 0;
  const long long v0 = (long long)subject * voxels_per_subject + (g0 + tid) * 4;
  const float* p_t = p + v0;
  const unsigned char* t_t = target + v0;
  const unsigned char* d_t = pred + v0;
  const unsigned char* m_t = HAS_MASK ? mask + v0 : nullptr;

  // Two groups of four voxels per thread in flight.  The first loads leave before the tables are built (their latency hides
  // the prologue); afterwards each register set is reloaded as soon as its voxels are counted, so the loads of the next
  // step travel while the other set is being counted.  Loads and atomics are volatile asm: the compiler keeps their order.
  float4 pa = make_float4(0.f, 0.f, 0.f, 0.f), pb = pa;
  unsigned int ta = 0u, da = 0u, ma = 0u, tb = 0u, db = 0u, mb = 0u;
  auto load = [&](int i, float4& p4, unsigned int& t4, unsigned int& d4, unsigned int& m4) {
    if (i < n_mine) {
      const size_t o = (size_t)i * (kAtomThreads * 4);
      p4 = ld_stream_f4(p_t + o);
      t4 = ld_stream_u32(t_t + o);
      d4 = ld_stream_u32(d_t + o);
      if (HAS_MASK) m4 = ld_stream_u32(m_t + o);
    }
  };
  load(0, pa, ta, da, ma);
  load(1, pb, tb, db, mb);
  ATOM_CLK(1);

  {
    uint4* z = reinterpret_cast<uint4*>(smem_atom);
    for (int i = tid; i < n_seg * (kAtomSegBytes / 16); i += kAtomThreads) z[i] = make_uint4(0u, 0u, 0u, 0u);
  }
  // the tables out of the kernel parameters: per-thread indices into the constant bank serialise (32 replays per warp),
  // so three single threads in three different warps copy them with uniform reads while the other warps clear the cells
  for (int i = ap.n_list + tid; i < kBreakPad + 4; i += kAtomThreads) s_list[i] = __int_as_float(0x7f800000);
  for (int i = ap.n_list + tid; i < kBreakPad; i += kAtomThreads) s_bidx[i] = 0xffffffffu;
  if (tid == 32) {
    for (int i = 0; i < ap.n_list; ++i) { const float v = ap.list[i]; s_list[i] = v; s_bidx[i] = atom_bucket(__saturatef(v), sh); }
  } else if (tid == 64) {
    for (int i = 0; i < n_seg; ++i) s_attr[i] = ap.attr[i];
  } else if (tid == 96) {
    for (int i = 0; i < n_bins + 2; ++i) s_binseg[i] = ap.bin_seg[i];
  }
  for (int i = tid; i < kPartialSlots; i += kAtomThreads) s_slot[i] = 0u;
  ATOM_CLK(2);
  __syncthreads();
  ATOM_CLK(3);
  // bucket table: entries in lower buckets are <= every p of the bucket, entries of higher buckets above every p (the
  // bucket map is monotone), so a bucket needs #{entries below it} and its own entries.  One entry: {cell base, the entry};
  // several: {cell base | count, index of the first} — the flagged path scans them
  for (unsigned int b = tid; b < n_buckets; b += kAtomThreads) {
    int lo = 0, hi = 0;   // #{entries in lower buckets}, #{entries in buckets <= b}
#pragma unroll
    for (int step = kBreakPad / 2; step >= 1; step >>= 1) {
      if (s_bidx[lo + step - 1] < b) lo += step;
      if (s_bidx[hi + step - 1] <= b) hi += step;
    }
    const int inside = hi - lo;
    s_lut[b] = inside >= 2 ? make_uint2((unsigned int)lo * (unsigned int)kAtomSegBytes | (unsigned int)min(inside, 1023), (unsigned int)lo)
                           : make_uint2((unsigned int)lo * (unsigned int)kAtomSegBytes, inside == 1 ? __float_as_uint(s_list[lo]) : 0x7f800000u);
  }
  ATOM_CLK(4);
  __syncthreads();
  ATOM_TRACE(1);
  ATOM_CLK(5);

  const unsigned int cells_a = (unsigned int)__cvta_generic_to_shared(smem_atom) + lane * 4;   // + seg * 3072 + word * 1024 + code * 128
  const unsigned int lut_a = (unsigned int)__cvta_generic_to_shared(s_lut) - ((0x3f800000u >> sh) << 3);
  const int sh3 = sh - 3;
  const float q_scale = ap.q_scale, edge1 = ap.edge1;
  const unsigned int span_ulps = ap.span_ulps;
  double acc0 = 0.0;

  // one voxel whose cell offset is known: three atomics on the cell, bin 0 into the register accumulator.  `mbit` selects
  // the voxel's mask bit in the code word; ps = saturate(p) is 0 for negative values and NaN, so it can be added blindly.
  auto add_one = [&](float pv, float ps, unsigned int off, unsigned int code_off, unsigned int cw, unsigned int mbit) {
    const unsigned int a = cells_a + off + code_off;
    const unsigned int q = __float2uint_rz(pv * q_scale);   // exact inside the summed bins; anything elsewhere (never read)
    asm volatile(
        "red.shared.add.u32 [%0], 1;\n\t"
        "red.shared.add.u32 [%0 + 1024], %1;\n\t"
        "red.shared.add.u32 [%0 + 2048], %2;" ::"r"(a), "r"(q), "r"(q >> 16)
        : "memory");
    float x;   // ps inside the mask and inside bin 0, else +0
    asm("{\n\t.reg .pred pm, pb;\n\t.reg .b32 t;\n\t"
        "and.b32 t, %2, %3;\n\tsetp.ne.u32 pm, t, 0;\n\t"
        "setp.lt.and.f32 pb, %1, %4, pm;\n\t"
        "selp.f32 %0, %1, 0f00000000, pb;\n\t}"
        : "=f"(x)
        : "f"(ps), "r"(cw), "r"(mbit), "f"(edge1));
    acc0 += (double)x;
  };
  auto four = [&](const float4 pv, const unsigned int cw) {   // cw: prediction | target << 1 | mask << 2 per byte, bits normalised
    const float pq[4] = {pv.x, pv.y, pv.z, pv.w};
    float ps[4];
    unsigned int off[4], brk[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      ps[e] = __saturatef(pq[e]);
      const unsigned int la = lut_a + ((__float_as_uint(ps[e] + 1.0f) >> sh3) & 0xfffffff8u);
      asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(off[e]), "=r"(brk[e]) : "r"(la));
    }
    if ((off[0] | off[1] | off[2] | off[3]) & 1023u) {   // some bucket holds several entries
      // The entries of such a bucket are a cluster a few ulps wide (the float32 u(p) arithmetic is not monotone right at a
      // threshold): a value below the cluster or at / above its end skips all of it at once; only a value INSIDE the
      // cluster's span (span_ulps: the widest one, from the host) is compared with every entry.
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const unsigned int cnt = off[e] & 1023u;
        if (cnt) {
          const float first = s_list[brk[e]];
          off[e] -= cnt;
          const float pz = pq[e] > 0.0f ? pq[e] : 0.0f;   // -0, negative values and NaN sit at the bottom of bucket 0: always compared exactly
          if (__float_as_uint(pz) - __float_as_uint(first) < span_ulps) {
            for (unsigned int i = 0; i < cnt; ++i) off[e] += (pq[e] >= s_list[brk[e] + i]) ? (unsigned int)kAtomSegBytes : 0u;
          } else if (pq[e] >= first) {
            off[e] += cnt * (unsigned int)kAtomSegBytes;
          }
          brk[e] = 0x7f800000u;
        }
      }
    }
#pragma unroll
    for (int e = 0; e < 4; ++e)   // one segment up where p >= the bucket's break point
      asm("{\n\t.reg .pred pg;\n\tsetp.ge.f32 pg, %1, %2;\n\t@pg add.u32 %0, %0, %3;\n\t}"
          : "+r"(off[e])
          : "f"(pq[e]), "f"(__uint_as_float(brk[e])), "n"(kAtomSegBytes));
    const unsigned int c02 = (cw & 0x00070007u) << 7, c13 = ((cw >> 8) & 0x00070007u) << 7;   // code * 128 of voxels 0 | 2 and 1 | 3
    add_one(pq[0], ps[0], off[0], c02 & 0xffffu, cw, 0x00000004u);
    add_one(pq[1], ps[1], off[1], c13 & 0xffffu, cw, 0x00000400u);
    add_one(pq[2], ps[2], off[2], c02 >> 16, cw, 0x00040000u);
    add_one(pq[3], ps[3], off[3], c13 >> 16, cw, 0x04000000u);
  };
  auto codes = [&](unsigned int t4, unsigned int d4, unsigned int m4) {
    return swar_nonzero(d4) + 2u * swar_nonzero(t4) + (HAS_MASK ? 4u * swar_nonzero(m4) : 0x04040404u);
  };

  for (int i = 0; i < n_mine; i += 2) {
    four(pa, codes(ta, da, ma));
    load(i + 2, pa, ta, da, ma);
    if (i + 1 < n_mine) four(pb, codes(tb, db, mb));
    load(i + 3, pb, tb, db, mb);
  }
  if (tid == 0 && g0 < g1 && g1 == groups && gend * 4 < voxels_per_subject) {   // ragged last group of the subject
    const long long base_v = (long long)subject * voxels_per_subject;
    for (long long v = gend * 4; v < voxels_per_subject; ++v) {
      const long long a = base_v + v;
      const unsigned int code = (pred[a] != 0 ? 1u : 0u) | (target[a] != 0 ? 2u : 0u) | ((HAS_MASK ? mask[a] != 0 : true) ? 4u : 0u);
      const float pv = p[a];
      const unsigned int off = (pv >= 0.0f) ? (unsigned int)count_breaks<float, false>(pv, s_list, ap.top) * (unsigned int)kAtomSegBytes : 0u;   // false for NaN
      add_one(pv, __saturatef(pv), off, code * 128u, code, 4u);
    }
  }
  ATOM_CLK(6);
  __syncthreads();
  ATOM_TRACE(2);
  ATOM_CLK(7);

  // ---- block totals, one thread per cell: the 32 lane columns are read rotated by the thread index (bank == lane stays
  // conflict-free), every cell's count goes to the table slots it belongs to, the masked-in cells' exact sum(q) to their segment ----
  constexpr int kWarps = kAtomThreads / 32;
  const int n_slots = 3 * nb1 + 4 * ncls + 1;
  for (int c0 = 0; c0 < n_seg * 8; c0 += kAtomThreads) {
    const int c = c0 + tid;
    const bool valid = c < n_seg * 8;
    const int sg = valid ? c >> 3 : 0, code = c & 7;              // code = prediction | target << 1 | mask << 2
    const bool want_q = valid && code >= 4 && sg >= ap.seg_sum_lo && sg < ap.seg_sum_hi;
    const unsigned int* base = s_cells + sg * (kAtomSegBytes / 4) + code * 32;
    unsigned int cnt = 0u;
    unsigned long long sumq = 0ull;
    if (valid) {
#pragma unroll 8
      for (int l = 0; l < 32; ++l) cnt += base[(l + tid) & 31];
    }
    if (want_q) {
#pragma unroll 8
      for (int l = 0; l < 32; ++l) {
        const unsigned int w = base[256 + ((l + tid) & 31)], h = base[512 + ((l + tid) & 31)];
        sumq += ((unsigned long long)h << 16) + (unsigned long long)(w - (h << 16));   // sum(q) of the column
      }
    }
    const unsigned int at = s_attr[sg];
    const int k = (int)(at & 0xffu), j = (int)((at >> 8) & 0xffu);
    if (valid && cnt != 0u) {
      if (code & 4) {
        atomicAdd(&s_slot[k], cnt);
        if (code & 2) atomicAdd(&s_slot[nb1 + k], cnt);
      }
      const int r = (code & 3) == 3 ? 0 : ((code & 3) == 0 ? 1 : ((code & 3) == 1 ? 2 : 3));   // rows tp, tn, fp, fn
      atomicAdd(&s_slot[3 * nb1 + r * ncls + j], cnt);
      if ((at >> 16) & 1u) atomicAdd(&s_slot[n_slots - 1], cnt);
    }
    ATOM_CLK(8);
    sumq += __shfl_xor_sync(0xffffffffu, sumq, 1);    // codes 4..7 of a segment are four adjacent lanes
    sumq += __shfl_xor_sync(0xffffffffu, sumq, 2);
    if (valid && code == 4) s_segq[sg] = sumq;       // 0 outside the summed segments
  }
  {
    double v = acc0;
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (lane == 0) s_red[warp] = v;
  }
  ATOM_CLK(9);
  __syncthreads();
  ATOM_CLK(10);
  unsigned long long* acc = accs + (long long)subject * kPartialSlots;
  double* bin0_partials = reinterpret_cast<double*>(partials) + (long long)subject * blocks_per_subject;
  if (tid < n_slots) {
    if (tid >= 2 * nb1 && tid < 3 * nb1) {
      const int k = tid - 2 * nb1;
      if (k >= 1 && k < n_bins) {
        unsigned long long tot = 0ull;
        for (int sg = s_binseg[k]; sg < s_binseg[k + 1]; ++sg) tot += s_segq[sg];
        if (tot != 0ull) atomicAdd(&acc[tid], tot);
      }
    } else if (s_slot[tid] != 0u) {
      atomicAdd(&acc[tid], (unsigned long long)s_slot[tid]);
    }
  }
  if (warp == kWarps - 1) {   // the warps' bin-0 sums: a fixed tree again
    double v = lane < kWarps ? s_red[lane] : 0.0;
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (lane == 0) bin0_partials[blockIdx.x] = v;
  }
  ATOM_TRACE(3);
  ATOM_CLK(11);
  asm volatile("fence.acq_rel.gpu;" ::: "memory");   // release: reductions and the partial before the ticket (everything read back goes through L2)
  ATOM_CLK(12);
  __syncthreads();
  ATOM_TRACE(4);
  ATOM_CLK(13);
  unsigned int* tk = tickets + (long long)subject * kTicketsPerSubject;
  if (tid == 0) s_is_last = (atomicAdd(tk, 1u) == (unsigned int)blocks_per_subject - 1u);
  __syncthreads();
  ATOM_TRACE(5);
  ATOM_CLK(14);
  if (!s_is_last) return;
  asm volatile("fence.acq_rel.gpu;" ::: "memory");
  if (tid < n_slots) {
    const unsigned long long v = __ldcg(&acc[tid]);
    __stcg(&acc[tid], 0ull);   // zero at rest for the next stream-ordered call
    const int sl = tid;
    if (sl < nb1) out.count[(long long)subject * nb1 + sl] = v;
    else if (sl < 2 * nb1) out.positives[(long long)subject * nb1 + (sl - nb1)] = v;
    else if (sl < 3 * nb1) {
      if (sl != 2 * nb1) out.conf_sum[(long long)subject * nb1 + (sl - 2 * nb1)] = (double)v * ap.q_inv;   // exact below 2^53
    } else if (sl < n_slots - 1) out.ue_counts[(long long)subject * 4 * ncls + (sl - 3 * nb1)] = v;
    else if (out.invalid) out.invalid[subject] = v;
  }
  if (warp == kWarps - 1) {   // bin 0: the per-block float64 sums in block order (lane-strided runs, then a fixed tree)
    double v = 0.0;
    for (int b0 = 0; b0 < blocks_per_subject; b0 += 256) {   // eight independent loads per lane and trip
      double x[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) x[u] = b0 + 32 * u + lane < blocks_per_subject ? __ldcg(&bin0_partials[b0 + 32 * u + lane]) : 0.0;
#pragma unroll
      for (int u = 0; u < 8; ++u) v += x[u];
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (lane == 0) out.conf_sum[(long long)subject * nb1] = v;
  }
  if (tid == 0) *tk = 0u;
  ATOM_TRACE(6);
}
