"""Multi-GPU partitioning of the hot path: one process per GPU (torch.distributed, NCCL over NVLink on the GPU box,
gloo in the CPU tests).  The reference has no distributed code (its only multi-GPU mechanism is nn.DataParallel in
training, common/trainloop/context.py:223-233); the axes below are the ones SURVEY.md §8(e) identifies.

  subjects / slices   independent units -> contiguous shards, NO data-path collective; only the tiny per-subject
                      count tables are summed when a subject's slices span ranks
  MC samples          each rank runs a range of Philox sample ids over all slices, accumulates sum_t softmax in
                      fp32 (rcu_aggregate_partial), ONE all-reduce of the (N, K, H, W) sums, then entropy / mean on
                      every rank (rcu_aggregate_finish)
  ensemble members    same, with a member range per rank

Float caveat: the exchange changes the fp32 summation order of the probability sums, so means agree with the single-GPU
result to ~1 ulp, not bit-for-bit; integer tables stay exact for identical probability maps.  The sharded variance comes
from raw second moments (sum p^2 - T m^2, evaluated in float64 and clamped at 0): for confident pixels it agrees with the
single-GPU Welford result to ~1e-6 absolute, not to an ulp (tolerance written in tests/test_gpu_multi.py).

The exchange step itself has three routes: PeerExchange (fused reduce + finish over CUDA-IPC peer memory, the product
path on a B200 box), Comm (the library's own NCCL all-reduce) and torch.distributed (what the gloo CPU tests exercise).
"""
import ctypes

import numpy as np
import torch
import torch.distributed as dist

from . import _lib


def shard_bounds(n_items, world_size, rank):
    """Contiguous balanced shard [lo, hi) of `n_items` for `rank` (first n % world ranks get one extra)."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError('bad rank {} / world size {}'.format(rank, world_size))
    base, extra = divmod(int(n_items), world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_sizes(n_items, world_size):
    return [shard_bounds(n_items, world_size, r)[1] - shard_bounds(n_items, world_size, r)[0] for r in range(world_size)]


def _world(group=None):
    if not (dist.is_available() and dist.is_initialized()):
        return 1, 0
    return dist.get_world_size(group), dist.get_rank(group)


def allreduce_sum_(tensor, group=None):
    """In-place sum over ranks (no-op for a single process)."""
    world, _ = _world(group)
    if world > 1:
        dist.all_reduce(tensor, op=dist.ReduceOp.SUM, group=group)
    return tensor


def allreduce_metric_tables_(count, positives, conf_sum, ue_counts, group=None):
    """Sum the per-subject ECE / U-E tables over ranks: the integer tables travel as one int64 buffer (exact),
    the float64 confidence sums as a second one.  All arguments are tensors on the communication device and are
    updated in place; returns them."""
    world, _ = _world(group)
    if world == 1:
        return count, positives, conf_sum, ue_counts
    ints = torch.cat([count.reshape(-1), positives.reshape(-1), ue_counts.reshape(-1)]).to(torch.int64)
    dist.all_reduce(ints, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(conf_sum, op=dist.ReduceOp.SUM, group=group)
    a, b = count.numel(), count.numel() + positives.numel()
    count.copy_(ints[:a].view_as(count))
    positives.copy_(ints[a:b].view_as(positives))
    ue_counts.copy_(ints[b:].view_as(ue_counts))
    return count, positives, conf_sum, ue_counts


class Comm:
    """An NCCL communicator owned by the C-ABI (rcu_comm_*): the id is made on rank 0 and travels through torch.distributed
    (any backend), the collectives themselves are the library's — rcu_allreduce_probsum / rcu_allreduce_counts on the
    caller's CUDA stream."""

    def __init__(self, group=None, device=None):
        self.world, self.rank = _world(group)
        self.device = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
        ident = [None]
        if self.rank == 0:
            buf = (ctypes.c_ubyte * _lib.RCU_COMM_ID_BYTES)()
            _lib.check(_lib.lib().rcu_comm_unique_id(buf))
            ident[0] = bytes(buf)
        if self.world > 1:
            dist.broadcast_object_list(ident, src=0, group=group)
        raw = (ctypes.c_ubyte * _lib.RCU_COMM_ID_BYTES).from_buffer_copy(ident[0])
        handle = ctypes.c_void_p()
        _lib.check(_lib.lib().rcu_comm_create(raw, self.world, self.rank, int(self.device.index or 0), ctypes.byref(handle)))
        self._handle = handle

    def allreduce_probsum_(self, sums):
        """In-place fp32 sum over ranks of the (N, K, H, W) probability sums."""
        if not (sums.is_cuda and sums.dtype == torch.float32 and sums.is_contiguous()):
            raise ValueError('sums must be a contiguous float32 CUDA tensor')
        with torch.cuda.device(sums.device):
            _lib.check(_lib.lib().rcu_allreduce_probsum(self._handle, _lib.ptr(sums), sums.numel(), _lib.current_stream()))
        return sums

    def allreduce_metric_tables_(self, counts, conf_sum):
        """In-place sum over ranks of a flat int64 / uint64 count buffer and the float64 confidence sums (one grouped call)."""
        if not (counts.is_cuda and counts.dtype == torch.int64 and counts.is_contiguous() and conf_sum.is_cuda and
                conf_sum.dtype == torch.float64 and conf_sum.is_contiguous()):
            raise ValueError('counts must be contiguous int64 and conf_sum contiguous float64 CUDA tensors')
        with torch.cuda.device(counts.device):
            _lib.check(_lib.lib().rcu_allreduce_counts(self._handle, _lib.ptr(counts), counts.numel(), _lib.ptr(conf_sum), conf_sum.numel(),
                                                       _lib.current_stream()))
        return counts, conf_sum

    def close(self):
        if getattr(self, '_handle', None):
            _lib.lib().rcu_comm_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # interpreter shutdown
            pass


class _RegionView:
    """CUDA array interface over a slice of an exchange region (zero-copy torch.as_tensor)."""

    def __init__(self, ptr, shape, typestr, owner):
        self.__cuda_array_interface__ = {'shape': tuple(int(x) for x in shape), 'typestr': typestr, 'data': (int(ptr), False), 'version': 3,
                                         'strides': None}
        self._owner = owner


class PeerExchange:
    """The exchange regions behind rcu_aggregate_finish_peer for one problem size (n_images, H, W[, mi, var]) on the ranks
    of one box: this rank's region (where rcu_aggregate_partial writes its sums and where the finished outputs land) and
    the CUDA-IPC mappings of every other rank's region.  Handles travel through torch.distributed."""

    def __init__(self, n_images, h, w, want_mi=False, want_var=False, group=None, device=None):
        self.world, self.rank = _world(group)
        self.device = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
        self.n, self.h, self.w, self.mi, self.var = int(n_images), int(h), int(w), bool(want_mi), bool(want_var)
        self.layout = _lib.RcuPeerLayout()
        _lib.check(_lib.lib().rcu_peer_layout(self.n, self.h * self.w, int(self.mi), int(self.var), ctypes.byref(self.layout)))
        dev = int(self.device.index or 0)
        own = ctypes.c_void_p()
        hbuf = (ctypes.c_ubyte * _lib.RCU_IPC_HANDLE_BYTES)()
        _lib.check(_lib.lib().rcu_peer_region_alloc(self.layout.bytes, dev, ctypes.byref(own), hbuf))
        self._own = own
        handles = [None] * self.world
        if self.world > 1:
            dist.all_gather_object(handles, bytes(hbuf), group=group)
        self._ptrs = (ctypes.c_void_p * self.world)()
        self._opened = []
        for r in range(self.world):
            if r == self.rank:
                self._ptrs[r] = own.value
            else:
                mapped = ctypes.c_void_p()
                raw = (ctypes.c_ubyte * _lib.RCU_IPC_HANDLE_BYTES).from_buffer_copy(handles[r])
                _lib.check(_lib.lib().rcu_peer_region_open(raw, dev, ctypes.byref(mapped)))
                self._ptrs[r] = mapped.value
                self._opened.append(mapped)
        self.epoch = 0
        if self.world > 1:
            dist.barrier(group=group)   # every region is mapped (and zeroed) before the first collective call

    def _view(self, offset, shape, dtype):
        typestr = {torch.float32: '<f4', torch.uint8: '|u1'}[dtype]
        return torch.as_tensor(_RegionView(self._own.value + offset, shape, typestr, self), device=self.device)

    @property
    def sums(self):
        """(N, K, H, W) float32 view of this rank's partial-sum planes: the `sums` target of aggregate_partial."""
        return self._view(self.layout.off_sums, (self.n, self.layout.planes, self.h, self.w), torch.float32)

    def finish(self, total_samples, emit_prediction=False, emit_foreground=False):
        """Collective: barrier, fused peer reduce + finish, barrier.  Returns views of this rank's output planes (valid until
        the next finish on this exchange)."""
        self.epoch += 1
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().rcu_aggregate_finish_peer(self._ptrs, self.world, self.rank, self.epoch, int(total_samples), self.n,
                                                            self.h * self.w, int(self.mi), int(self.var), _lib.current_stream()))
        lay, shp1 = self.layout, (self.n, 1, self.h, self.w)
        out = {'probabilities': self._view(lay.off_mean, (self.n, 2, self.h, self.w), torch.float32),
               'entropy': self._view(lay.off_entropy, shp1, torch.float32)}
        if self.mi:
            out['mutual_info'] = self._view(lay.off_mi, shp1, torch.float32)
        if self.var:
            out['variance'] = self._view(lay.off_var, shp1, torch.float32)
        if emit_prediction:
            out['prediction'] = self._view(lay.off_prediction, (self.n, self.h, self.w), torch.uint8)
        if emit_foreground:
            out['foreground'] = self._view(lay.off_foreground, (self.n, self.h, self.w), torch.float32)
        return out

    def close(self):
        if getattr(self, '_own', None) is None:
            return
        torch.cuda.synchronize(self.device)
        for m in self._opened:
            _lib.lib().rcu_peer_region_close(m)
        self._opened = []
        _lib.lib().rcu_peer_region_free(self._own)
        self._own = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # interpreter shutdown
            pass


def gather_rows(rows, group=None):
    """All ranks receive the concatenation (in rank order) of every rank's list of picklable per-subject rows."""
    world, _ = _world(group)
    if world == 1:
        return list(rows)
    out = [None] * world
    dist.all_gather_object(out, list(rows), group=group)
    return [r for part in out for r in part]


def _partial_planes(want_mi, want_var):
    return 2 + (1 if want_mi else 0) + (2 if want_var else 0)


def aggregate_partial(logits, want_mi=False, want_var=False, out=None):
    """sum_t softmax (and optional sum_t H(p_t), sum_t p^2) of interleaved logits (T, N, H, W, 2) -> (N, K, H, W).
    `out`: write into this tensor (e.g. PeerExchange.sums) instead of a fresh one."""
    t, n, h, w, _ = logits.shape
    shape = (n, _partial_planes(want_mi, want_var), h, w)
    if out is not None and (tuple(out.shape) != shape or out.dtype != torch.float32 or not out.is_contiguous()):
        raise ValueError('out must be a contiguous float32 tensor of shape {}'.format(shape))
    sums = torch.empty(shape, dtype=torch.float32, device=logits.device) if out is None else out
    with torch.cuda.device(logits.device):
        _lib.check(_lib.lib().rcu_aggregate_partial(_lib.ptr(logits), 0, t, n, h * w, int(want_mi), int(want_var), _lib.ptr(sums),
                                                    _lib.current_stream()))
    return sums


def aggregate_finish(sums, total_samples, has_mi=False, has_var=False, emit_prediction=False, emit_foreground=False):
    """Turn (all-reduced) sums into the MultiPredictionSummary outputs."""
    n, k, h, w = sums.shape
    dev = sums.device
    out = {'probabilities': torch.empty((n, 2, h, w), dtype=torch.float32, device=dev),
           'entropy': torch.empty((n, 1, h, w), dtype=torch.float32, device=dev)}
    if has_mi:
        out['mutual_info'] = torch.empty((n, 1, h, w), dtype=torch.float32, device=dev)
    if has_var:
        out['variance'] = torch.empty((n, 1, h, w), dtype=torch.float32, device=dev)
    if emit_prediction:
        out['prediction'] = torch.empty((n, h, w), dtype=torch.uint8, device=dev)
    if emit_foreground:
        out['foreground'] = torch.empty((n, h, w), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().rcu_aggregate_finish(_lib.ptr(sums), int(total_samples), n, h * w, int(has_mi), int(has_var),
                                                   _lib.ptr(out['probabilities']), _lib.ptr(out['entropy']), _lib.ptr(out.get('mutual_info')),
                                                   _lib.ptr(out.get('variance')), _lib.ptr(out.get('prediction')), _lib.ptr(out.get('foreground')),
                                                   _lib.current_stream()))
    return out


def _reduce_and_finish(sums, total, want_mi, want_var, emit_prediction, emit_foreground, group, comm, exchange):
    """The exchange step: fused peer-memory reduce + finish (`exchange`, sums already sit in its region), the library's
    NCCL all-reduce (`comm`) or torch.distributed's (gloo in the CPU tests), then the finishing pass."""
    if exchange is not None:
        return exchange.finish(total, emit_prediction, emit_foreground)
    if comm is not None:
        comm.allreduce_probsum_(sums)
    else:
        allreduce_sum_(sums, group)
    return aggregate_finish(sums, total, want_mi, want_var, emit_prediction, emit_foreground)


def mc_predict_sample_sharded(engine, images, mc_steps, group=None, seed=None, slice_index0=0, want_mi=False, want_var=False,
                              emit_prediction=False, emit_foreground=False, comm=None, exchange=None):
    """MC dropout with the T samples split over the ranks of `group` (every rank holds all slices of the batch).

    Sample ids are global (rank r runs Philox samples [lo, hi)), so the union over ranks is exactly the single-GPU
    sample set.  Returns the MultiPredictionSummary outputs (identical on every rank) plus 'ws_probabilities' on
    rank 0 (the deterministic weight-scaling pass of McPredictStep is run once, not per rank).
    Exchange step: `exchange` (a PeerExchange of matching size) -> fused reduce + finish over NVLink peer memory;
    `comm` (a Comm) -> the library's NCCL all-reduce; neither -> torch.distributed.all_reduce."""
    from .steps import softmax_planar
    world, rank = _world(group)
    lo, hi = shard_bounds(mc_steps, world, rank)
    det = rank == 0
    n_local = (hi - lo) + (1 if det else 0)
    out_ws = None
    n, _, h, w = images.shape
    target = exchange.sums if exchange is not None else None
    if n_local > 0:
        logits = engine.forward_samples(images, n_local, dropout_mode=1, det_first=det, seed=seed, slice_index0=slice_index0, sample0=lo)
        if det:
            out_ws = softmax_planar(logits[0])
            logits = logits[1:]
    if hi > lo:
        sums = aggregate_partial(logits, want_mi, want_var, out=target)
    elif target is not None:
        sums = target.zero_()
    else:
        sums = torch.zeros((n, _partial_planes(want_mi, want_var), h, w), dtype=torch.float32, device=engine.device)
    out = _reduce_and_finish(sums, mc_steps, want_mi, want_var, emit_prediction, emit_foreground, group, comm, exchange)
    if out_ws is not None:
        out['ws_probabilities'] = out_ws
    return out


def ensemble_member_sharded(local_engines, n_members, images, group=None, emit_prediction=False, emit_foreground=False, comm=None,
                            exchange=None):
    """Ensemble with the members split over ranks: `local_engines` are this rank's members (shard_bounds order).
    Exchange step as in mc_predict_sample_sharded."""
    world, rank = _world(group)
    n, _, h, w = images.shape
    dev = local_engines[0].device if local_engines else torch.device('cuda', torch.cuda.current_device())
    target = exchange.sums if exchange is not None else None
    if local_engines:
        logits = torch.stack([e.forward_samples(images, 1, dropout_mode=0)[0] for e in local_engines])
        sums = aggregate_partial(logits, out=target)
    elif target is not None:
        sums = target.zero_()
    else:
        sums = torch.zeros((n, 2, h, w), dtype=torch.float32, device=dev)
    return _reduce_and_finish(sums, n_members, False, False, emit_prediction, emit_foreground, group, comm, exchange)
