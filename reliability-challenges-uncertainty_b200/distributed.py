"""Multi-GPU partitioning of the hot path: one process per GPU (torch.distributed, NCCL over NVLink on the GPU box,
gloo in the CPU tests).  The reference has no distributed code (its only multi-GPU mechanism is nn.DataParallel in
training, common/trainloop/context.py:223-233); the axes below are the ones SURVEY.md §8(e) identifies.

  subjects / slices   independent units -> contiguous shards, NO data-path collective; only the tiny per-subject
                      count tables are summed when a subject's slices span ranks
  MC samples          each rank runs a range of Philox sample ids over all slices, accumulates sum_t softmax in
                      fp32 (rcu_aggregate_partial), ONE all-reduce of the (N, K, H, W) sums, then entropy / mean on
                      every rank (rcu_aggregate_finish)
  ensemble members    same, with a member range per rank

Float caveat: the all-reduce changes the fp32 summation order of the probability sums, so means agree with the
single-GPU result to ~1 ulp, not bit-for-bit; integer tables stay exact for identical probability maps.
"""
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


def aggregate_partial(logits, want_mi=False, want_var=False):
    """sum_t softmax (and optional sum_t H(p_t), sum_t p^2) of interleaved logits (T, N, H, W, 2) -> (N, K, H, W)."""
    t, n, h, w, _ = logits.shape
    sums = torch.empty((n, _partial_planes(want_mi, want_var), h, w), dtype=torch.float32, device=logits.device)
    with torch.cuda.device(logits.device):
        _lib.check(_lib.lib().rcu_aggregate_partial(_lib.ptr(logits), 0, t, n, h * w, int(want_mi), int(want_var), _lib.ptr(sums),
                                                    _lib.current_stream()))
    return sums


def aggregate_finish(sums, total_samples, has_mi=False, has_var=False, emit_prediction=False):
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
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().rcu_aggregate_finish(_lib.ptr(sums), int(total_samples), n, h * w, int(has_mi), int(has_var),
                                                   _lib.ptr(out['probabilities']), _lib.ptr(out['entropy']), _lib.ptr(out.get('mutual_info')),
                                                   _lib.ptr(out.get('variance')), _lib.ptr(out.get('prediction')), None, _lib.current_stream()))
    return out


def mc_predict_sample_sharded(engine, images, mc_steps, group=None, seed=None, slice_index0=0, want_mi=False, want_var=False,
                              emit_prediction=False):
    """MC dropout with the T samples split over the ranks of `group` (every rank holds all slices of the batch).

    Sample ids are global (rank r runs Philox samples [lo, hi)), so the union over ranks is exactly the single-GPU
    sample set.  Returns the MultiPredictionSummary outputs (identical on every rank) plus 'ws_probabilities' on
    rank 0 (the deterministic weight-scaling pass of McPredictStep is run once, not per rank)."""
    from .steps import softmax_planar
    world, rank = _world(group)
    lo, hi = shard_bounds(mc_steps, world, rank)
    det = rank == 0
    n_local = (hi - lo) + (1 if det else 0)
    out_ws = None
    n, _, h, w = images.shape
    if n_local > 0:
        logits = engine.forward_samples(images, n_local, dropout_mode=1, det_first=det, seed=seed, slice_index0=slice_index0, sample0=lo)
        if det:
            out_ws = softmax_planar(logits[0])
            logits = logits[1:]
    if hi > lo:
        sums = aggregate_partial(logits, want_mi, want_var)
    else:
        sums = torch.zeros((n, _partial_planes(want_mi, want_var), h, w), dtype=torch.float32, device=engine.device)
    allreduce_sum_(sums, group)
    out = aggregate_finish(sums, mc_steps, want_mi, want_var, emit_prediction)
    if out_ws is not None:
        out['ws_probabilities'] = out_ws
    return out


def ensemble_member_sharded(local_engines, n_members, images, group=None, emit_prediction=False):
    """Ensemble with the members split over ranks: `local_engines` are this rank's members (shard_bounds order)."""
    world, rank = _world(group)
    n, _, h, w = images.shape
    dev = local_engines[0].device if local_engines else torch.device('cuda', torch.cuda.current_device())
    if local_engines:
        logits = torch.stack([e.forward_samples(images, 1, dropout_mode=0)[0] for e in local_engines])
        sums = aggregate_partial(logits)
    else:
        sums = torch.zeros((n, 2, h, w), dtype=torch.float32, device=dev)
    allreduce_sum_(sums, group)
    return aggregate_finish(sums, n_members, emit_prediction=emit_prediction)
