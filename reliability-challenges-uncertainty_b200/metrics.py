"""Device entry points of the calibration / uncertainty-error metrics (thin wrappers over the C-ABI).

Inputs may be numpy arrays (the reference's EvaluationStrategy protocol hands numpy around) or CUDA tensors
(the in-memory pipeline: probabilities never leave HBM).  numpy inputs are copied to the current CUDA device;
results come back as small numpy integer / float64 tables.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from . import _torch_ext
from . import tables


def _device():
    if not torch.cuda.is_available():
        raise _lib.RcuError('rcu_b200 metrics need a CUDA device (there is no CPU fallback)')
    return torch.device('cuda', torch.cuda.current_device())


def _pick_device(*xs):
    """The device the call runs on: that of the first CUDA tensor among the inputs, else the current CUDA device (numpy inputs
    are copied there).  Every C call below is made under torch.cuda.device(that device), so tensors on cuda:1 work while
    cuda:0 is current, and the per-device workspace matches the data."""
    for x in xs:
        if torch.is_tensor(x) and x.is_cuda:
            return x.device
    return _device()


def _to_device(x, dtype, name, device=None):
    """Flat contiguous CUDA tensor of `dtype` for a numpy array / torch tensor (bool -> uint8)."""
    if x is None:
        return None
    if torch.is_tensor(x) and x.is_cuda and x.dtype == dtype and x.is_contiguous():   # in-memory pipeline: nothing to do
        return x if x.dim() == 1 else x.view(-1)
    if isinstance(x, np.ndarray):
        if x.dtype == np.bool_:
            x = x.view(np.uint8)
        t = torch.from_numpy(np.ascontiguousarray(x))
    elif torch.is_tensor(x):
        t = x
        if t.dtype == torch.bool:
            t = t.to(torch.uint8)
    else:
        raise ValueError("object of type '{}' must be '{}'".format(type(x).__name__, np.ndarray.__name__))
    if t.dtype != dtype:
        t = t.to(dtype)
    t = t.to(device if device is not None else _device(), non_blocking=True).contiguous().view(-1)
    return t


class _Workspace:
    """Scratch for the deterministic two-stage reductions, one per (device, stream): allocated once, reused.  The kernels keep
    tickets and partial tables in it, so two streams of one device (a background hook next to the main loop) must not share
    a buffer; calls on ONE stream are ordered by the stream itself."""
    _cache = {}

    @classmethod
    def get(cls, n_subjects):
        dev = torch.cuda.current_device()     # callers switch to the tensors' device first (torch.cuda.device(...))
        key = (dev, int(torch.cuda.current_stream().cuda_stream))
        need = int(_lib.lib().rcu_metrics_workspace_bytes(int(n_subjects)))
        buf = cls._cache.get(key)
        if buf is None or buf.numel() < need:
            buf = torch.empty(need, dtype=torch.uint8, device=_device())
            _lib.check(_lib.lib().rcu_metrics_workspace_init(_lib.ptr(buf), buf.numel(), _lib.current_stream()))
            cls._cache[key] = buf
        return buf


_EDGES = {}
_BREAKS = {}


def _edges_for(n_bins):
    """(float32 edge array, ctypes pointer) of the calibration bins, built once per n_bins."""
    e = _EDGES.get(n_bins)
    if e is None:
        e = _EDGES[n_bins] = _f32_array(tables.calibration_edges_f32(n_bins))
    return e


def _breaks_for(breaks, seg):
    """ctypes views of a break table, cached per table object (the hooks reuse one table for every subject)."""
    key = (id(breaks), id(seg))
    e = _BREAKS.get(key)
    if e is None or e[0] is not breaks or e[1] is not seg:
        if len(_BREAKS) > 64:
            _BREAKS.clear()
        b32, b32_p = _f32_array(breaks)
        seg8 = np.ascontiguousarray(seg, dtype=np.uint8)
        e = _BREAKS[key] = (breaks, seg, b32, b32_p, seg8, seg8.ctypes.data_as(_lib.c_uint8_p))
    return e[3], e[5]


_HOST_TABLES = {}


def _host_tables(n_bins, breaks, seg):
    """Host torch tensors (edges float32, breaks float32, seg uint8) for the torch-extension operators, cached like _breaks_for."""
    key = (n_bins, id(breaks), id(seg))
    e = _HOST_TABLES.get(key)
    if e is None or e[0] is not breaks or e[1] is not seg:
        if len(_HOST_TABLES) > 64:
            _HOST_TABLES.clear()
        e = _HOST_TABLES[key] = (breaks, seg, torch.from_numpy(np.ascontiguousarray(tables.calibration_edges_f32(n_bins), dtype=np.float32).copy()),
                                 torch.from_numpy(np.ascontiguousarray(breaks, dtype=np.float32).copy()),
                                 torch.from_numpy(np.ascontiguousarray(seg, dtype=np.uint8).copy()))
    return e[2], e[3], e[4]


def _f32_array(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(_lib.c_float_p)


def _check_lengths(n, vps, n_subjects, **arrays):
    for k, v in arrays.items():
        if v is not None and v.numel() != n:
            raise ValueError('"{}" has {} elements, expected {}'.format(k, v.numel(), n))
    if vps * n_subjects != n:
        raise ValueError('{} elements do not split into {} subjects'.format(n, n_subjects))


def calibration_tables(p, target, mask=None, n_bins=10, threshold_range=None, n_subjects=1, sync=True):
    """ECE reliability tables of `n_subjects` equally sized subjects laid out back to back.

    Returns (count int64[S, n_bins+1], positives int64[S, n_bins+1], conf_sum float64[S, n_bins+1]); the last
    slot counts values outside every bin (p < 0, p >= 1+1e-8, NaN).  With sync=False CUDA tensors are returned.
    """
    dev = _pick_device(p, target, mask)
    p_d = _to_device(p, torch.float32, 'probabilities', dev)
    t_d = _to_device(target, torch.uint8, 'target', dev)
    m_d = _to_device(mask, torch.uint8, 'mask', dev)
    n = p_d.numel()
    vps = n // max(n_subjects, 1)
    _check_lengths(n, vps, n_subjects, target=t_d, mask=m_d)
    count = torch.zeros((n_subjects, n_bins + 1), dtype=torch.int64, device=dev)
    positives = torch.zeros_like(count)
    conf = torch.zeros((n_subjects, n_bins + 1), dtype=torch.float64, device=dev)
    if n > 0:
        edges, edges_p = _f32_array(tables.calibration_edges_f32(n_bins))
        lo, hi = (float('nan'), float('nan')) if threshold_range is None else (float(threshold_range[0]), float(threshold_range[1]))
        with torch.cuda.device(dev):
            ws = _Workspace.get(n_subjects)
            _lib.check(_lib.lib().rcu_calib_hist(_lib.ptr(p_d), _lib.ptr(t_d), _lib.ptr(m_d), vps, n_subjects, edges_p, n_bins, lo, hi,
                                                 _lib.ptr(count), _lib.ptr(positives), _lib.ptr(conf), _lib.ptr(ws), ws.numel(),
                                                 _lib.current_stream()))
    if not sync:
        return count, positives, conf
    return count.cpu().numpy(), positives.cpu().numpy(), conf.cpu().numpy()


def _ue_tables_for(kind, thresholds):
    if kind == 'p':
        return tables.uncertainty_break_table(thresholds)
    if kind == 'u32':
        return tables.threshold_breaks_f32(thresholds)
    ths = np.asarray(thresholds, dtype=np.float64)
    order = np.argsort(ths, kind='stable')
    return ths[order], np.arange(len(ths) + 1, dtype=np.uint8), order


def ue_tables(values, prediction, target, thresholds=tables.SWEEP_THRESHOLDS, mask=None, kind='p', n_subjects=1, sync=True,
              break_table=None):
    """Joint (confusion class x thresholds-exceeded) histogram for all thresholds in one pass.

    kind 'p'   : `values` is the float32 foreground probability; uncertainty = H([1-p, p]) / ln 2 in the
                 reference's arithmetic (ToEntropy after AddBackgroundProbabilities), classified through an exact
                 float32 break table.
    kind 'u32' : `values` is a float32 uncertainty map;  kind 'u64': a float64 uncertainty map (ToEntropy output).
    Returns (table int64[S, 4, K+1] rows tp, tn, fp, fn; invalid int64[S]; order) — `order` maps the k-th smallest
    threshold back to the caller's index.
    """
    if kind not in ('p', 'u32', 'u64'):
        raise ValueError('unknown kind "{}"'.format(kind))
    breaks, seg, order = break_table if break_table is not None else _ue_tables_for(kind, thresholds)
    n_classes = len(order) + 1
    dev = _pick_device(values, prediction, target, mask)
    v_d = _to_device(values, torch.float64 if kind == 'u64' else torch.float32, 'values', dev)
    d_d = _to_device(prediction, torch.uint8, 'prediction', dev)
    t_d = _to_device(target, torch.uint8, 'target', dev)
    m_d = _to_device(mask, torch.uint8, 'mask', dev)
    n = v_d.numel()
    vps = n // max(n_subjects, 1)
    _check_lengths(n, vps, n_subjects, prediction=d_d, target=t_d, mask=m_d)
    table = torch.zeros((n_subjects, 4, n_classes), dtype=torch.int64, device=dev)
    invalid = torch.zeros((n_subjects,), dtype=torch.int64, device=dev)
    if n > 0:
        seg = np.ascontiguousarray(seg, dtype=np.uint8)
        seg_p = seg.ctypes.data_as(_lib.c_uint8_p)
        if kind == 'u64':
            b64 = np.ascontiguousarray(breaks, dtype=np.float64)
            b32_p, b64_p = None, b64.ctypes.data_as(_lib.c_double_p)
            vk = 2
        else:
            b32, b32_p = _f32_array(breaks)
            b64_p = None
            vk = 0 if kind == 'p' else 1
        with torch.cuda.device(dev):
            ws = _Workspace.get(n_subjects)
            _lib.check(_lib.lib().rcu_ue_hist(_lib.ptr(v_d), vk, _lib.ptr(d_d), _lib.ptr(t_d), _lib.ptr(m_d), vps, n_subjects, b32_p, b64_p,
                                              len(breaks), seg_p, n_classes, _lib.ptr(table), _lib.ptr(invalid), _lib.ptr(ws), ws.numel(),
                                              _lib.current_stream()))
    if not sync:
        return table, invalid, order
    return table.cpu().numpy(), invalid.cpu().numpy(), order


def eval_fused(p, prediction, target, mask=None, n_bins=10, thresholds=tables.SWEEP_THRESHOLDS, n_subjects=1, sync=True,
               break_table=None):
    """ECE tables (masked) and the U-E joint table (unmasked) in ONE pass over (p, prediction, target, mask).

    Returns (count, positives, conf_sum, ue_table, invalid, order) as in calibration_tables / ue_tables.
    """
    breaks, seg, order = break_table if break_table is not None else tables.uncertainty_break_table(thresholds)
    n_classes = len(order) + 1
    dev = _pick_device(p, prediction, target, mask)
    p_d = _to_device(p, torch.float32, 'probabilities', dev)
    d_d = _to_device(prediction, torch.uint8, 'prediction', dev)
    t_d = _to_device(target, torch.uint8, 'target', dev)
    m_d = _to_device(mask, torch.uint8, 'mask', dev)
    n = p_d.numel()
    vps = n // max(n_subjects, 1)
    _check_lengths(n, vps, n_subjects, prediction=d_d, target=t_d, mask=m_d)
    # one allocation for all five result tables (the kernel's last block writes every slot, nothing needs zeroing):
    # the single-subject call is otherwise bound by five tiny memset launches
    nb1 = n_bins + 1
    width = 3 * nb1 + 4 * n_classes + 1
    o1, o2, o3, o4 = n_subjects * nb1, 2 * n_subjects * nb1, 3 * n_subjects * nb1, 3 * n_subjects * nb1 + n_subjects * 4 * n_classes
    ext = _torch_ext.ops() if n > 0 else None
    if ext is not None:
        # torch-extension binding: one registered operator call (tensors in, the flat table out, current stream inside)
        with torch.cuda.device(dev):
            ws = _Workspace.get(n_subjects)
            edges_t, breaks_t, seg_t = _host_tables(n_bins, breaks, seg)
            flat = ext.eval_fused(p_d, d_d, t_d, m_d, edges_t, breaks_t, seg_t, n_subjects, n_classes, ws)
    else:
        flat = (torch.empty if n > 0 else torch.zeros)((n_subjects * width,), dtype=torch.int64, device=dev)
    if n > 0 and ext is None:
        _, edges_p = _edges_for(n_bins)
        b32_p, seg_p = _breaks_for(breaks, seg)
        with torch.cuda.device(dev):
            ws = _Workspace.get(n_subjects)
            base = flat.data_ptr()
            _lib.check(_lib.lib().rcu_eval_fused(_lib.ptr(p_d), _lib.ptr(d_d), _lib.ptr(t_d), _lib.ptr(m_d), vps, n_subjects, edges_p, n_bins,
                                                 b32_p, len(breaks), seg_p, n_classes, base, base + 8 * o1, base + 8 * o2, base + 8 * o3,
                                                 base + 8 * o4, _lib.ptr(ws), ws.numel(), _lib.current_stream()))
    if not sync:
        return (flat[:o1].view(n_subjects, nb1), flat[o1:o2].view(n_subjects, nb1), flat[o2:o3].view(torch.float64).view(n_subjects, nb1),
                flat[o3:o4].view(n_subjects, 4, n_classes), flat[o4:], order)
    host = flat.cpu().numpy()      # ONE device->host copy for all five tables
    return (host[:o1].reshape(n_subjects, nb1), host[o1:o2].reshape(n_subjects, nb1),
            host[o2:o3].view(np.float64).reshape(n_subjects, nb1), host[o3:o4].reshape(n_subjects, 4, n_classes), host[o4:], order)


def philox_keep_scale_host(seed, p_drop, site_channels, slice_index0, n_slices, sample0, n_samples):
    """Host copy of the engine's Dropout2d keep-scale stream (no GPU needed): float32[n_samples, n_slices, sum(C)]."""
    sc = (ctypes.c_int * len(site_channels))(*[int(c) for c in site_channels])
    out = np.zeros((n_samples, n_slices, int(sum(site_channels))), dtype=np.float32)
    _lib.check(_lib.lib().rcu_philox_masks_host(int(seed), float(p_drop), sc, len(site_channels), int(slice_index0), int(n_slices),
                                                int(sample0), int(n_samples), out.ctypes.data_as(_lib.c_float_p)))
    return out


def philox_keep_scale(seed, p_drop, site_channels, slice_index0, n_slices, sample0, n_samples):
    """Device version of philox_keep_scale_host (same stream, CUDA tensor)."""
    sc = (ctypes.c_int * len(site_channels))(*[int(c) for c in site_channels])
    out = torch.empty((n_samples, n_slices, int(sum(site_channels))), dtype=torch.float32, device=_device())
    _lib.check(_lib.lib().rcu_philox_masks(int(seed), float(p_drop), sc, len(site_channels), int(slice_index0), int(n_slices),
                                           int(sample0), int(n_samples), _lib.ptr(out), _lib.current_stream()))
    return out


def confusion_counts(prediction, target, n_subjects=1, sync=True):
    """tp, tn, fp, fn per subject with pymia's ConfusionMatrix semantics (== 1 / == 0 comparisons)."""
    dev = _pick_device(prediction, target)
    d_d = _to_device(prediction, torch.uint8, 'prediction', dev)
    t_d = _to_device(target, torch.uint8, 'target', dev)
    n = d_d.numel()
    vps = n // max(n_subjects, 1)
    _check_lengths(n, vps, n_subjects, target=t_d)
    out = torch.zeros((n_subjects, 4), dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().rcu_confusion(_lib.ptr(d_d), _lib.ptr(t_d), vps, n_subjects, _lib.ptr(out), _lib.current_stream()))
    return out.cpu().numpy() if sync else out


# --------------------------------------------------------------------------------------------------
# evaluation-side preparations (csrc/prepare.cu)
# --------------------------------------------------------------------------------------------------
def border_mask(binary_label_map, distance_in=1, distance_out=1):
    """common/utils/labelhelper.py:12-20 `boarder_mask` (mask only): uint8 CUDA tensor of the map's shape."""
    shape = tuple(binary_label_map.shape)
    if len(shape) not in (2, 3):
        raise ValueError('border mask needs a 2-D or 3-D label map, got shape {}'.format(shape))
    label = _to_device(binary_label_map if torch.is_tensor(binary_label_map) else np.asarray(binary_label_map) != 0, torch.uint8, 'label')
    d = (1,) * (3 - len(shape)) + shape
    out = torch.empty(label.numel(), dtype=torch.uint8, device=label.device)
    _lib.check(_lib.lib().rcu_border_mask(_lib.ptr(label), int(d[0]), int(d[1]), int(d[2]), int(distance_in), int(distance_out),
                                          _lib.ptr(out), _lib.current_stream()))
    return out.view(shape)


def minmax(values):
    """(min, max) as numpy float32 scalars — `entry_np.min(), entry_np.max()` (rechun/eval/analysis.py:176)."""
    v = _to_device(values, torch.float32, 'values')
    out = torch.empty(3, dtype=torch.int32, device=v.device)
    _lib.check(_lib.lib().rcu_minmax(_lib.ptr(v), v.numel(), _lib.ptr(out), _lib.current_stream()))
    k = out.cpu().numpy().view(np.uint32)
    if k[2]:
        return np.float32('nan'), np.float32('nan')   # numpy's min / max propagate NaN
    return np.float32(_lib.lib().rcu_minmax_decode(int(k[0]))), np.float32(_lib.lib().rcu_minmax_decode(int(k[1])))


def confidence_to_foreground(uncertainty, prediction, rescale='subject', epsilon=1e-5):
    """`rescale_uncertainties` + `uncertainty_to_foreground_probabilities` (rechun/eval/helper.py:7-22) in one pass.

    rescale: 'subject' (RescaleSubjectMinMax, analysis.py:169-178: float32 min / max of this map), a (min, max) pair of
    Python floats (RescaleLinear, analysis.py:154-166) or None (the map is already in [0, 1]).  Returns the foreground
    pseudo-probability as a float32 CUDA tensor of the input's shape; raises ValueError where the reference does."""
    shape = tuple(uncertainty.shape)
    if tuple(prediction.shape) != shape:
        raise ValueError('shapes must agree. Found {} and {}'.format(shape, tuple(prediction.shape)))
    u = _to_device(uncertainty, torch.float32, 'uncertainty')
    pred = _to_device(prediction, torch.uint8, 'prediction')
    lo = rng = 0.0
    if rescale == 'subject':
        mn, mx = minmax(u)
        lo, rng = float(mn), float(np.float32(mx) - np.float32(mn))          # float32 scalars: float32 arithmetic
    elif rescale is not None:
        mn, mx = rescale
        lo, rng = float(np.float32(mn)), float(np.float32(float(mx) - float(mn)))   # Python floats: range in double, then weak-cast
    out = torch.empty_like(u)
    invalid = torch.empty(2, dtype=torch.int64, device=u.device)
    _lib.check(_lib.lib().rcu_confidence_to_foreground(_lib.ptr(u), _lib.ptr(pred), u.numel(), int(rescale is not None), lo, rng,
                                                       float(np.float32(1 - 2 * epsilon)), float(np.float32(epsilon)), _lib.ptr(out),
                                                       _lib.ptr(invalid), _lib.current_stream()))
    bad, bad_pred = (int(v) for v in invalid.cpu())
    if bad_pred:
        raise ValueError('Found class larger than 1. Only works for binary problems')
    if bad:
        raise ValueError('Found {} values outside [0, 1] (rescale the uncertainty first)'.format(bad))
    return out.view(shape)
