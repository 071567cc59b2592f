"""ctypes binding of librcu_b200.so — the C-ABI declared in include/rcu_b200.h.

There is deliberately no fallback: if the library is missing (not built) every entry point raises, and if it
is loaded on a box without an sm_100 GPU the compute calls fail with the library's own error message.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'librcu_b200.so')
LIB_PATH = os.environ.get('RCU_B200_LIB', LIB_PATH)   # developer override (A/B builds)

RCU_ABI_VERSION = 3
RCU_OK, RCU_EINVAL, RCU_ECUDA, RCU_ENOTSUP, RCU_ENOMEM, RCU_ENCCL = 0, -1, -2, -3, -4, -5
RCU_COMM_ID_BYTES, RCU_IPC_HANDLE_BYTES = 128, 64
RCU_MAX_BINS, RCU_MAX_UE_CLASSES, RCU_MAX_BREAKS = 32, 32, 96

c_void_p, c_int, c_int64, c_uint64, c_size_t, c_float = (ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_uint64,
                                                         ctypes.c_size_t, ctypes.c_float)
c_float_p = ctypes.POINTER(ctypes.c_float)
c_double_p = ctypes.POINTER(ctypes.c_double)
c_int_p = ctypes.POINTER(ctypes.c_int)
c_uint8_p = ctypes.POINTER(ctypes.c_uint8)


class RcuConvUnit(ctypes.Structure):
    _fields_ = [('weight', c_float_p), ('bias', c_float_p), ('bn_weight', c_float_p), ('bn_bias', c_float_p),
                ('bn_mean', c_float_p), ('bn_var', c_float_p), ('c_in', c_int), ('c_out', c_int), ('has_dropout', c_int)]


class RcuUnetDesc(ctypes.Structure):
    _fields_ = [('in_channels', c_int), ('depth', c_int), ('start_filters', c_int), ('nb_classes', c_int),
                ('p_drop', c_float), ('bn_eps', c_float), ('units', ctypes.POINTER(RcuConvUnit)), ('n_units', c_int),
                ('upconvs', ctypes.POINTER(RcuConvUnit)), ('n_upconvs', c_int), ('head', RcuConvUnit),
                ('sigma_unit', ctypes.POINTER(RcuConvUnit)), ('sigma_head', ctypes.POINTER(RcuConvUnit)),
                ('residuals', ctypes.POINTER(RcuConvUnit)), ('n_residuals', c_int)]


class RcuUnetOutputs(ctypes.Structure):
    _fields_ = [('logits', c_void_p), ('sigma', c_void_p), ('features', c_void_p), ('postnet', c_void_p),
                ('postnet_logits', c_void_p), ('logit_diff', c_void_p)]


class RcuPeerLayout(ctypes.Structure):
    _fields_ = [('planes', c_int), ('off_sums', c_size_t), ('off_mean', c_size_t), ('off_entropy', c_size_t), ('off_mi', c_size_t),
                ('off_var', c_size_t), ('off_foreground', c_size_t), ('off_prediction', c_size_t), ('bytes', c_size_t)]


# name -> (restype, argtypes); mirrors include/rcu_b200.h one to one
PROTOTYPES = {
    'rcu_abi_version': (c_int, []),
    'rcu_last_error': (ctypes.c_char_p, []),
    'rcu_device_check': (c_int, [c_int]),
    'rcu_metrics_workspace_bytes': (c_size_t, [c_int]),
    'rcu_metrics_workspace_init': (c_int, [c_void_p, c_size_t, c_void_p]),
    'rcu_calib_hist': (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_float_p, c_int, c_float, c_float, c_void_p,
                               c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    'rcu_ue_hist': (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_float_p, c_double_p, c_int,
                            c_uint8_p, c_int, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    'rcu_eval_fused': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_float_p, c_int, c_float_p, c_int,
                               c_uint8_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t,
                               c_void_p]),
    'rcu_confusion': (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p]),
    'rcu_aggregate': (c_int, [c_void_p, c_int, c_int, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                              c_void_p, c_void_p, c_void_p]),
    'rcu_aggregate_ws': (c_int, [c_void_p, c_int, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                 c_void_p, c_void_p, c_void_p]),
    'rcu_aggregate_ws_diff': (c_int, [c_void_p, c_int, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                      c_void_p, c_void_p, c_void_p]),
    'rcu_aggregate_partial': (c_int, [c_void_p, c_int, c_int, c_int64, c_int64, c_int, c_int, c_void_p, c_void_p]),
    'rcu_aggregate_finish': (c_int, [c_void_p, c_int, c_int64, c_int64, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                     c_void_p, c_void_p, c_void_p]),
    'rcu_philox_masks': (c_int, [c_uint64, c_float, c_int_p, c_int, c_int64, c_int64, c_int, c_int, c_void_p, c_void_p]),
    'rcu_philox_masks_host': (c_int, [c_uint64, c_float, c_int_p, c_int, c_int64, c_int64, c_int, c_int, c_float_p]),
    'rcu_border_mask': (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    'rcu_minmax': (c_int, [c_void_p, c_int64, c_void_p, c_void_p]),
    'rcu_minmax_decode': (c_float, [ctypes.c_uint32]),
    'rcu_confidence_to_foreground': (c_int, [c_void_p, c_void_p, c_int64, c_int, c_float, c_float, c_float, c_float, c_void_p,
                                             c_void_p, c_void_p]),
    'rcu_unet_create': (c_int, [ctypes.POINTER(RcuUnetDesc), c_int, ctypes.POINTER(c_void_p)]),
    'rcu_unet_destroy': (None, [c_void_p]),
    'rcu_unet_plan': (c_int, [c_void_p, c_int, c_int, c_int, ctypes.POINTER(c_size_t)]),
    'rcu_unet_bind_workspace': (c_int, [c_void_p, c_void_p, c_size_t]),
    'rcu_unet_forward': (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_uint64, c_int64, c_int, c_void_p,
                                 c_void_p, c_void_p]),
    'rcu_unet_forward_ex': (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_uint64, c_int64, c_int, c_void_p,
                                    ctypes.POINTER(RcuUnetOutputs), c_void_p]),
    'rcu_postnet_create': (c_int, [ctypes.POINTER(RcuConvUnit), c_int, ctypes.POINTER(RcuConvUnit), c_float, c_int,
                                   ctypes.POINTER(c_void_p)]),
    'rcu_postnet_destroy': (None, [c_void_p]),
    'rcu_postnet_forward': (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_void_p]),
    'rcu_unet_total_dropout_channels': (c_int, [c_void_p]),
    'rcu_unet_debug_activation': (c_int, [c_void_p, c_int, c_void_p, c_size_t, c_void_p]),
    'rcu_unet_set_conv_impl': (c_int, [c_void_p, c_int]),
    'rcu_unet_set_halo_mask': (c_int, [c_void_p, c_uint64]),
    'rcu_unet_set_first_layer_dedup': (c_int, [c_void_p, c_int]),
    'rcu_unet_last_launch_count': (c_int64, [c_void_p]),
    'rcu_unet_enable_timing': (c_int, [c_void_p, c_int]),
    'rcu_unet_num_ops': (c_int, [c_void_p]),
    'rcu_unet_op_info': (c_int, [c_void_p, c_int, c_int_p, ctypes.POINTER(c_int64), c_int_p, c_int_p, c_int_p, c_int_p]),
    'rcu_unet_op_executed_macs': (c_int, [c_void_p, c_int, ctypes.POINTER(c_int64)]),
    'rcu_unet_read_timing': (c_int, [c_void_p, c_float_p, ctypes.POINTER(c_int64), c_int]),
    'rcu_comm_unique_id': (c_int, [c_void_p]),
    'rcu_comm_create': (c_int, [c_void_p, c_int, c_int, c_int, ctypes.POINTER(c_void_p)]),
    'rcu_comm_destroy': (None, [c_void_p]),
    'rcu_allreduce_probsum': (c_int, [c_void_p, c_void_p, c_int64, c_void_p]),
    'rcu_allreduce_counts': (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_void_p]),
    'rcu_peer_layout': (c_int, [c_int64, c_int64, c_int, c_int, ctypes.POINTER(RcuPeerLayout)]),
    'rcu_peer_region_alloc': (c_int, [c_size_t, c_int, ctypes.POINTER(c_void_p), c_void_p]),
    'rcu_peer_region_open': (c_int, [c_void_p, c_int, ctypes.POINTER(c_void_p)]),
    'rcu_peer_region_close': (c_int, [c_void_p]),
    'rcu_peer_region_free': (c_int, [c_void_p]),
    'rcu_aggregate_finish_peer': (c_int, [ctypes.POINTER(c_void_p), c_int, c_int, ctypes.c_uint32, c_int, c_int64, c_int64, c_int, c_int,
                                          c_void_p]),
}

_lib = None


class RcuError(RuntimeError):
    pass


def lib():
    """The loaded library (loads on first use; raises if it has not been built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RcuError('{} is missing — run `python __graft_entry__.py` (or {}/build.py) to compile the CUDA '
                           'extension; there is no CPU fallback'.format(LIB_PATH, _HERE))
        handle = ctypes.CDLL(LIB_PATH)
        for name, (restype, argtypes) in PROTOTYPES.items():
            fn = getattr(handle, name)  # AttributeError here = header/library drift
            fn.restype = restype
            fn.argtypes = argtypes
        if handle.rcu_abi_version() != RCU_ABI_VERSION:
            raise RcuError('librcu_b200.so ABI version {} != {} (rebuild: python __graft_entry__.py)'.format(handle.rcu_abi_version(), RCU_ABI_VERSION))
        _lib = handle
    return _lib


def last_error():
    msg = lib().rcu_last_error()
    return msg.decode('utf-8', 'replace') if msg else ''


def check(rc):
    """Turn a status code into the exception the reference would raise for the same mistake."""
    if rc == RCU_OK:
        return
    msg = last_error()
    if rc == RCU_EINVAL:
        raise ValueError(msg)
    if rc == RCU_ENOTSUP:
        raise NotImplementedError(msg)
    raise RcuError('rcu_b200 error {}: {}'.format(rc, msg))


def ptr(t):
    """Device/host address of a torch tensor (or None)."""
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def current_stream():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
