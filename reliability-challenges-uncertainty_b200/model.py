"""B200UNet — drop-in for the reference's `context.model` on the inference path.

`B200UNet.from_reference(module)` consumes a loaded reference `UNet` (common/model/unet.py:123-186: its
state_dict and the placement of its Dropout2d modules) and runs the forward through rcu_unet_* (tcgen05
implicit-GEMM convolutions, BN folded, ReLU + Dropout2d keep-scale in the epilogue, MC samples folded into the
batch).  It honours the reference's MC-dropout switch: `th.set_dropout_mode(model, True)`
(common/utils/torchhelper.py:44-50) flips every nn.Dropout2d submodule to train mode — this module owns one such
submodule as its switch, so the unmodified `McPredictStep` loop drives it correctly (each stochastic call is the
next Philox sample).  The fused steps in steps.py use `forward_samples` instead and get all T+1 passes in one go.
"""
import ctypes
import os

import numpy as np
import torch
import torch.nn as nn

from . import _lib

BN_EPS = 1e-5  # nn.BatchNorm2d default (common/model/unet.py:17)


def block_dropout_flags(dropout, dropout_center, level, depth, is_down):
    """(first conv, second conv) of the block at `level` carry a Dropout2d — the placement rule of
    common/model/unet.py:63-82 (_get_dropout / _get_dropout_mode)."""
    if dropout is None:
        return (False, False)
    if dropout_center is None:
        return (True, True)
    if level == depth:
        return (False, False)
    if level + dropout_center >= depth:
        return (False, True) if is_down else (True, False)
    return (False, False)


def unit_layout(in_channels, depth, start_filters, dropout, dropout_center):
    """[(state_dict prefix, c_in, c_out, has_dropout)] for every Conv2dBnRelu in forward order, and
    [(prefix, c_in, c_out)] for the up-path `upconv` convolutions."""
    units, upconvs = [], []
    c_in, c_out = in_channels, start_filters
    for lvl in range(depth):
        flags = block_dropout_flags(dropout, dropout_center, lvl, depth, True)
        units.append(('down_convs.%d.block.block.0.conv2d_batch_relu' % lvl, c_in, c_out, flags[0]))
        units.append(('down_convs.%d.block.block.1.conv2d_batch_relu' % lvl, c_out, c_out, flags[1]))
        c_in, c_out = c_out, 2 * c_out
    flags = block_dropout_flags(dropout, dropout_center, depth, depth, True)
    units.append(('bottom_convs.block.0.conv2d_batch_relu', c_in, c_out, flags[0]))
    units.append(('bottom_convs.block.1.conv2d_batch_relu', c_out, c_out, flags[1]))
    for j in range(depth):
        lvl = depth - 1 - j
        c_in, c_out = c_out, c_out // 2
        upconvs.append(('up_convs.%d.upconv.1' % j, c_in, c_out))
        flags = block_dropout_flags(dropout, dropout_center, lvl, depth, False)
        units.append(('up_convs.%d.block.block.0.conv2d_batch_relu' % j, 2 * c_out, c_out, flags[0]))
        units.append(('up_convs.%d.block.block.1.conv2d_batch_relu' % j, c_out, c_out, flags[1]))
    units.append(('conv_cls.0.conv2d_batch_relu', c_out, c_out, dropout is not None))
    return units, upconvs


def _f32(t):
    return np.ascontiguousarray(t.detach().to('cpu', torch.float32).numpy())


_WORKSPACES = {}


def shared_workspace(device, nbytes):
    """The per-device activation workspace (a uint8 CUDA tensor of at least `nbytes`), shared by all engines of the device.
    Sharing assumes what the reference's test loop does: one Python thread, forwards ordered on one stream."""
    device = torch.device(device)
    key = (device.type, device.index if device.index is not None else torch.cuda.current_device())
    buf = _WORKSPACES.get(key)
    if buf is None or buf.numel() < nbytes:
        if buf is not None:
            _WORKSPACES.pop(key)
            del buf
        buf = torch.empty(int(nbytes), dtype=torch.uint8, device=device)
        assert buf.data_ptr() % 1024 == 0
        _WORKSPACES[key] = buf
    return buf


def release_workspaces():
    """Drop the shared activation workspaces (engines re-acquire one on their next forward)."""
    _WORKSPACES.clear()


class B200UNet(nn.Module):
    """See module docstring.  Not trainable: there are no parameters, only a device handle."""

    DEFAULT_CHUNK_IMAGES = int(os.environ.get('RCU_B200_CHUNK_IMAGES', '147'))

    def __init__(self, state_dict, nb_classes=2, in_channels=4, depth=4, start_filters=32, dropout=0.05,
                 dropout_center=None, device=None, seed=20, chunk_images=None, sigma_out=None, provide_features=False):
        super().__init__()
        if nb_classes != 2:
            raise NotImplementedError('the B200 hot path is binary (nb_classes == 2)')
        state_dict = {(k[len('module.'):] if k.startswith('module.') else k): v for k, v in state_dict.items()}
        # residual=True (unet.py:42-60, 144-148): every block adds a 1x1 convolution of its input and its last conv has no ReLU
        self.residual = any(k.endswith('.residual.weight') for k in state_dict)
        has_sigma_weights = any(k.startswith('conv_sigma.') for k in state_dict)
        # sigma_out=True (unet.py:162-164): a second head on the same features, forward returns (logits, sigma)
        self.sigma_out = has_sigma_weights if sigma_out is None else bool(sigma_out)
        if self.sigma_out and not has_sigma_weights:
            raise ValueError('sigma_out=True but the state_dict has no conv_sigma weights')
        self.nb_classes, self.in_channels, self.depth, self.start_filters = nb_classes, in_channels, depth, start_filters
        self.dropout, self.dropout_center = dropout, dropout_center
        self.seed = int(seed)
        self.chunk_images = int(chunk_images or self.DEFAULT_CHUNK_IMAGES)
        # the reference toggles MC dropout by flipping nn.Dropout2d modules (torchhelper.py:44-50); this is ours
        self.mc_dropout_switch = nn.Dropout2d(p=dropout if dropout is not None else 0.0)
        self.provide_features = bool(provide_features)   # unet.py:178-179: forward() keeps the conv_cls input in .features
        self.features = None
        self._next_sample = 0      # Philox sample index of the next free-running stochastic forward() call
        self._next_slice = 0       # run-global slice counter for forward() calls
        self._handle = None
        self._plan = None
        self._workspace = None
        self._device = torch.device(device) if device is not None else torch.device('cuda', torch.cuda.current_device())
        units, upconvs = unit_layout(in_channels, depth, start_filters, dropout, dropout_center)
        self.site_channels = [co for (_, _, co, d) in units if d]
        self._create(state_dict, units, upconvs)
        self.eval()  # inference engine: the Dropout2d switch starts in eval mode, like load_from_checkpoint leaves the model (context.py:321)

    # ------------------------------------------------------------------ construction
    @classmethod
    def from_reference(cls, module, device=None, seed=20, chunk_images=None):
        """Build from a loaded reference `common.model.unet.UNet` (any object with its attribute layout)."""
        if isinstance(module, nn.DataParallel):
            module = module.module
        sd = module.state_dict()
        depth = len(module.down_convs)
        w0 = sd['down_convs.0.block.block.0.conv2d_batch_relu.conv.weight']
        start_filters, in_channels = int(w0.shape[0]), int(w0.shape[1])
        nb_classes = int(sd['conv_cls.1.weight'].shape[0])
        if not any(k.endswith('.bn.weight') for k in sd):
            raise NotImplementedError('bn=False nets are outside the B200 hot path')
        for up in module.up_convs:
            if isinstance(up.upconv, nn.ConvTranspose2d):
                raise NotImplementedError('transpose=True up-convolutions are outside the B200 hot path')
        # dropout placement is read off the module tree, not guessed from constructor arguments
        drop_ps, has_drop = [], {}
        for name, m in module.named_modules():
            if isinstance(m, nn.Dropout2d):
                drop_ps.append(m.p)
                has_drop[name.rsplit('.', 1)[0]] = True
        dropout = drop_ps[0] if drop_ps else None
        if any(abs(p - dropout) > 0 for p in drop_ps):
            raise NotImplementedError('per-layer dropout probabilities differ')
        # recover dropout_center by matching the observed placement
        for center in [None] + list(range(0, depth + 1)):
            units, _ = unit_layout(in_channels, depth, start_filters, dropout, center)
            if all(bool(has_drop.get(p, False)) == d for (p, _, _, d) in units):
                dropout_center = center
                break
        else:
            raise NotImplementedError('unrecognised Dropout2d placement')
        if device is None:
            p0 = next(module.parameters())
            device = p0.device if p0.is_cuda else None
        net = cls(sd, nb_classes, in_channels, depth, start_filters, dropout, dropout_center, device, seed, chunk_images,
                  sigma_out=getattr(module, 'conv_sigma', None) is not None,
                  provide_features=bool(getattr(module, 'provide_features', False)))
        net.mc_dropout_switch.train(any(m.training for m in module.modules() if isinstance(m, nn.Dropout2d)))
        return net

    def _create(self, sd, units, upconvs):
        keep = []  # numpy arrays referenced by the ctypes structs

        def arr(key):
            if key not in sd:
                raise ValueError('state_dict is missing "{}"'.format(key))
            a = _f32(sd[key])
            keep.append(a)
            return a.ctypes.data_as(_lib.c_float_p)

        def unit_struct(prefix, c_in, c_out, has_dropout, bn=True, conv_key='.conv'):
            u = _lib.RcuConvUnit()
            u.weight = arr(prefix + conv_key + '.weight')
            u.bias = arr(prefix + conv_key + '.bias')
            if bn:
                u.bn_weight = arr(prefix + '.bn.weight')
                u.bn_bias = arr(prefix + '.bn.bias')
                u.bn_mean = arr(prefix + '.bn.running_mean')
                u.bn_var = arr(prefix + '.bn.running_var')
            w = sd[prefix + conv_key + '.weight']
            if tuple(w.shape[:2]) != (c_out, c_in):
                raise ValueError('{}: weight shape {} does not match the topology ({}, {})'.format(prefix, tuple(w.shape), c_out, c_in))
            u.c_in, u.c_out, u.has_dropout = c_in, c_out, int(bool(has_dropout))
            return u

        unit_arr = (_lib.RcuConvUnit * len(units))(*[unit_struct(p, ci, co, d) for (p, ci, co, d) in units])
        up_arr = (_lib.RcuConvUnit * len(upconvs))(*[unit_struct(p, ci, co, False, bn=False, conv_key='') for (p, ci, co) in upconvs])
        desc = _lib.RcuUnetDesc()
        desc.in_channels, desc.depth, desc.start_filters, desc.nb_classes = self.in_channels, self.depth, self.start_filters, self.nb_classes
        desc.p_drop = float(self.dropout) if self.dropout is not None else 0.0
        desc.bn_eps = BN_EPS
        desc.units, desc.n_units = unit_arr, len(units)
        desc.upconvs, desc.n_upconvs = up_arr, len(upconvs)
        desc.head = unit_struct('conv_cls.1', self.start_filters, self.nb_classes, False, bn=False, conv_key='')
        if self.residual:
            names = (['down_convs.%d.block.residual' % l for l in range(self.depth)] + ['bottom_convs.residual'] +
                     ['up_convs.%d.block.residual' % j for j in range(self.depth)])
            res = []
            for name in names:
                w = sd[name + '.weight'] if name + '.weight' in sd else None
                if w is None:
                    raise ValueError('state_dict is missing "{}.weight" (residual=True nets carry one per block)'.format(name))
                res.append(unit_struct(name, int(w.shape[1]), int(w.shape[0]), False, bn=False, conv_key=''))
            res_arr = (_lib.RcuConvUnit * len(res))(*res)
            desc.residuals, desc.n_residuals = res_arr, len(res)
        if self.sigma_out:
            sf = self.start_filters
            sigma_unit = unit_struct('conv_sigma.0.conv2d_batch_relu', sf, sf, self.dropout is not None)
            sigma_head = unit_struct('conv_sigma.1', sf, self.nb_classes, False, bn=False, conv_key='')
            desc.sigma_unit, desc.sigma_head = ctypes.pointer(sigma_unit), ctypes.pointer(sigma_head)
            if self.dropout is not None:
                self.site_channels.append(sf)   # conv_sigma.0's Dropout2d is the last site in forward order
        handle = ctypes.c_void_p()
        dev_index = self._device.index if self._device.index is not None else torch.cuda.current_device()
        _lib.check(_lib.lib().rcu_unet_create(ctypes.byref(desc), int(dev_index), ctypes.byref(handle)))
        self._handle = handle
        del keep

    def __del__(self):
        try:
            h = self.__dict__.get('_handle')
            if h:
                self.__dict__['_handle'] = None
                _lib.lib().rcu_unet_destroy(h)
        except Exception:  # interpreter shutdown
            pass

    # ------------------------------------------------------------------ nn.Module surface
    def to(self, *args, **kwargs):  # weights already live on the device chosen at construction
        return self

    @property
    def device(self):
        return self._device

    @property
    def total_dropout_channels(self):
        return int(sum(self.site_channels))

    def set_conv_impl(self, impl):
        """0 = tcgen05 (product path), 1 = CUDA-core cross-check kernels (tests only), 2 = tcgen05 per-tap kernel only,
        3 = tcgen05 without the pixel-pair kernel (A/B)."""
        _lib.check(_lib.lib().rcu_unet_set_conv_impl(self._handle, int(impl)))

    def set_first_layer_dedup(self, enable):
        """First-layer dedup (default on): the first unit's output is stored once per slice and the dropped channels of each
        (sample, slice) are patched into the consumer's tiles; False materialises it per (sample, slice) (A/B, debug)."""
        _lib.check(_lib.lib().rcu_unet_set_first_layer_dedup(self._handle, int(bool(enable))))

    def set_halo_mask(self, mask):
        """Debug: bit i lets conv i (execution order) use the halo-tile kernel; default all ones."""
        _lib.check(_lib.lib().rcu_unet_set_halo_mask(self._handle, int(mask) & 0xFFFFFFFFFFFFFFFF))

    def last_launch_count(self):
        return int(_lib.lib().rcu_unet_last_launch_count(self._handle))

    def reset_stream(self, seed=None, slice_index=0, sample=0):
        """Rewind the free-running Philox position used by forward() in MC mode."""
        if seed is not None:
            self.seed = int(seed)
        self._next_slice, self._next_sample = int(slice_index), int(sample)

    def _ensure_plan(self, h, w, n_samples):
        need = max(self.chunk_images, n_samples)
        if self._plan is None or self._plan[:2] != (h, w) or self._plan[2] < need:
            ws = ctypes.c_size_t()
            _lib.check(_lib.lib().rcu_unet_plan(self._handle, int(h), int(w), int(need), ctypes.byref(ws)))
            self._plan = (h, w, need, int(ws.value))
            self._workspace = None
        # the activation workspace is ours, not the library's: one buffer per device, shared by every engine on it (their
        # forwards are ordered on the current stream), grown on demand
        arena = shared_workspace(self._device, self._plan[3])
        if self._workspace is not arena:
            with torch.cuda.device(self._device):
                _lib.check(_lib.lib().rcu_unet_bind_workspace(self._handle, _lib.ptr(arena), arena.numel()))
            self._workspace = arena

    def workspace_bytes(self):
        """Bytes of activation workspace the current plan needs (0 before the first forward)."""
        return 0 if self._plan is None else self._plan[3]

    def forward_samples(self, images, n_samples=1, dropout_mode=0, det_first=False, seed=None, slice_index0=0, sample0=0,
                        scale=None, diff=False):
        """All samples of all slices in one call.

        images: (N, C, H, W) tensor (moved to the device as float32).  Returns pixel-interleaved logits
        float32 (n_samples, N, H, W, 2) — or, with diff=True, their differences l0 - l1, float32 (n_samples, N, H, W):
        all a softmax over the two classes needs, at half the bytes (rcu_unet_outputs.logit_diff).
        dropout_mode: 0 eval, 1 Philox MC dropout (sample ids sample0...),
        2 caller-supplied keep-scale table `scale` float32 (n_stochastic, N, total_dropout_channels)."""
        out = self.forward_outputs(images, n_samples, dropout_mode, det_first, seed, slice_index0, sample0, scale, logit_diff=diff)
        return out['logit_diff' if diff else 'logits']

    def forward_outputs(self, images, n_samples=1, dropout_mode=0, det_first=False, seed=None, slice_index0=0, sample0=0,
                        scale=None, sigma=False, features=False, postnet=None, logit_diff=False):
        """forward_samples plus the optional outputs of rcu_unet_forward_ex, as a dict:
          'logits'          (n_samples, N, H, W, 2) interleaved — or, with logit_diff=True, INSTEAD of it
          'logit_diff'      (n_samples, N, H, W) = l0 - l1 of the class head
          'sigma'           same layout, raw conv_sigma output (sigma=True, sigma_out nets)
          'features'        (n_samples, N, start_filters, H, W) float32 = UNet.features (features=True)
          'postnet_logits'  (n_samples, N, H, W, 2): `postnet` (a B200PostNet) applied to the features inside the same
                            call, while they are still bf16 in the workspace"""
        if images.dim() != 4 or images.shape[1] != self.in_channels:
            raise ValueError('expected images of shape (N, {}, H, W), got {}'.format(self.in_channels, tuple(images.shape)))
        x = images.to(self._device, torch.float32).contiguous()
        n, _, h, w = x.shape
        self._ensure_plan(h, w, n_samples)
        outputs = _lib.RcuUnetOutputs()
        if logit_diff:
            result = {'logit_diff': torch.empty((n_samples, n, h, w), dtype=torch.float32, device=self._device)}
            outputs.logit_diff = result['logit_diff'].data_ptr()
        else:
            result = {'logits': torch.empty((n_samples, n, h, w, 2), dtype=torch.float32, device=self._device)}
            outputs.logits = result['logits'].data_ptr()
        pair_shape = (n_samples, n, h, w, 2)
        if sigma:
            if not self.sigma_out:
                raise ValueError('this net has no sigma head (sigma_out=False)')
            result['sigma'] = torch.empty(pair_shape, dtype=torch.float32, device=self._device)
            outputs.sigma = result['sigma'].data_ptr()
        if features:
            result['features'] = torch.empty((n_samples, n, self.start_filters, h, w), dtype=torch.float32, device=self._device)
            outputs.features = result['features'].data_ptr()
        if postnet is not None:
            result['postnet_logits'] = torch.empty(pair_shape, dtype=torch.float32, device=self._device)
            outputs.postnet = postnet._handle
            outputs.postnet_logits = result['postnet_logits'].data_ptr()
        scale_d = None
        if dropout_mode == 2:
            n_stoch = n_samples - (1 if det_first else 0)
            scale_d = torch.as_tensor(scale, dtype=torch.float32).to(self._device).contiguous()
            if tuple(scale_d.shape) != (n_stoch, n, self.total_dropout_channels):
                raise ValueError('scale must have shape {}, got {}'.format((n_stoch, n, self.total_dropout_channels), tuple(scale_d.shape)))
        with torch.cuda.device(self._device):
            _lib.check(_lib.lib().rcu_unet_forward_ex(self._handle, _lib.ptr(x), n, int(n_samples), int(dropout_mode),
                                                      int(bool(det_first)), int(self.seed if seed is None else seed),
                                                      int(slice_index0), int(sample0), _lib.ptr(scale_d), ctypes.byref(outputs),
                                                      _lib.current_stream()))
        return result

    def forward(self, x):
        """`model(images)` of the reference: logits (N, 2, H, W) float32 (a channels-last strided view), or
        `(logits, sigma)` for a sigma_out net (unet.py:181-186); with provide_features the conv_cls input is kept in
        `self.features` as (N, start_filters, H, W) float32 (unet.py:178-179).

        Deterministic while the Dropout2d switch is in eval mode; in train mode (th.set_dropout_mode(model, True))
        every call is the next MC sample of the Philox stream for these slices."""
        kw = dict(sigma=self.sigma_out, features=self.provide_features)
        if self.mc_dropout_switch.training and self.dropout:
            out = self.forward_outputs(x, 1, dropout_mode=1, slice_index0=self._next_slice, sample0=self._next_sample, **kw)
            self._next_sample += 1
        else:
            out = self.forward_outputs(x, 1, dropout_mode=0, **kw)
        if self.provide_features:
            self.features = out['features'][0]
        logits = out['logits'][0].permute(0, 3, 1, 2)
        if self.sigma_out:
            return logits, out['sigma'][0].permute(0, 3, 1, 2)
        return logits

    def enable_timing(self, enable=True):
        """Bracket every kernel launch of the forward with CUDA events (read them with read_timing)."""
        _lib.check(_lib.lib().rcu_unet_enable_timing(self._handle, int(bool(enable))))

    def op_table(self):
        """[{kind, macs_per_image, c_in, c_out, h, w}] for every op of the planned schedule (last = coefficient table)."""
        n = int(_lib.lib().rcu_unet_num_ops(self._handle))
        out = []
        for i in range(n):
            kind, ci, co, h, w = (ctypes.c_int() for _ in range(5))
            macs = ctypes.c_int64()
            _lib.check(_lib.lib().rcu_unet_op_info(self._handle, i, ctypes.byref(kind), ctypes.byref(macs), ctypes.byref(ci),
                                                   ctypes.byref(co), ctypes.byref(h), ctypes.byref(w)))
            ex = ctypes.c_int64()
            _lib.check(_lib.lib().rcu_unet_op_executed_macs(self._handle, i, ctypes.byref(ex)))
            out.append({'kind': ('first_conv', 'conv', 'maxpool', 'coef')[kind.value], 'macs_per_image': int(macs.value),
                        'executed_macs_per_image': int(ex.value), 'c_in': ci.value, 'c_out': co.value, 'h': h.value, 'w': w.value})
        return out

    def read_timing(self):
        """(ms float32[n_ops], launches int64[n_ops]) accumulated since the last read; synchronises the events."""
        n = int(_lib.lib().rcu_unet_num_ops(self._handle))
        ms = np.zeros(n, dtype=np.float32)
        launches = np.zeros(n, dtype=np.int64)
        _lib.check(_lib.lib().rcu_unet_read_timing(self._handle, ms.ctypes.data_as(_lib.c_float_p),
                                                   launches.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), n))
        return ms, launches

    def debug_activation(self, index, shape):
        """fp32 copy of the index-th internal activation (NHWC) of the last chunk of the last forward."""
        out = torch.empty(shape, dtype=torch.float32, device=self._device)
        _lib.check(_lib.lib().rcu_unet_debug_activation(self._handle, int(index), _lib.ptr(out), out.numel(), _lib.current_stream()))
        return out


class B200PostNet(nn.Module):
    """`context.model` drop-in for common/model/postnet.py:6-17 PostNet(in_channels=32, nb_classes=2, nb_convs, dropout):
    the auxiliary-feature method's per-pixel stack of 1x1 Conv2dBnRelu units + 1x1 logits conv, in eval mode
    (bin-dl/brats_test_auxiliary_feat.py:61-80).  `forward(features)` takes the (N, 32, H, W) tensor UNet.features holds;
    `B200UNet.forward_outputs(..., postnet=self)` runs it fused with the U-Net forward instead."""

    def __init__(self, state_dict, device=None):
        super().__init__()
        sd = {(k[len('module.'):] if k.startswith('module.') else k): v for k, v in state_dict.items()}
        self._handle = None
        self._device = torch.device(device) if device is not None else torch.device('cuda', torch.cuda.current_device())
        n_convs = 0
        while 'convs.{}.conv2d_batch_relu.conv.weight'.format(n_convs) in sd:
            n_convs += 1
        if n_convs == 0 or 'conv_logits.weight' not in sd:
            raise ValueError('state_dict is not a PostNet (convs.N.conv2d_batch_relu.*, conv_logits.*)')
        if int(sd['conv_logits.weight'].shape[0]) != 2:
            raise NotImplementedError('the B200 hot path is binary (nb_classes == 2)')
        self.nb_convs = n_convs
        self.in_channels = int(sd['conv_logits.weight'].shape[1])
        keep = []

        def arr(key):
            a = _f32(sd[key])
            keep.append(a)
            return a.ctypes.data_as(_lib.c_float_p)

        def unit(prefix, conv_key, bn, c_in, c_out):
            u = _lib.RcuConvUnit()
            u.weight, u.bias = arr(prefix + conv_key + '.weight'), arr(prefix + conv_key + '.bias')
            if bn:
                u.bn_weight, u.bn_bias = arr(prefix + '.bn.weight'), arr(prefix + '.bn.bias')
                u.bn_mean, u.bn_var = arr(prefix + '.bn.running_mean'), arr(prefix + '.bn.running_var')
            u.c_in, u.c_out, u.has_dropout = c_in, c_out, 0
            return u
        c = self.in_channels
        for i in range(n_convs):
            w = sd['convs.{}.conv2d_batch_relu.conv.weight'.format(i)]
            if tuple(w.shape) != (c, c, 1, 1):
                raise ValueError('convs.{}: weight shape {} is not ({}, {}, 1, 1)'.format(i, tuple(w.shape), c, c))
        units = (_lib.RcuConvUnit * n_convs)(*[unit('convs.{}.conv2d_batch_relu'.format(i), '.conv', True, c, c) for i in range(n_convs)])
        head = unit('conv_logits', '', False, c, 2)
        handle = ctypes.c_void_p()
        dev_index = self._device.index if self._device.index is not None else torch.cuda.current_device()
        _lib.check(_lib.lib().rcu_postnet_create(units, n_convs, ctypes.byref(head), BN_EPS, int(dev_index), ctypes.byref(handle)))
        self._handle = handle
        self.eval()

    @classmethod
    def from_reference(cls, module, device=None):
        if isinstance(module, nn.DataParallel):
            module = module.module
        if device is None:
            p0 = next(module.parameters())
            device = p0.device if p0.is_cuda else None
        return cls(module.state_dict(), device)

    def __del__(self):
        try:
            h = self.__dict__.get('_handle')
            if h:
                self.__dict__['_handle'] = None
                _lib.lib().rcu_postnet_destroy(h)
        except Exception:  # interpreter shutdown
            pass

    def to(self, *args, **kwargs):
        return self

    def forward(self, features):
        if features.dim() != 4 or features.shape[1] != self.in_channels:
            raise ValueError('expected features of shape (N, {}, H, W), got {}'.format(self.in_channels, tuple(features.shape)))
        x = features.to(self._device, torch.float32).contiguous()
        n, _, h, w = x.shape
        logits = torch.empty((n, h, w, 2), dtype=torch.float32, device=self._device)
        with torch.cuda.device(self._device):
            _lib.check(_lib.lib().rcu_postnet_forward(self._handle, _lib.ptr(x), n, h * w, _lib.ptr(logits), _lib.current_stream()))
        return logits.permute(0, 3, 1, 2)
