"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference's stochastic-inference + uncertainty path.

The reference is pure Python; its arithmetic on this path is torch (CPU) and numpy.  This module restates
that arithmetic with the same library primitives, independent of the reference tree, so it can travel to
the GPU box (where /root/reference does not exist).  It is *pinned*: tests/test_oracle_golden.py checks
every function here against fixtures under tests/golden/ that tests/golden/make_golden.py produced by
running the unmodified reference (through oracle/ref_shim.py) in the authoring container.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
module, and only as the checker / the CPU baseline.  The product package never imports it.

Reference citations (path:line relative to the reference root):
  U-Net forward ............ common/model/unet.py:8-23,26-39,63-82,85-120,123-186; common/model/helpers.py:5-16
  MC / ensemble steps ...... rechun/dl/customsteps.py:16-39; bin-dl/brats_test_ensemble.py:78-94
  deterministic step ....... common/trainloop/steps.py:69-90
  summary .................. rechun/dl/customsteps.py:50-71; common/utils/torchhelper.py:53-54
  ECE ...................... common/evalutation/numpyfunctions.py:6-23,26-48,51-69,72-83
  U-E counts / ratios ...... common/evalutation/numpyfunctions.py:86-125; common/evalutation/eval.py:145-226
  Dice / accuracy / cm ..... common/evalutation/numpyfunctions.py:128-151 (+ pymia 0.2.1 metric classes, restated)
  eval-side preparation .... rechun/eval/helper.py:25-47; rechun/eval/analysis.py:147-151,189-203
  threshold sweep / table .. bin-eval/eval_uncertainty.py:176-202,239; bin-analysis/table_ece_ue_bnf_dice.py:56-59
  sigma head / aleatoric ... common/model/unet.py:162-164,181-186; bin-dl/brats_test_aleatoric.py:51-73
  features / PostNet ....... common/model/unet.py:178-179; common/model/postnet.py:6-17; bin-dl/brats_test_auxiliary_feat.py:61-80
  auxiliary segm input ..... bin-dl/brats_test_auxiliary_segm.py:48-69
  border mask .............. common/utils/labelhelper.py:12-20
  confidence -> p .......... rechun/eval/helper.py:7-22
  CSV rows / report tables . rechun/eval/hook.py:27-93; bin-analysis/table_supplmat_ece_dataset_vs_meansubject.py:59-104;
                             bin-analysis/table_ece_ue_bnf_dice.py:30-73,132-143 (the last one UNPINNED: the installed
                             pandas 3 rejects the reference's groupby.mean over string columns, so it cannot be run)
"""
import math
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-5  # nn.BatchNorm2d default, common/model/unet.py:17
SWEEP_THRESHOLDS = (0.05, 0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9, 0.95)  # bin-eval/eval_uncertainty.py:239


# ------------------------------------------------------------------------------------------------
# U-Net topology (common/model/unet.py:128-164)
# ------------------------------------------------------------------------------------------------
class UNetConfig:
    """Constructor arguments of the reference UNet that the hot path supports."""

    def __init__(self, nb_classes=2, in_channels=4, depth=4, start_filters=32, dropout=0.05, dropout_center=None,
                 sigma_out=False, residual=False):
        self.sigma_out = sigma_out
        self.residual = residual      # ConvResidualBlock instead of ConvBlock (unet.py:42-60, 133)
        self.nb_classes = nb_classes
        self.in_channels = in_channels
        self.depth = depth
        self.start_filters = start_filters
        self.dropout = dropout
        self.dropout_center = dropout_center


def _block_dropout_flags(cfg, level, is_down):
    """Which of the two convs of the block at `level` carry a Dropout2d (unet.py:63-82)."""
    if cfg.dropout is None:
        return (False, False)
    if cfg.dropout_center is None:
        return (True, True)
    if level == cfg.depth:
        return (False, False)
    if level + cfg.dropout_center >= cfg.depth:
        return (False, True) if is_down else (True, False)
    return (False, False)


def conv_sites(cfg):
    """Ordered list of every 3x3 'conv [-> dropout] -> bn -> relu' unit, in forward order.

    Each entry: (state_dict prefix, c_in, c_out, has_dropout).  The order is the order in which the
    reference's forward visits them, which is also the order dropout sites are numbered in.
    """
    sites = []
    c_in, c_out = cfg.in_channels, cfg.start_filters
    for lvl in range(cfg.depth):
        flags = _block_dropout_flags(cfg, lvl, True)
        for rep in range(2):
            sites.append(('down_convs.%d.block.block.%d.conv2d_batch_relu' % (lvl, rep),
                          c_in if rep == 0 else c_out, c_out, flags[rep]))
        c_in, c_out = c_out, c_out * 2
    flags = _block_dropout_flags(cfg, cfg.depth, True)
    for rep in range(2):
        sites.append(('bottom_convs.block.%d.conv2d_batch_relu' % rep, c_in if rep == 0 else c_out, c_out, flags[rep]))
    for j, lvl in enumerate(range(cfg.depth - 1, -1, -1)):
        c_in, c_out = c_out, c_out // 2
        flags = _block_dropout_flags(cfg, lvl, False)
        for rep in range(2):
            sites.append(('up_convs.%d.block.block.%d.conv2d_batch_relu' % (j, rep),
                          2 * c_out if rep == 0 else c_out, c_out, flags[rep]))
    sites.append(('conv_cls.0.conv2d_batch_relu', c_out, c_out, cfg.dropout is not None))
    return sites


SIGMA_PREFIX = 'conv_sigma.0.conv2d_batch_relu'


def residual_prefix(cfg, block):
    """state_dict prefix of the `residual` 1x1 conv of block number `block` in forward order (down 0.., bottom, up 0..)."""
    if block < cfg.depth:
        return 'down_convs.%d.block.residual' % block
    if block == cfg.depth:
        return 'bottom_convs.residual'
    return 'up_convs.%d.block.residual' % (block - cfg.depth - 1)


def dropout_sites(cfg):
    """[(prefix, channels)] of the units that own a Dropout2d, in forward order (conv_sigma.0 of a sigma_out net
    runs after conv_cls, unet.py:181-185)."""
    sites = [(p, co) for (p, ci, co, d) in conv_sites(cfg) if d]
    if cfg.sigma_out and cfg.dropout is not None:
        sites.append((SIGMA_PREFIX, cfg.start_filters))
    return sites


def init_state_dict(cfg, seed):
    """Reproduce `torch.manual_seed(seed); UNet(**cfg).state_dict()` without the reference tree.

    Parameter creation order follows the reference constructor (unet.py:134-164): for every block the conv
    (kaiming-uniform weight, then uniform bias — torch's nn.Conv2d.reset_parameters) then the BN (no RNG);
    in each UpConv the block is built *before* the upconv conv (unet.py:155-157: the block is an argument
    expression evaluated before UpConv.__init__ creates `upconv`); finally conv_cls.0 then conv_cls.1.
    Pinned bit-for-bit against the reference by tests/golden (same torch version on both boxes).
    """
    gen_state = torch.random.get_rng_state()
    torch.manual_seed(seed)
    sd = OrderedDict()

    def conv(prefix_w, c_out, c_in, k):
        w = torch.empty(c_out, c_in, k, k)
        torch.nn.init.kaiming_uniform_(w, a=math.sqrt(5))
        bound = 1 / math.sqrt(c_in * k * k)
        b = torch.empty(c_out)
        torch.nn.init.uniform_(b, -bound, bound)
        sd[prefix_w + '.weight'] = w
        sd[prefix_w + '.bias'] = b

    def unit(prefix, c_in, c_out):
        conv(prefix + '.conv', c_out, c_in, 3)
        sd[prefix + '.bn.weight'] = torch.ones(c_out)
        sd[prefix + '.bn.bias'] = torch.zeros(c_out)
        sd[prefix + '.bn.running_mean'] = torch.zeros(c_out)
        sd[prefix + '.bn.running_var'] = torch.ones(c_out)
        sd[prefix + '.bn.num_batches_tracked'] = torch.tensor(0, dtype=torch.long)

    sites = conv_sites(cfg)
    n_down = 2 * cfg.depth + 2
    for i, (p, ci, co, _) in enumerate(sites[:n_down]):
        unit(p, ci, co)
        if cfg.residual and i % 2 == 1:   # ConvResidualBlock.__init__ creates `residual` after the block's convs (unet.py:54-55)
            conv(residual_prefix(cfg, i // 2), co, sites[i - 1][1], 1)
    k = n_down
    for j in range(cfg.depth):
        (p0, ci0, co0, _), (p1, ci1, co1, _) = sites[k], sites[k + 1]
        unit(p0, ci0, co0)
        unit(p1, ci1, co1)
        if cfg.residual:
            conv(residual_prefix(cfg, cfg.depth + 1 + j), co1, ci0, 1)
        conv('up_convs.%d.upconv.1' % j, co0, 2 * co0, 3)
        k += 2
    p, ci, co, _ = sites[k]
    unit(p, ci, co)
    conv('conv_cls.1', cfg.nb_classes, co, 1)
    if cfg.sigma_out:  # unet.py:162-164
        unit(SIGMA_PREFIX, co, co)
        conv('conv_sigma.1', cfg.nb_classes, co, 1)
    torch.random.set_rng_state(gen_state)

    # state_dict order of the reference module tree: inside UpConv, `block` is registered before `upconv`
    return sd


def randomize_statistics(sd, seed, logit_gain=24.0):
    """Give a random-init net non-degenerate behaviour (random-init probabilities sit at ~0.49 everywhere).

    BN affine/running stats are drawn from a seeded generator and the 1x1 head is scaled so logits spread
    over several units.  Pure test-data synthesis (no reference counterpart).
    """
    g = torch.Generator().manual_seed(seed)
    out = OrderedDict()
    for k, v in sd.items():
        if k.endswith('bn.weight'):
            out[k] = 0.75 + 0.5 * torch.rand(v.shape, generator=g)
        elif k.endswith('bn.bias'):
            out[k] = 0.1 * torch.randn(v.shape, generator=g)
        elif k.endswith('bn.running_mean'):
            out[k] = 0.1 * torch.randn(v.shape, generator=g)
        elif k.endswith('bn.running_var'):
            out[k] = 0.5 + torch.rand(v.shape, generator=g)
        elif k.startswith('conv_cls.1.weight'):
            out[k] = v * logit_gain
        else:
            out[k] = v.clone()
    return out


# ------------------------------------------------------------------------------------------------
# U-Net forward (common/model/unet.py:166-186) with optional injected Dropout2d keep-masks
# ------------------------------------------------------------------------------------------------
def _unit(x, sd, prefix, p_drop, keep, activation=True):
    y = F.conv2d(x, sd[prefix + '.conv.weight'], sd[prefix + '.conv.bias'], padding=1)
    if keep is not None:
        # nn.Dropout2d in train mode: per-(n, c) Bernoulli(1-p) noise divided by (1-p), multiplied in.
        noise = keep.to(y.dtype) / (1.0 - p_drop)
        y = y * noise[:, :, None, None]
    y = F.batch_norm(y, sd[prefix + '.bn.running_mean'], sd[prefix + '.bn.running_var'],
                     sd[prefix + '.bn.weight'], sd[prefix + '.bn.bias'], False, 0.0, BN_EPS)
    return F.relu(y) if activation else y


def unet_forward(sd, x, cfg, keep_masks=None, return_all=False):
    """Logits (N, nb_classes, H, W) of the reference UNet in eval mode.

    keep_masks: None (all Dropout2d in eval mode = identity) or a list, one entry per dropout site in
    forward order, of (N, C) {0,1} tensors = the Bernoulli keep decisions of Dropout2d in train mode.
    return_all: a dict with 'logits', 'features' (the conv_cls input, unet.py:178-179) and, for sigma_out nets,
    'sigma' (unet.py:185-186).
    """
    sites = conv_sites(cfg)
    masks = iter(keep_masks) if keep_masks is not None else None

    def run(x, idx, activation=True):
        prefix, _, _, has_do = sites[idx]
        keep = next(masks) if (masks is not None and has_do) else None
        return _unit(x, sd, prefix, cfg.dropout, keep, activation)

    def block(x, idx, number):
        """ConvBlock (unet.py:26-39) or ConvResidualBlock (:42-60): the last unit without ReLU, plus residual(x)."""
        y = run(x, idx)
        if not cfg.residual:
            return run(y, idx + 1)
        y = run(y, idx + 1, activation=False)
        rp = residual_prefix(cfg, number)
        return y + F.conv2d(x, sd[rp + '.weight'], sd[rp + '.bias'])

    skips = []
    idx = 0
    for lvl in range(cfg.depth):
        x = block(x, idx, lvl)
        idx += 2
        skips.append(x)
        x = F.max_pool2d(x, 2)
    x = block(x, idx, cfg.depth)
    idx += 2
    for j in range(cfg.depth):
        skip = skips[-(j + 1)]
        up = F.interpolate(x, scale_factor=2, mode='nearest')
        up = F.conv2d(up, sd['up_convs.%d.upconv.1.weight' % j], sd['up_convs.%d.upconv.1.bias' % j], padding=1)
        if up.shape[-2:] != skip.shape[-2:]:
            dy = skip.shape[-2] - up.shape[-2]
            dx = skip.shape[-1] - up.shape[-1]
            up = F.pad(up, (dx // 2, dx // 2 + dx % 2, dy // 2, dy // 2 + dy % 2))
        x = torch.cat((up, skip), 1)
        x = block(x, idx, cfg.depth + 1 + j)
        idx += 2
    features = x
    x = run(features, idx)
    logits = F.conv2d(x, sd['conv_cls.1.weight'], sd['conv_cls.1.bias'])
    if not return_all:
        return logits
    out = {'logits': logits, 'features': features}
    if cfg.sigma_out:
        keep = next(masks) if (masks is not None and cfg.dropout is not None) else None
        y = _unit(features, sd, SIGMA_PREFIX, cfg.dropout, keep)
        out['sigma'] = F.conv2d(y, sd['conv_sigma.1.weight'], sd['conv_sigma.1.bias'])
    return out


def predict_deterministic(sd, images, cfg):
    """SegmentationPredictStep(do_probs=True) (common/trainloop/steps.py:69-90)."""
    logits = unet_forward(sd, images.float(), cfg)
    return {'logits': logits, 'probabilities': F.softmax(logits, 1)}


def predict_mc(sd, images, cfg, mc_steps, keep_masks_per_step):
    """McPredictStep (rechun/dl/customsteps.py:16-39) with the dropout decisions supplied by the caller.

    keep_masks_per_step[t] is the per-site mask list of sample t (see unet_forward).
    """
    images = images.float()
    out = {'ws_probabilities': F.softmax(unet_forward(sd, images, cfg), 1)}
    probs = []
    for t in range(mc_steps):
        probs.append(F.softmax(unet_forward(sd, images, cfg, keep_masks_per_step[t]), 1))
    out['multi_probabilities'] = torch.stack(probs)
    return out


def predict_ensemble(state_dicts, images, cfg):
    """EnsemblePredictionStep (bin-dl/brats_test_ensemble.py:78-94)."""
    images = images.float()
    return {'multi_probabilities': torch.stack([F.softmax(unet_forward(sd, images, cfg), 1) for sd in state_dicts])}


def predict_aleatoric(sd, images, cfg, is_log_sigma=False):
    """AleatoricPredictStep (bin-dl/brats_test_aleatoric.py:51-73)."""
    out = unet_forward(sd, images.float(), cfg, return_all=True)
    sigma = out['sigma'].exp() if is_log_sigma else out['sigma'].abs()
    return {'logits': out['logits'], 'sigma': sigma, 'probabilities': F.softmax(out['logits'], 1)}


def postnet_init_state_dict(in_channels, nb_classes, nb_convs, seed):
    """`torch.manual_seed(seed); PostNet(in_channels, nb_classes, nb_convs).state_dict()` (common/model/postnet.py:8-12)."""
    gen_state = torch.random.get_rng_state()
    torch.manual_seed(seed)
    sd = OrderedDict()

    def conv(prefix, c_out, c_in):
        w = torch.empty(c_out, c_in, 1, 1)
        torch.nn.init.kaiming_uniform_(w, a=math.sqrt(5))
        bound = 1 / math.sqrt(c_in)
        b = torch.empty(c_out)
        torch.nn.init.uniform_(b, -bound, bound)
        sd[prefix + '.weight'] = w
        sd[prefix + '.bias'] = b
    for i in range(nb_convs):
        p = 'convs.%d.conv2d_batch_relu' % i
        conv(p + '.conv', in_channels, in_channels)
        sd[p + '.bn.weight'] = torch.ones(in_channels)
        sd[p + '.bn.bias'] = torch.zeros(in_channels)
        sd[p + '.bn.running_mean'] = torch.zeros(in_channels)
        sd[p + '.bn.running_var'] = torch.ones(in_channels)
        sd[p + '.bn.num_batches_tracked'] = torch.tensor(0, dtype=torch.long)
    conv('conv_logits', nb_classes, in_channels)
    torch.random.set_rng_state(gen_state)
    return sd


def postnet_forward(sd, x):
    """PostNet.forward in eval mode (common/model/postnet.py:14-17)."""
    i = 0
    while 'convs.%d.conv2d_batch_relu.conv.weight' % i in sd:
        p = 'convs.%d.conv2d_batch_relu' % i
        x = F.conv2d(x, sd[p + '.conv.weight'], sd[p + '.conv.bias'])
        x = F.batch_norm(x, sd[p + '.bn.running_mean'], sd[p + '.bn.running_var'], sd[p + '.bn.weight'], sd[p + '.bn.bias'],
                         False, 0.0, BN_EPS)
        x = F.relu(x)
        i += 1
    return F.conv2d(x, sd['conv_logits.weight'], sd['conv_logits.bias'])


def predict_aux_feat(sd_unet, cfg, sd_postnet, images):
    """SegmentationPredictStep(test_model) of bin-dl/brats_test_auxiliary_feat.py:61-80."""
    out = unet_forward(sd_unet, images.float(), cfg, return_all=True)
    return {'segm_probabilities': F.softmax(out['logits'], 1), 'features': out['features'],
            'probabilities': F.softmax(postnet_forward(sd_postnet, out['features']), 1)}


def predict_aux_segm(sd, images, labels, cfg):
    """SegmentationPredictStep of bin-dl/brats_test_auxiliary_segm.py:48-69 (cfg.in_channels = image channels + 1)."""
    pred = labels.long()[:, 1]
    inpt = torch.cat([images.float(), pred.unsqueeze(1).float()], dim=1)
    logits = unet_forward(sd, inpt, cfg)
    return {'logits': logits, 'probabilities': F.softmax(logits, 1), 'orig_prediction': pred.unsqueeze(1)}


def torch_entropy(p, dim=-1, keepdim=False):
    """th.entropy (common/utils/torchhelper.py:53-54): natural log, 0·log0 := 0."""
    return -torch.where(p > 0, p * p.log(), torch.zeros((), dtype=p.dtype)).sum(dim=dim, keepdim=keepdim)


def summarize(multi_probabilities, do_mi=False, do_var=False):
    """MultiPredictionSummary (rechun/dl/customsteps.py:50-71)."""
    out = {}
    probabilities = multi_probabilities.mean(dim=0)
    out['probabilities'] = probabilities
    entropy = torch_entropy(probabilities, dim=1, keepdim=True)
    out['entropy'] = entropy
    if do_mi:
        out['mutual_info'] = entropy - torch_entropy(multi_probabilities, dim=2, keepdim=True).mean(dim=0)
    if do_var:
        out['variance'] = multi_probabilities.var(dim=0).mean(dim=1, keepdim=True)
    return out


# ------------------------------------------------------------------------------------------------
# Philox4x32-10 keep-mask stream (new in the build; host restatement of csrc/philox.cuh)
# ------------------------------------------------------------------------------------------------
_PHILOX_M0 = np.uint64(0xD2511F53)
_PHILOX_M1 = np.uint64(0xCD9E8D57)
_PHILOX_W0 = 0x9E3779B9
_PHILOX_W1 = 0xBB67AE85
_MASK32 = np.uint64(0xFFFFFFFF)


def philox4x32_10(counter, key):
    """Vectorised Philox4x32-10.  counter: (..., 4) uint32, key: (2,) ints.  Returns (..., 4) uint32."""
    c = [np.asarray(counter[..., i], dtype=np.uint64) for i in range(4)]
    k0, k1 = int(key[0]) & 0xFFFFFFFF, int(key[1]) & 0xFFFFFFFF
    for _ in range(10):
        p0 = _PHILOX_M0 * c[0]
        p1 = _PHILOX_M1 * c[2]
        hi0, lo0 = p0 >> np.uint64(32), p0 & _MASK32
        hi1, lo1 = p1 >> np.uint64(32), p1 & _MASK32
        c = [hi1 ^ c[1] ^ np.uint64(k0), lo1, hi0 ^ c[3] ^ np.uint64(k1), lo0]
        k0 = (k0 + _PHILOX_W0) & 0xFFFFFFFF
        k1 = (k1 + _PHILOX_W1) & 0xFFFFFFFF
    return np.stack(c, axis=-1).astype(np.uint32)


def dropout_threshold(p):
    """keep  <=>  u32 >= ceil(p * 2**32)   (u32 uniform on [0, 2**32))."""
    return min(0xFFFFFFFF, int(math.ceil(p * 4294967296.0)))


def philox_keep_masks(cfg, seed, sample, slice_index0, n_slices):
    """Keep decisions of MC sample `sample` for slices [slice_index0, slice_index0+n_slices).

    Stream definition (must match csrc/philox.cuh): for dropout site s, slice g (a run-global index so the
    stream is independent of batching and of the GPU a slice lands on), sample t and channel c:
        r = philox4x32_10(counter=(c // 4, s, g, t), key=(seed_lo, seed_hi));  keep = r[c % 4] >= thr(p)
    Returns a list (one per site) of (n_slices, C) uint8 tensors.
    """
    thr = dropout_threshold(cfg.dropout)
    key = (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    masks = []
    for s, (_, ch) in enumerate(dropout_sites(cfg)):
        groups = (ch + 3) // 4
        ctr = np.zeros((n_slices, groups, 4), dtype=np.uint32)
        ctr[..., 0] = np.arange(groups, dtype=np.uint32)[None, :]
        ctr[..., 1] = s
        ctr[..., 2] = (slice_index0 + np.arange(n_slices, dtype=np.uint32))[:, None]
        ctr[..., 3] = sample
        r = philox4x32_10(ctr, key).reshape(n_slices, groups * 4)[:, :ch]
        masks.append(torch.from_numpy((r >= np.uint32(thr)).astype(np.uint8)))
    return masks


# ------------------------------------------------------------------------------------------------
# Calibration: ECE reliability bins (common/evalutation/numpyfunctions.py:6-83)
# ------------------------------------------------------------------------------------------------
def calibration_edges(n_bins=10):
    return np.linspace(0., 1. + 1e-8, n_bins + 1)  # numpyfunctions.py:53 (float64)


def calibration_tables(probabilities, target, n_bins=10, threshold_range=None, mask=None):
    """Un-compacted per-bin tables: (count int64[n], positives float64[n], confidence_sum float64[n]).

    Follows binary_calibration/_binary_calibration (numpyfunctions.py:26-69) up to the three bincounts.
    Values outside [0, 1+1e-8) make np.bincount return more than n_bins entries in the reference; the
    tables returned here then have that longer length too (callers decide the policy).
    """
    if probabilities.ndim > target.ndim:
        if probabilities.shape[-1] > 2:
            raise ValueError('can only evaluate the calibration for binary classification')
        elif probabilities.shape[-1] == 2:
            probabilities = probabilities[..., 1]
        else:
            probabilities = np.squeeze(probabilities, axis=-1)
    if mask is not None:
        probabilities, target = probabilities[mask], target[mask]
    if threshold_range is not None:
        lo, hi = threshold_range
        keep = np.logical_and(probabilities < hi, probabilities > lo)
        probabilities, target = probabilities[keep], target[keep]
    p = probabilities.flatten()
    t = target.flatten()
    ids = np.digitize(p, calibration_edges(n_bins)) - 1
    conf_sum = np.bincount(ids, weights=p, minlength=n_bins)
    positives = np.bincount(ids, weights=t, minlength=n_bins)
    count = np.bincount(ids, minlength=n_bins)
    return count, positives, conf_sum


def ece_from_tables(count, positives, conf_sum, bin_weighting='proportion', n_dim=3):
    """ECE and the compacted bin arrays from the three tables (numpyfunctions.py:6-23,65-83)."""
    non_zero = count != 0
    pos_frac = positives[non_zero] / count[non_zero]
    mean_conf = conf_sum[non_zero] / count[non_zero]
    cnt = count[non_zero]
    if bin_weighting == 'proportion':
        w = cnt / cnt.sum()
    elif bin_weighting == 'log_proportion':
        w = np.log(cnt) / np.log(cnt).sum()
    elif bin_weighting == 'power_proportion':
        w = cnt ** (1 / n_dim) / (cnt ** (1 / n_dim)).sum()
    elif bin_weighting == 'mean_proportion':
        w = 1 / non_zero.sum()
    else:
        raise ValueError('unknown bin weighting "{}"'.format(bin_weighting))
    ece = (np.abs(mean_conf - pos_frac) * w).sum()
    return ece, {'bins_count': cnt, 'bins_avg_confidence': mean_conf, 'bins_positive_fraction': pos_frac,
                 'bins_non_zero': non_zero}


def ece_binary(probabilities, target, n_bins=10, threshold_range=None, mask=None, bin_weighting='proportion'):
    count, positives, conf_sum = calibration_tables(probabilities, target, n_bins, threshold_range, mask)
    return ece_from_tables(count, positives, conf_sum, bin_weighting, target.ndim)


# ------------------------------------------------------------------------------------------------
# Eval-side preparation (rechun/eval/helper.py:25-47, rechun/eval/analysis.py:147-151,189-203)
# ------------------------------------------------------------------------------------------------
def add_background_probability(p_foreground):
    if p_foreground.max() > 1:
        raise ValueError('Found value larger than 1: "{}"'.format(p_foreground.max()))
    if p_foreground.min() < 0:
        raise ValueError('Found value smaller than 0: "{}"'.format(p_foreground.min()))
    return np.stack([1 - p_foreground, p_foreground], axis=-1)


def numpy_entropy(p, axis=-1):
    # numpyfunctions.py:166-168 — the float list literal makes np.where promote to float64 before the sum
    with np.errstate(divide='ignore', invalid='ignore'):  # log(0) and 0*inf of the discarded where-branch
        return -np.where(p > 0, p * np.log(p), [0.0]).sum(axis=axis)


def normalized_entropy(prob_2class):
    """ToEntropy (analysis.py:189-203): H(prob)/ln 2, float64."""
    return numpy_entropy(prob_2class) / np.log(2)


# ------------------------------------------------------------------------------------------------
# Confusion / Dice / accuracy (numpyfunctions.py:128-151 via pymia 0.2.1 metric classes)
# ------------------------------------------------------------------------------------------------
def boarder_mask(binary_label_map, distance_in=1, distance_out=1):
    """common/utils/labelhelper.py:12-20: voxels within `distance_in` of the background (inside the object) and within
    `distance_out` of the object (outside).  Returns (dist_in + dist_out, mask)."""
    import scipy.ndimage
    m = np.asarray(binary_label_map).astype(bool)
    dist_in = scipy.ndimage.distance_transform_edt(m)
    dist_out = scipy.ndimage.distance_transform_edt(~m)
    return dist_in + dist_out, (dist_in <= distance_in) * (dist_out <= distance_out)


def rescale_uncertainties(uncertainty, min_, max_, epsilon=1e-5):
    """rechun/eval/helper.py:18-21."""
    rescaled = (uncertainty - min_) / (max_ - min_)
    return rescaled * (1 - 2 * epsilon) + epsilon


def uncertainty_to_foreground_probabilities(uncertainty, prediction):
    """rechun/eval/helper.py:7-15."""
    if prediction.shape != uncertainty.shape:
        raise ValueError('shapes must agree. Found {} and {}'.format(uncertainty.shape, prediction.shape))
    if uncertainty.max() > 1 or uncertainty.min() < 0 or prediction.max() > 1:
        raise ValueError('values outside the valid range')
    p = uncertainty * 0.5
    p[prediction == 1] = 1 - p[prediction == 1]
    return p


def confusion(prediction, target):
    tp = np.sum(np.logical_and(prediction == 1, target == 1))
    tn = np.sum(np.logical_and(prediction == 0, target == 0))
    fp = np.sum(np.logical_and(prediction == 1, target == 0))
    fn = np.sum(np.logical_and(prediction == 0, target == 1))
    return tp, tn, fp, fn, prediction.size


def dice_from_counts(tp, fp, fn):
    if tp == 0 and (tp + fp + fn) == 0:
        return 1.
    return 2 * tp / (2 * tp + fp + fn)


def accuracy_from_counts(tp, tn, fp, fn):
    s = tp + tn + fp + fn
    return (tp + tn) / s if s != 0 else 0


def dice(prediction, target):
    tp, tn, fp, fn, _ = confusion(prediction, target)
    return dice_from_counts(tp, fp, fn)


def accuracy(prediction, target):
    tp, tn, fp, fn, _ = confusion(prediction, target)
    return accuracy_from_counts(tp, tn, fp, fn)


# ------------------------------------------------------------------------------------------------
# Uncertainty-error overlap (numpyfunctions.py:86-125, eval.py:145-226)
# ------------------------------------------------------------------------------------------------
def uncertainty_counts(prediction, target, thresholded_uncertainty, mask=None):
    """tp, tn, fp, fn and their intersections with the thresholded-uncertainty map (numpyfunctions.py:86-107)."""
    if mask is not None:
        prediction, target, thresholded_uncertainty = prediction[mask], target[mask], thresholded_uncertainty[mask]
    classes = (np.logical_and(target, prediction), np.logical_and(~target, ~prediction),
               np.logical_and(~target, prediction), np.logical_and(target, ~prediction))
    with_u = tuple(np.logical_and(c, thresholded_uncertainty).sum() for c in classes)
    plain = tuple(c.sum() for c in classes)
    return plain + with_u  # tp, tn, fp, fn, tpu, tnu, fpu, fnu


def error_dice(fp, fn, tpu, tnu, fpu, fnu):
    if (fnu + fpu) == 0 and (fn + fp + fnu + fpu + tnu + tpu) == 0:
        return 1.
    return (2 * (fnu + fpu)) / (fn + fp + fnu + fpu + tnu + tpu)


def error_recall(fp, fn, fpu, fnu):
    if (fnu + fpu) == 0 and (fn + fp) == 0:
        return 1.
    return (fnu + fpu) / (fn + fp)


def error_precision(tpu, tnu, fpu, fnu):
    if (fnu + fpu) == 0 and (fnu + fpu + tpu + tnu) == 0:
        return 1.
    return (fnu + fpu) / (fnu + fpu + tpu + tnu)


def uncertainty_error_dice(prediction, target, uncertainty, threshold, prefix='', target_boarder=None):
    """UncertaintyErrorDiceNumpy.__call__ (eval.py:156-173)."""
    target = target.astype(bool)
    prediction = prediction.astype(bool)
    mask = None if target_boarder is None else ~target_boarder
    tp, tn, fp, fn, tpu, tnu, fpu, fnu = uncertainty_counts(prediction, target, uncertainty > threshold, mask)
    return {prefix + 'precision': error_precision(tpu, tnu, fpu, fnu), prefix + 'recall': error_recall(fp, fn, fpu, fnu),
            prefix + 'dice': error_dice(fp, fn, tpu, tnu, fpu, fnu)}


def uncertainty_and_correction(prediction, target, uncertainty, threshold):
    """UncertaintyAndCorrectionEvalNumpy.__call__ (eval.py:182-226), materialising the corrected maps."""
    target = target.astype(bool)
    prediction = prediction.astype(bool)
    thr_u = uncertainty > threshold
    tp, tn, fp, fn, tpu, tnu, fpu, fnu = uncertainty_counts(prediction, target, thr_u)
    r = {'tpu': tpu, 'tnu': tnu, 'fpu': fpu, 'fnu': fnu, 'tp': tp, 'tn': tn, 'fp': fp, 'fn': fn}
    with np.errstate(divide='ignore', invalid='ignore'):
        ratio = r['tpu'] / r['fpu']
        jaccard = r['tp'] / (r['tp'] + r['fp'] + r['fn'])
    r['dice_benefit'] = ratio < jaccard
    r['accuracy_benefit'] = ratio < 1
    r['dice'] = dice(prediction, target)
    r['accuracy'] = accuracy(prediction, target)
    corrected = prediction.copy()
    corrected[thr_u] = 0
    r['corrected_dice'] = dice(corrected, target)
    r['corrected_accuracy'] = accuracy(corrected, target)
    r['dice_benefit_correct'] = (r['corrected_dice'] > r['dice']) == r['dice_benefit']
    r['accuracy_benefit_correct'] = (r['corrected_accuracy'] > r['accuracy']) == r['accuracy_benefit']
    corrected = prediction.copy()
    corrected[thr_u] = 1
    r['corrected_add_dice'] = dice(corrected, target)
    r['corrected_add_accuracy'] = accuracy(corrected, target)
    return r


def sweep(prediction, target, uncertainty, thresholds=SWEEP_THRESHOLDS):
    """CorrectionAction (bin-eval/eval_uncertainty.py:195-202): one result dict per threshold."""
    return [uncertainty_and_correction(prediction, target, uncertainty, th) for th in thresholds]


def ue_table_row(r):
    """bin-analysis/table_ece_ue_bnf_dice.py:56-59 per-subject derived columns."""
    denom = r['fn'] + r['fp'] + r['fnu'] + r['fpu'] + r['tnu'] + r['tpu']
    return {'benefit': r['corrected_dice'] > r['dice'],
            'ue': (2 * (r['fnu'] + r['fpu'])) / denom if denom else float('nan')}


# ------------------------------------------------------------------------------------------------
# Report-side reductions
# ------------------------------------------------------------------------------------------------
def bins_csv_row(results):
    """WriteBinsCsvHook.on_subject + WriteCsvHook._unfold_results (rechun/eval/hook.py:49-62,75-93) -> {column: value}."""
    nz = np.asarray(results['bins_non_zero'])
    row = {}
    for key, value in results.items():
        if key in ('bins_count', 'bins_avg_confidence', 'bins_positive_fraction'):
            full = np.zeros_like(nz, dtype=np.asarray(value).dtype)
            full[nz] = value
            value = full
        if isinstance(value, np.ndarray):
            value = value.tolist()
        if isinstance(value, (list, tuple)):
            digits = len(str(len(value)))
            for i, v in enumerate(value):
                row['{}_{:0{}d}'.format(key, i, digits)] = v
        else:
            row[key] = value
    return row


def dataset_vs_mean_subject_ece(rows):
    """table_supplmat_ece_dataset_vs_meansubject.py:59-104 on the unfolded calibration rows of one test id."""
    def block(prefix, dtype):
        return np.array([[r['%s_%02d' % (prefix, b)] for b in range(10)] for r in rows], dtype=dtype)
    nonzero = block('bins_non_zero', bool)
    conf = np.ma.array(block('bins_avg_confidence', np.float64), mask=~nonzero)
    pos = np.ma.array(block('bins_positive_fraction', np.float64), mask=~nonzero)
    cnt = np.ma.array(block('bins_count', np.int64), mask=~nonzero)
    bin_sum = cnt.sum(axis=0)
    avg_conf = (conf * cnt).sum(axis=0) / bin_sum
    pos_frac = (pos * cnt).sum(axis=0) / bin_sum
    ece = (np.abs(conf - pos) * (cnt / cnt.sum(axis=1, keepdims=True))).sum(axis=1)
    assert np.allclose(ece.data, [r['ece'] for r in rows])      # the reference's own consistency check (:77-78)
    return {'ece': ece.mean(), 'ds_ece': (np.abs(avg_conf - pos_frac) * bin_sum / bin_sum.sum()).sum()}


def best_threshold_summary(sweeps, ece, dice, thresholds=SWEEP_THRESHOLDS):
    """table_ece_ue_bnf_dice.py:30-73,132-143 for one test id.  sweeps[s][threshold] = uncertainty_and_correction()."""
    table = {}
    for name in ('benefit', 'error'):
        per_th = []
        for th in thresholds:
            vals = []
            for s in sweeps:
                r = s[th]
                if name == 'benefit':
                    vals.append(float((r['corrected_dice'] - r['dice']) > 0))
                else:
                    with np.errstate(divide='ignore', invalid='ignore'):
                        vals.append((2 * (r['fnu'] + r['fpu'])) / np.float64(r['fn'] + r['fp'] + r['fnu'] + r['fpu'] + r['tnu'] + r['tpu']))
            vals = np.asarray(vals, dtype=np.float64)             # DataFrame.mean skips NaN subjects (0 / 0 U-E Dice)
            per_th.append(np.nan if np.all(np.isnan(vals)) else np.nanmean(vals))
        per_th = np.array(per_th)
        k = int(np.nanargmax(per_th))                            # Series.idxmax: first maximum, NaN skipped
        table[name], table[name + '_threshold'] = per_th[k], float(thresholds[k])
    table['ece'], table['dice'] = float(np.mean(ece)), float(np.mean(dice))
    return table
