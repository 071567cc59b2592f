"""Synthetic weights for benchmarks, smoke runs and profiling drivers (there is no network for checkpoints).

`random_unet_state_dict` returns a state dict with the key layout of the reference `UNet`
(common/model/unet.py:134-164; `model.unit_layout` names the units): convolutions get torch's default nn.Conv2d
initialisation from a seeded generator, BatchNorm affine / running statistics are drawn non-degenerate and the 1x1 head
is scaled so that the logits spread over several units (a plain random init predicts p ~ 0.49 everywhere, which would
put every voxel into one reliability bin).  Pure data synthesis: nothing here is reference arithmetic.
"""
import math
from collections import OrderedDict

import torch

from .model import unit_layout


def _conv(sd, prefix, c_out, c_in, k, g):
    fan_in = c_in * k * k
    bound_w = math.sqrt(6.0 / ((1 + 5.0) * fan_in))          # kaiming_uniform_(a=sqrt(5))
    sd[prefix + '.weight'] = (torch.rand((c_out, c_in, k, k), generator=g) * 2 - 1) * bound_w
    sd[prefix + '.bias'] = (torch.rand((c_out,), generator=g) * 2 - 1) / math.sqrt(fan_in)


def _unit(sd, prefix, c_in, c_out, k, g):
    _conv(sd, prefix + '.conv', c_out, c_in, k, g)
    sd[prefix + '.bn.weight'] = 0.75 + 0.5 * torch.rand((c_out,), generator=g)
    sd[prefix + '.bn.bias'] = 0.1 * torch.randn((c_out,), generator=g)
    sd[prefix + '.bn.running_mean'] = 0.1 * torch.randn((c_out,), generator=g)
    sd[prefix + '.bn.running_var'] = 0.5 + torch.rand((c_out,), generator=g)
    sd[prefix + '.bn.num_batches_tracked'] = torch.tensor(0, dtype=torch.long)


def random_unet_state_dict(in_channels=4, depth=4, start_filters=32, nb_classes=2, seed=20, sigma_out=False, logit_gain=24.0):
    g = torch.Generator().manual_seed(seed)
    sd = OrderedDict()
    units, upconvs = unit_layout(in_channels, depth, start_filters, 0.05, None)
    for prefix, c_in, c_out, _ in units:
        _unit(sd, prefix, c_in, c_out, 3, g)
    for prefix, c_in, c_out in upconvs:
        _conv(sd, prefix, c_out, c_in, 3, g)
    _conv(sd, 'conv_cls.1', nb_classes, start_filters, 1, g)
    sd['conv_cls.1.weight'] *= logit_gain
    if sigma_out:
        _unit(sd, 'conv_sigma.0.conv2d_batch_relu', start_filters, start_filters, 3, g)
        _conv(sd, 'conv_sigma.1', nb_classes, start_filters, 1, g)
    return sd


def random_postnet_state_dict(in_channels=32, nb_classes=2, nb_convs=3, seed=21, logit_gain=4.0):
    g = torch.Generator().manual_seed(seed)
    sd = OrderedDict()
    for i in range(nb_convs):
        _unit(sd, 'convs.%d.conv2d_batch_relu' % i, in_channels, in_channels, 1, g)
    _conv(sd, 'conv_logits', nb_classes, in_channels, 1, g)
    sd['conv_logits.weight'] *= logit_gain
    return sd
