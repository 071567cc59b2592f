"""GPU parity tests of the tcgen05 U-Net forward against the reference (golden fixtures made by the unmodified
reference) and against the oracle restatement at BraTS / ISIC sizes.

Stated tolerance (bf16 operands, fp32 accumulation, fp32 reference): for nets whose logits spread over several
units (std ~1, |max| ~3-4) the foreground probability differs from the fp32 reference by at most 1.2e-2 (measured
1e-3 ... 8e-3) and by less than 2e-3 on average; logits by at most 0.08.  Integer-exact properties (chunking / batching invariance,
injected == generated masks, tcgen05 == cross-check up to one bf16 ulp) are asserted exactly.
"""
import numpy as np
import pytest
import torch

from rcu_b200 import metrics, model
from oracle import restate as R
from helpers import GOLDEN_CONFIGS

pytestmark = pytest.mark.gpu
torch.set_grad_enabled(False)
P_MAX, P_MEAN, LOGIT_MAX = 1.2e-2, 2e-3, 0.08


def _net(name_or_kw, seed=20, **extra):
    kw = GOLDEN_CONFIGS[name_or_kw] if isinstance(name_or_kw, str) else name_or_kw
    cfg = R.UNetConfig(**kw)
    sd = R.randomize_statistics(R.init_state_dict(cfg, seed), 7)
    net = model.B200UNet(sd, in_channels=cfg.in_channels, dropout=cfg.dropout, dropout_center=cfg.dropout_center, **extra)
    return cfg, sd, net


def _close(logits_nchw, ref_logits):
    got = logits_nchw.cpu()
    assert got.shape == ref_logits.shape and got.dtype == torch.float32
    assert (got - ref_logits).abs().max().item() <= LOGIT_MAX
    dp = (torch.softmax(got, 1) - torch.softmax(ref_logits, 1)).abs()
    assert dp.max().item() <= P_MAX and dp.mean().item() <= P_MEAN


@pytest.mark.parametrize('name', sorted(GOLDEN_CONFIGS))
def test_deterministic_and_mc_logits_match_reference_golden(golden_unet, name):
    cfg, sd, net = _net(name)
    x = torch.from_numpy(golden_unet[name + '/input'])
    _close(net(x.cuda()), torch.from_numpy(golden_unet[name + '/logits']))
    # MC: the golden multi_probabilities were produced by the reference with the Philox keep masks injected
    ref_multi = torch.from_numpy(golden_unet[name + '/multi_probabilities'])
    T = ref_multi.shape[0]
    logits = net.forward_samples(x, T + 1, dropout_mode=1, det_first=True, seed=20, slice_index0=0, sample0=0)
    probs = torch.softmax(logits.permute(0, 1, 4, 2, 3), 2).cpu()
    dp = (probs[1:] - ref_multi).abs()
    assert dp.max().item() <= P_MAX and dp.mean().item() <= P_MEAN
    assert (probs[0] - torch.from_numpy(golden_unet[name + '/ws_probabilities'])).abs().max().item() <= P_MAX
    # masks really matter: a different seed moves the samples well beyond the tolerance
    other = net.forward_samples(x, T + 1, dropout_mode=1, det_first=True, seed=21)
    assert (torch.softmax(other[1:].permute(0, 1, 4, 2, 3), 2).cpu() - ref_multi).abs().max().item() > 2 * P_MAX
    assert torch.equal(other[0], logits[0])


def test_reference_supplied_masks_equal_generated_masks():
    cfg, sd, net = _net('brats')
    x = torch.randn(3, 4, 48, 64)
    a = net.forward_samples(x, 4, dropout_mode=1, det_first=True, seed=20, slice_index0=11, sample0=2)
    scale = metrics.philox_keep_scale_host(20, cfg.dropout, net.site_channels, 11, 3, 2, 3)
    b = net.forward_samples(x, 4, dropout_mode=2, det_first=True, scale=scale)
    assert torch.equal(a, b)
    # arbitrary caller-made decisions (a torch Bernoulli draw, like nn.Dropout2d's) follow the oracle with those masks
    g = torch.Generator().manual_seed(5)
    keep = [(torch.rand(3, c, generator=g) >= cfg.dropout).to(torch.uint8) for c in net.site_channels]
    scale = torch.cat(keep, 1).float()[None] / (1 - cfg.dropout)
    got = net.forward_samples(x, 1, dropout_mode=2, scale=scale)[0].permute(0, 3, 1, 2)
    _close(got, R.unet_forward(sd, x, cfg, keep))
    with pytest.raises(ValueError):
        net.forward_samples(x, 1, dropout_mode=2, scale=scale[:, :2])


def test_tcgen05_path_matches_cuda_core_cross_check_layer_by_layer():
    cfg, sd, net_tc = _net('brats', chunk_images=64)
    _, _, net_ck = _net('brats', chunk_images=64)
    net_ck.set_conv_impl(1)
    net_tc.set_first_layer_dedup(False)   # activation 0 per (sample, slice), as the cross-check path stores it (the dedup has its own test)
    n, h, w = 3, 48, 64
    x = torch.randn(n, 4, h, w)
    out_tc = net_tc.forward_samples(x, 2, dropout_mode=1, det_first=True, seed=5)
    out_ck = net_ck.forward_samples(x, 2, dropout_mode=1, det_first=True, seed=5)
    dims = [(h, w, 32), (h, w, 32)]
    c = 32
    for l in range(1, 5):
        dims += [(h >> l, w >> l, c), (h >> l, w >> l, 2 * c), (h >> l, w >> l, 2 * c)]
        c *= 2
    for l in range(3, -1, -1):
        c //= 2
        dims += [(h >> l, w >> l, c)] * 3
    for i, (hh, ww, cc) in enumerate(dims):
        a = net_ck.debug_activation(i, (2 * n, hh, ww, cc))
        b = net_tc.debug_activation(i, (2 * n, hh, ww, cc))
        # identical bf16 inputs, fp32 accumulation in a different order: at most one bf16 ulp apart
        assert (a - b).abs().max().item() <= 2 ** -7 * max(a.abs().max().item(), 1.0), i
    assert (out_tc - out_ck).abs().max().item() <= 0.05


@pytest.mark.parametrize('shape,in_ch', [((240, 240), 4), ((256, 256), 3), ((192, 256), 3), ((16, 16), 4), ((16, 48), 4)])
def test_full_size_slices_match_oracle(shape, in_ch):
    cfg, sd, net = _net(dict(in_channels=in_ch))
    n = 3 if shape[0] >= 192 else 5
    g = torch.Generator().manual_seed(shape[0] + in_ch)
    x = torch.randn(n, in_ch, *shape, generator=g)
    _close(net(x.cuda()), R.unet_forward(sd, x, cfg))


def test_chunking_and_batching_do_not_change_results():
    cfg, sd, big = _net('brats', chunk_images=64)
    _, _, small = _net('brats', chunk_images=8)
    x = torch.randn(7, 4, 32, 48)
    a = big.forward_samples(x, 4, dropout_mode=1, det_first=True, seed=3, slice_index0=100)
    b = small.forward_samples(x, 4, dropout_mode=1, det_first=True, seed=3, slice_index0=100)   # 2 slices per chunk
    assert torch.equal(a, b)
    # the Philox stream is keyed by the run-global slice index: two calls == one call
    c = torch.cat([big.forward_samples(x[:3], 4, dropout_mode=1, det_first=True, seed=3, slice_index0=100),
                   big.forward_samples(x[3:], 4, dropout_mode=1, det_first=True, seed=3, slice_index0=103)], dim=1)
    assert torch.equal(a, c)
    # sample sharding: samples [0,3) == samples [0,1) + [1,3)
    d = torch.cat([big.forward_samples(x, 1, dropout_mode=1, seed=3, slice_index0=100, sample0=0),
                   big.forward_samples(x, 2, dropout_mode=1, seed=3, slice_index0=100, sample0=1)], dim=0)
    assert torch.equal(a[1:], d)


def test_dropout_switch_follows_set_dropout_mode_semantics():
    cfg, sd, net = _net('brats')
    x = torch.randn(2, 4, 32, 32).cuda()
    net.eval()
    det = net(x)
    assert torch.equal(det, net(x))
    for m in net.modules():                      # th.set_dropout_mode(model, True)
        if isinstance(m, torch.nn.Dropout2d):
            m.train()
    net.reset_stream(seed=20)
    s0, s1 = net(x), net(x)
    assert not torch.equal(s0, s1) and not torch.equal(s0, det)
    folded = net.forward_samples(x, 2, dropout_mode=1, seed=20)
    assert torch.equal(torch.stack([s0, s1]), folded.permute(0, 1, 4, 2, 3))
    for m in net.modules():
        if isinstance(m, torch.nn.Dropout2d):
            m.eval()
    assert torch.equal(det, net(x))


def test_unsupported_configurations_fail_loudly():
    cfg, sd, net = _net('brats')
    with pytest.raises(NotImplementedError):
        net(torch.randn(1, 4, 40, 40).cuda())          # not a multiple of 2^depth: F.pad branch is off the hot path
    with pytest.raises(ValueError):
        net(torch.randn(1, 3, 32, 32).cuda())
    with pytest.raises(NotImplementedError):
        model.B200UNet(sd, nb_classes=3)
    bad = dict(sd)
    bad['conv_sigma.1.weight'] = torch.zeros(2, 32, 1, 1)     # a sigma head without its conv_sigma.0 unit
    with pytest.raises(ValueError):
        model.B200UNet(bad)
    bad = dict(sd)
    bad['down_convs.0.block.residual.weight'] = torch.zeros(32, 4, 1, 1)
    with pytest.raises(ValueError):
        model.B200UNet(bad)                                    # a residual=True net with the other blocks' residual convs missing
    with pytest.raises(ValueError):
        net.forward_outputs(torch.randn(1, 4, 32, 32), sigma=True)   # sigma requested from a sigma_out=False net
    missing = {k: v for k, v in sd.items() if 'bottom_convs.block.1' not in k}
    with pytest.raises(ValueError):
        model.B200UNet(missing)


def test_free_running_mc_is_statistically_equivalent_to_torch_dropout():
    """Free-running MC (our Philox stream vs torch's Bernoulli stream) cannot agree mask for mask; the summary
    statistics must: mean predictive entropy and ECE of our T=20 run lie within 4 sigma of the seed-to-seed spread
    of the oracle run with torch-drawn masks (R = 8 seeds)."""
    cfg, sd, net = _net(dict(in_channels=4, dropout=0.2))
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 4, 64, 64, generator=g)
    target = (torch.rand(2, 64, 64, generator=g) < 0.4).numpy().astype(np.uint8)
    T = 20

    def stats(mean_p):
        ent = R.torch_entropy(mean_p, 1).mean().item()
        ece, _ = R.ece_binary(mean_p.permute(0, 2, 3, 1).numpy(), target)
        return ent, ece
    ref = []
    for seed in range(8):
        gg = torch.Generator().manual_seed(100 + seed)
        probs = []
        for t in range(T):
            keep = [(torch.rand(2, c, generator=gg) >= cfg.dropout).to(torch.uint8) for _, c in R.dropout_sites(cfg)]
            probs.append(torch.softmax(R.unet_forward(sd, x, cfg, keep), 1))
        ref.append(stats(torch.stack(probs).mean(0)))
    ref = np.array(ref)
    ours = net.forward_samples(x, T, dropout_mode=1, seed=20)
    mine = stats(torch.softmax(ours.permute(0, 1, 4, 2, 3), 2).mean(0).cpu())
    for k in range(2):
        mu, sd_ = ref[:, k].mean(), ref[:, k].std(ddof=1)
        assert abs(mine[k] - mu) <= 4 * sd_ + 2e-3, (k, mine[k], mu, sd_)


def test_brats_t20_injected_masks_match_oracle_at_full_resolution():
    """BASELINE config 3 at its real shape: 4x240x240 slices, T = 20 stochastic samples + the weight-scaling pass, the Philox
    keep masks injected into the oracle's restatement of McPredictStep (rechun/dl/customsteps.py:16-39) — every sample,
    the summary and the argmax are compared, not just properties of the run."""
    cfg, sd, net = _net('brats')
    n, T = 2, 20
    x = torch.randn(n, 4, 240, 240, generator=torch.Generator().manual_seed(240))
    logits = net.forward_samples(x, T + 1, dropout_mode=1, det_first=True, seed=20, slice_index0=5, sample0=0)
    scale = metrics.philox_keep_scale_host(20, cfg.dropout, net.site_channels, 5, n, 0, T)
    keep, off = [], 0
    spans = []
    for c in net.site_channels:
        spans.append((off, c))
        off += c
    for t in range(T):
        keep.append([torch.from_numpy((scale[t, :, o:o + c] > 0).astype(np.float32)) for (o, c) in spans])
    ref = R.predict_mc(sd, x, cfg, T, keep)
    probs = torch.softmax(logits.permute(0, 1, 4, 2, 3), 2).cpu()
    dp = (probs[1:] - ref['multi_probabilities']).abs()
    assert dp.max().item() <= P_MAX and dp.mean().item() <= P_MEAN, (dp.max().item(), dp.mean().item())
    assert (probs[0] - ref['ws_probabilities']).abs().max().item() <= P_MAX
    from rcu_b200 import steps
    out = steps.summarize(steps.LazyMultiProbabilities(logits[1:]), do_mi=True, do_var=True, emit_prediction=True, ws_logits=logits[0])
    summ = R.summarize(ref['multi_probabilities'], do_mi=True, do_var=True)
    assert (out['probabilities'].cpu() - summ['probabilities']).abs().max().item() <= P_MAX
    assert (out['entropy'].cpu() - summ['entropy']).abs().max().item() <= 2e-2          # dH/dp <= ln((1-p)/p): flat near 0.5, steep in the tails
    assert (out['mutual_info'].cpu() - summ['mutual_info']).abs().max().item() <= 1e-2
    assert (out['variance'].cpu() - summ['variance']).abs().max().item() <= 2e-3
    assert (out['ws_probabilities'].cpu() - ref['ws_probabilities']).abs().max().item() <= P_MAX
    pm = summ['probabilities']
    clear = (pm[:, 1] - pm[:, 0]).abs() > 4 * P_MAX                                        # argmax can only flip where the classes are within tolerance
    assert torch.equal(out['prediction'].cpu()[clear], (pm[:, 1] > pm[:, 0]).to(torch.uint8)[clear])


def test_free_running_mc_entropy_distribution_matches_nn_dropout2d():
    """free-running MC with the engine's Philox stream against runs whose masks come from nn.Dropout2d itself
    (torch's feature dropout on an all-ones (N, C, 1, 1) tensor: exactly the noise Dropout2d multiplies with).  The per-voxel
    predictive-entropy distributions are compared with the two-sample Kolmogorov-Smirnov distance; voxels are spatially
    correlated, so the yardstick is not a p-value but the KS distance between independent torch-seeded runs."""
    from scipy.stats import ks_2samp
    cfg, sd, net = _net(dict(in_channels=4, dropout=0.2))
    x = torch.randn(2, 4, 64, 64, generator=torch.Generator().manual_seed(0))
    T, runs = 20, 6
    drop = torch.nn.Dropout2d(cfg.dropout).train()

    def entropy_map(mean_p):
        return R.torch_entropy(mean_p, 1).reshape(-1).numpy()
    torch_runs = []
    for seed in range(runs):
        torch.manual_seed(300 + seed)
        probs = []
        for t in range(T):
            keep = [(drop(torch.ones(2, c, 1, 1)) > 0).reshape(2, c).to(torch.uint8) for _, c in R.dropout_sites(cfg)]
            probs.append(torch.softmax(R.unet_forward(sd, x, cfg, keep), 1))
        torch_runs.append(entropy_map(torch.stack(probs).mean(0)))
    ours = net.forward_samples(x, T, dropout_mode=1, seed=20)
    mine = entropy_map(torch.softmax(ours.permute(0, 1, 4, 2, 3), 2).mean(0).cpu())
    d_tt = np.array([ks_2samp(torch_runs[i], torch_runs[j]).statistic for i in range(runs) for j in range(i + 1, runs)])
    d_ot = np.array([ks_2samp(mine, r).statistic for r in torch_runs])
    assert d_ot.mean() <= d_tt.mean() + 3 * d_tt.std(ddof=1) + 0.01, (d_ot, d_tt)
    assert d_ot.max() <= d_tt.max() + 3 * d_tt.std(ddof=1) + 0.02, (d_ot, d_tt)


def test_residual_net_matches_reference_golden_and_oracle():
    """residual=True nets (ConvResidualBlock, common/model/unet.py:42-60): block output = conv - [dropout] - bn of the second
    unit WITHOUT ReLU plus a 1x1 convolution of the block input.  Against the unmodified reference's outputs
    (tests/golden/residual_golden.npz: deterministic logits, `UNet.features`, one stochastic pass with injected masks), and
    layer by layer against the CUDA-core cross-check path."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'residual_golden.npz'))
    cfg = R.UNetConfig(in_channels=4, residual=True)
    sd = R.randomize_statistics(R.init_state_dict(cfg, 20), 7)
    net = model.B200UNet(sd, in_channels=4, dropout=cfg.dropout, provide_features=True)
    assert net.residual
    x = torch.from_numpy(g['input'])
    _close(net(x.cuda()), torch.from_numpy(g['logits']))
    feat = net.features.cpu()
    ref_feat = torch.from_numpy(g['features'])
    assert (feat - ref_feat).abs().max().item() <= 0.05 * max(1.0, ref_feat.abs().max().item())
    mc = net.forward_samples(x, 1, dropout_mode=1, seed=20, slice_index0=0, sample0=0)[0].permute(0, 3, 1, 2)
    _close(mc, torch.from_numpy(g['mc_logits']))
    # tcgen05 path against the CUDA-core kernels over the same bf16 data
    ck = model.B200UNet(sd, in_channels=4, dropout=cfg.dropout)
    ck.set_conv_impl(1)
    a = net.forward_samples(x, 2, dropout_mode=1, det_first=True, seed=5)
    b = ck.forward_samples(x, 2, dropout_mode=1, det_first=True, seed=5)
    assert (a - b).abs().max().item() <= 0.05
    # a residual=False state dict of the same topology still takes the plain path
    plain = model.B200UNet(R.randomize_statistics(R.init_state_dict(R.UNetConfig(in_channels=4), 20), 7), in_channels=4, dropout=cfg.dropout)
    assert not plain.residual


@pytest.mark.parametrize('impl', [0, 1, 2, 3])
def test_logit_difference_output_is_the_difference_of_the_logit_pair(impl):
    """rcu_unet_outputs.logit_diff: the head stores l0 - l1 instead of (l0, l1) — the same float32 subtraction softmax2 does on
    the pair, so it must equal the difference of the pair output bit for bit, in every convolution implementation (pixel-pair
    kernel, CUDA-core cross-check, per-tap kernel, pixel-row halo kernel) and with MC dropout."""
    cfg, sd, net = _net('brats')
    net.set_conv_impl(impl)
    x = torch.randn(3, 4, 48, 64, generator=torch.Generator().manual_seed(5))
    pair = net.forward_samples(x, 4, dropout_mode=1, det_first=True, seed=9, slice_index0=17)
    diff = net.forward_samples(x, 4, dropout_mode=1, det_first=True, seed=9, slice_index0=17, diff=True)
    assert diff.shape == pair.shape[:-1] and diff.dtype == torch.float32
    assert torch.equal(diff, pair[..., 0] - pair[..., 1])
    with pytest.raises(Exception):   # the two outputs are alternatives
        out = model._lib.RcuUnetOutputs()
        out.logits, out.logit_diff = pair.data_ptr(), diff.data_ptr()
        import ctypes
        model._lib.check(model._lib.lib().rcu_unet_forward_ex(net._handle, model._lib.ptr(x.cuda()), 3, 4, 1, 1, 9, 17, 0, None, ctypes.byref(out),
                                                              model._lib.current_stream()))


@pytest.mark.parametrize('shape,n,t,det_first', [((48, 64), 3, 4, True), ((240, 240), 2, 6, True), ((64, 48), 5, 3, False), ((16, 16), 2, 2, True),
                                                  ((240, 240), 9, 21, True)])
def test_first_layer_dedup_is_bit_identical(shape, n, t, det_first):
    """First-layer dedup (default): the first unit's output is stored once per slice (deterministic / every-channel-kept variant)
    and the pixel-pair kernel patches each (sample, slice)'s dropped channels into its landed tiles.  The patched tiles hold
    exactly the values the per-sample tensor held, so every logit must be bit-identical to the run with the dedup off — MC
    with and without the deterministic first sample, eval mode, several chunks per call, image borders and partial tiles."""
    cfg, sd, on = _net('brats', chunk_images=63)
    _, _, off = _net('brats', chunk_images=63)
    off.set_first_layer_dedup(False)
    x = torch.randn(n, 4, *shape, generator=torch.Generator().manual_seed(n * 31 + t))
    a = on.forward_samples(x, t, dropout_mode=1, det_first=det_first, seed=11, slice_index0=5)
    b = off.forward_samples(x, t, dropout_mode=1, det_first=det_first, seed=11, slice_index0=5)
    assert torch.equal(a, b)
    assert torch.equal(on.forward_samples(x, 1, dropout_mode=0), off.forward_samples(x, 1, dropout_mode=0))
    with pytest.raises(NotImplementedError):   # the per-sample tensor does not exist in this mode
        on.debug_activation(0, (n * 1, shape[0], shape[1], 32))
    # a caller-supplied scale table may hold anything: that mode keeps the per-sample tensor (and still agrees)
    scale = metrics.philox_keep_scale_host(11, cfg.dropout, on.site_channels, 5, n, 0, t - (1 if det_first else 0))
    c = on.forward_samples(x, t, dropout_mode=2, det_first=det_first, scale=scale)
    assert torch.equal(c, a)
