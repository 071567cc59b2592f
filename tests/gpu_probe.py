"""Developer probe for the GPU box: compact diagnostics of every kernel family against the oracle.

    python tools/gpu_probe.py metrics | aggregate | unet_check | unet_tc | unet_full

Each stage prints PASS/FAIL lines; `unet_tc` compares the tcgen05 path layer by layer against the CUDA-core
cross-check kernels so that a descriptor / layout mistake is localised to the first diverging layer.
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rcu_b200  # noqa: E402
from rcu_b200 import metrics, tables, evaluation, model, steps  # noqa: E402
from oracle import restate as R  # noqa: E402


def report(name, ok, extra=''):
    print('%s %s %s' % ('PASS' if ok else 'FAIL', name, extra), flush=True)
    return ok


def synth_maps(n, seed=20):
    rng = np.random.default_rng(seed)
    p = rng.beta(0.3, 0.3, size=n).astype(np.float32)
    adv = [0.0, 1e-45, 0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9, 1.0, 1 - 2 ** -24]
    adv = np.array(adv, dtype=np.float32)
    adv = np.concatenate([adv, np.nextafter(adv, np.float32(2))[:-2], np.nextafter(adv, np.float32(-1))[1:]])
    br, _, _ = tables.uncertainty_break_table()
    adv = np.concatenate([adv, br, np.nextafter(br, np.float32(0)), np.nextafter(br, np.float32(2))])
    adv = adv[(adv >= 0) & (adv <= 1)]
    adv = adv[:n]
    p[:len(adv)] = adv
    target = (rng.random(n) < p).astype(np.uint8)
    mask = (rng.random(n) < 0.25)
    pred = (p > 0.5).astype(np.uint8)
    flip = rng.random(n) < 0.05
    pred[flip] ^= 1
    return p, target, mask, pred


def stage_metrics():
    ok = True
    for n, s in ((1000003, 1), (155 * 240 * 240, 1), (4 * 65536, 4), (37, 1)):
        p, target, mask, pred = synth_maps(n)
        for use_mask in (False, True):
            m = mask if use_mask else None
            cnt, pos, conf = metrics.calibration_tables(p, target, m, n_subjects=s)
            vps = n // s
            for j in range(s):
                sl = slice(j * vps, (j + 1) * vps)
                c0, p0, s0 = R.calibration_tables(p[sl], target[sl], mask=(mask[sl] if use_mask else None))
                good = (np.array_equal(cnt[j, :10], c0[:10]) and np.array_equal(pos[j, :10], p0[:10].astype(np.int64)) and
                        np.allclose(conf[j, :10], s0[:10], rtol=1e-12, atol=0) and cnt[j, 10] == 0)
                ok &= report('calib n=%d S=%d mask=%s subj=%d' % (n, s, use_mask, j), good,
                             '' if good else '\n got %s\n exp %s\n %s %s' % (cnt[j], c0, conf[j], s0))
        prob2 = R.add_background_probability(p)
        unc = R.normalized_entropy(prob2)
        for kind, vals in (('p', p), ('u64', unc), ('u32', unc.astype(np.float32))):
            tab, inv, order = metrics.ue_tables(vals, pred, target, kind=kind, n_subjects=s)
            vps = n // s
            good = True
            for j in range(s):
                sl = slice(j * vps, (j + 1) * vps)
                u_ref = unc[sl] if kind != 'u32' else unc[sl].astype(np.float32)
                for k, th in enumerate(R.SWEEP_THRESHOLDS):
                    exp = R.uncertainty_counts(pred[sl].astype(bool), target[sl].astype(bool), u_ref > th)
                    got = tables.counts_at_threshold(tab[j], k)
                    if tuple(int(v) for v in exp) != tuple(int(v) for v in got):
                        good = False
                        print('  mismatch kind=%s subj=%d th=%s exp=%s got=%s' % (kind, j, th, exp, got))
            ok &= report('ue kind=%s n=%d S=%d' % (kind, n, s), good and int(inv.sum()) == 0)
        c, po, cf, tab, inv, order = metrics.eval_fused(p, pred, target, mask, n_subjects=s)
        c2, po2, cf2 = metrics.calibration_tables(p, target, mask, n_subjects=s)
        tab2, _, _ = metrics.ue_tables(p, pred, target, kind='p', n_subjects=s)
        good = np.array_equal(c, c2) and np.array_equal(po, po2) and np.array_equal(cf, cf2) and np.array_equal(tab, tab2)
        ok &= report('fused == separate n=%d S=%d' % (n, s), good)
        cm = metrics.confusion_counts(pred, target, n_subjects=s)
        exp = [R.confusion(pred[j * (n // s):(j + 1) * (n // s)], target[j * (n // s):(j + 1) * (n // s)])[:4] for j in range(s)]
        ok &= report('confusion n=%d S=%d' % (n, s), np.array_equal(cm, np.array(exp, dtype=np.int64)))
    # determinism of the float64 sums
    p, target, mask, pred = synth_maps(155 * 240 * 240)
    a = metrics.calibration_tables(p, target, mask)[2]
    b = metrics.calibration_tables(p, target, mask)[2]
    ok &= report('conf_sum deterministic', np.array_equal(a, b))
    # timing (HBM-resident inputs)
    pd, td, md, dd = (torch.from_numpy(x.view(np.uint8) if x.dtype == bool else x).cuda() for x in (p, target, mask, pred))
    bt = tables.uncertainty_break_table()
    for name, fn, nbytes in (('calib', lambda: metrics.calibration_tables(pd, td, md, sync=False), 6),
                             ('ue_p', lambda: metrics.ue_tables(pd, dd, td, kind='p', sync=False, break_table=bt), 6),
                             ('fused', lambda: metrics.eval_fused(pd, dd, td, md, sync=False, break_table=bt), 7)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        print('TIME %s %.4f ms/subject  %.1f GB/s (incl. launch + output alloc)' % (name, ms, nbytes * p.size / ms / 1e6))
    return ok


def stage_aggregate():
    ok = True
    torch.manual_seed(0)
    for (t, n, h, w) in ((5, 3, 48, 64), (20, 2, 240, 240), (1, 1, 16, 16)):
        logits = (torch.randn(t, n, h, w, 2) * 3).cuda()
        lazy = steps.LazyMultiProbabilities(logits)
        out = steps.summarize(lazy, do_mi=t > 1, do_var=t > 1, emit_prediction=True)
        probs = torch.softmax(logits.cpu().permute(0, 1, 4, 2, 3), 2)
        ref = R.summarize(probs, do_mi=t > 1, do_var=t > 1)
        for k in ref:
            d = (out[k].cpu() - ref[k]).abs().max().item()
            ok &= report('aggregate %s T=%d' % (k, t), d < 2e-6, 'maxdiff %.3g' % d)
        pred_ref = (ref['probabilities'][:, 1] > ref['probabilities'][:, 0]).to(torch.uint8)
        agree = (out['prediction'].cpu() == pred_ref).float().mean().item()
        ok &= report('aggregate prediction T=%d' % t, agree > 0.9999, 'agree %.6f' % agree)
        multi = lazy.materialize().cpu()
        d = (multi - probs).abs().max().item()
        ok &= report('materialize T=%d' % t, d < 1e-6, 'maxdiff %.3g' % d)
        out2 = steps.summarize(probs.cuda(), do_mi=t > 1, do_var=t > 1)
        for k in ref:
            d = (out2[k].cpu() - ref[k]).abs().max().item()
            ok &= report('aggregate(from probs) %s T=%d' % (k, t), d < 2e-6, 'maxdiff %.3g' % d)
    host = metrics.philox_keep_scale_host(20, 0.05, [32, 32, 64], 7, 5, 2, 3)
    dev = metrics.philox_keep_scale(20, 0.05, [32, 32, 64], 7, 5, 2, 3).cpu().numpy()
    ok &= report('philox device == host', np.array_equal(host, dev))
    cfg = R.UNetConfig()
    masks = R.philox_keep_masks(cfg, 20, 3, 100, 2)
    sc = metrics.philox_keep_scale_host(20, cfg.dropout, [c for _, c in R.dropout_sites(cfg)], 100, 2, 3, 1)
    keep = np.concatenate([m.numpy() for m in masks], axis=1)
    ok &= report('philox host == oracle', np.array_equal(sc[0] > 0, keep.astype(bool)))
    return ok


def make_net(in_ch=4, seed=20):
    cfg = R.UNetConfig(in_channels=in_ch)
    sd = R.randomize_statistics(R.init_state_dict(cfg, seed), 7)
    return cfg, sd


def stage_unet(impl, n=2, h=48, w=64, T=3):
    ok = True
    cfg, sd = make_net()
    net = model.B200UNet(sd, in_channels=cfg.in_channels, dropout=cfg.dropout, chunk_images=64)
    net.set_conv_impl(impl)
    torch.manual_seed(1)
    x = torch.randn(n, cfg.in_channels, h, w)
    ref = R.unet_forward(sd, x, cfg)
    got = net.forward_samples(x, 1)[0].permute(0, 3, 1, 2).cpu()
    torch.cuda.synchronize()
    d = (got - ref).abs().max().item()
    scale = ref.abs().max().item()
    ok &= report('unet impl=%d deterministic logits' % impl, d < 0.05 * scale, 'maxdiff %.4g (logit scale %.3g, std %.3g)' % (d, scale, ref.std().item()))
    # MC with injected masks == oracle with the same masks
    masks = [R.philox_keep_masks(cfg, 20, t, 0, n) for t in range(T)]
    refs = torch.stack([R.unet_forward(sd, x, cfg, masks[t]) for t in range(T)])
    got = net.forward_samples(x, T + 1, dropout_mode=1, det_first=True, seed=20)
    g = got[1:].permute(0, 1, 4, 2, 3).cpu()
    d = (g - refs).abs().max().item()
    ok &= report('unet impl=%d philox MC logits vs oracle with same masks' % impl, d < 0.05 * scale, 'maxdiff %.4g' % d)
    d0 = (got[0].permute(0, 3, 1, 2).cpu() - ref).abs().max().item()
    ok &= report('unet impl=%d det_first sample' % impl, d0 < 0.05 * scale, 'maxdiff %.4g' % d0)
    sc = metrics.philox_keep_scale_host(20, cfg.dropout, net.site_channels, 0, n, 0, T)
    got2 = net.forward_samples(x, T + 1, dropout_mode=2, det_first=True, scale=sc)
    ok &= report('unet impl=%d injected scale == philox' % impl, torch.equal(got2, got))
    return ok


def stage_unet_tc():
    """Layer-by-layer tcgen05 vs cross-check."""
    ok = True
    cfg, sd = make_net()
    n, h, w = 3, 48, 64
    torch.manual_seed(1)
    x = torch.randn(n, cfg.in_channels, h, w)
    nets = []
    for impl in (1, 0):
        net = model.B200UNet(sd, in_channels=cfg.in_channels, dropout=cfg.dropout, chunk_images=64)
        net.set_conv_impl(impl)
        net.set_first_layer_dedup(False)   # activation 0 per (sample, slice), so that it can be read back and compared
        nets.append(net)
    outs = [net.forward_samples(x, 2, dropout_mode=1, det_first=True, seed=5) for net in nets]
    torch.cuda.synchronize()
    # op list mirrors rcu_unet_plan: first, conv, (pool, conv, conv) x depth, (up, cat, conv) x depth, cls
    dims = [(h, w, 32), (h, w, 32)]
    c = 32
    for l in range(1, 5):
        hh, ww = h >> l, w >> l
        dims += [(hh, ww, c), (hh, ww, 2 * c), (hh, ww, 2 * c)]
        c *= 2
    for l in range(3, -1, -1):
        hh, ww = h >> l, w >> l
        c //= 2
        dims += [(hh, ww, c), (hh, ww, c), (hh, ww, c)]
    n_img = 2 * n
    for i, (hh, ww, cc) in enumerate(dims):
        a = nets[0].debug_activation(i, (n_img, hh, ww, cc)).cpu()
        b = nets[1].debug_activation(i, (n_img, hh, ww, cc)).cpu()
        d = (a - b).abs().max().item()
        s = a.abs().max().item()
        bad = (a - b).abs() > 0.02 * max(s, 1e-3)
        ok &= report('layer %2d %dx%dx%d' % (i, hh, ww, cc), not bad.any().item(),
                     'maxdiff %.4g scale %.3g  nbad %d first_bad %s' % (d, s, int(bad.sum()), tuple(bad.nonzero()[0].tolist()) if bad.any() else ''))
    d = (outs[0] - outs[1]).abs().max().item()
    ok &= report('logits tc vs check', d < 0.05 * outs[0].abs().max().item(), 'maxdiff %.4g' % d)
    return ok


def stage_unet_full():
    cfg, sd = make_net()
    net = model.B200UNet(sd, in_channels=cfg.in_channels, dropout=cfg.dropout)
    x = torch.randn(155, 4, 240, 240).cuda()
    for T in (0, 20):
        for _ in range(2):
            net.forward_samples(x[:32], T + 1, dropout_mode=1 if T else 0, det_first=bool(T))
        torch.cuda.synchronize()
        t0 = time.time()
        lg = net.forward_samples(x, T + 1, dropout_mode=1 if T else 0, det_first=bool(T))
        torch.cuda.synchronize()
        dt = time.time() - t0
        vs = 155 * 57600 * max(T, 1)
        print('TIME unet full subject T=%d: %.1f ms -> %.3g voxel-samples/s (%d launches)' % (T, dt * 1e3, vs / dt, net.last_launch_count()))
        t0 = time.time()
        out = steps.summarize(steps.LazyMultiProbabilities(lg[1:] if T else lg))
        torch.cuda.synchronize()
        print('TIME aggregate T=%d: %.3f ms' % (T, (time.time() - t0) * 1e3))
    return True


def stage_halo_layers():
    """Each halo-eligible conv alone against the per-tap tcgen05 kernel (identical inputs: every other layer is per-tap)."""
    cfg, sd = make_net()
    n, h, w = 3, 48, 64
    x = torch.randn(n, 4, h, w, generator=torch.Generator().manual_seed(3))
    ref = model.B200UNet(sd, in_channels=4, dropout=cfg.dropout, chunk_images=64)
    ref.set_conv_impl(2)
    out_ref = ref.forward_samples(x, 2, dropout_mode=1, det_first=True, seed=5)
    ops = ref.op_table()
    conv_ops = [i for i, o in enumerate(ops) if o['kind'] == 'conv']
    ok = True
    for ci, op_i in enumerate(conv_ops):
        o = ops[op_i]
        net = model.B200UNet(sd, in_channels=4, dropout=cfg.dropout, chunk_images=64)
        net.set_halo_mask(1 << ci)
        out = net.forward_samples(x, 2, dropout_mode=1, det_first=True, seed=5)
        shape = (2 * n, o['h'], o['w'], o['c_out'])
        if ci == len(conv_ops) - 1:
            d = (out - out_ref).abs().max().item()
            s = out_ref.abs().max().item()
            ok &= report('halo conv %2d (head) %s' % (ci, o), d <= 2e-2 * s, 'logits maxdiff %.4g scale %.3g' % (d, s))
            continue
        a = ref.debug_activation(op_i, shape).cpu()
        b = net.debug_activation(op_i, shape).cpu()
        d = (a - b).abs()
        s = a.abs().max().item()
        bad = d > 2 ** -7 * max(s, 1.0)
        ok &= report('halo conv %2d c_in=%d c_out=%d %dx%d' % (ci, o['c_in'], o['c_out'], o['h'], o['w']), not bad.any().item(),
                     'maxdiff %.4g scale %.3g nbad %d first_bad %s' % (d.max().item(), s, int(bad.sum()),
                                                                       tuple(bad.nonzero()[0].tolist()) if bad.any() else ''))
    return ok


if __name__ == '__main__':
    stage = sys.argv[1]
    fn = {'metrics': stage_metrics, 'aggregate': stage_aggregate, 'unet_check': lambda: stage_unet(1),
          'unet_tc_e2e': lambda: stage_unet(0), 'unet_tc': stage_unet_tc, 'unet_full': stage_unet_full, 'halo_layers': stage_halo_layers}[stage]
    good = fn()
    print('STAGE %s %s' % (stage, 'OK' if good else 'FAILED'))
    sys.exit(0 if good else 1)
