#!/usr/bin/env python
"""Benchmark of the stochastic-inference + uncertainty hot path (BASELINE.json's metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one process per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU arithmetic (oracle port) on the host cores

One step = one synthetic BraTS subject (155 slices of 4x240x240) through
    McPredictStep(mc=20) [T stochastic forwards + the deterministic weight-scaling forward, folded into the batch]
    -> MultiPredictionSummary [fused softmax / mean / entropy / argmax]
    -> ECE reliability tables (T2>0-style mask) + Dice/confusion + the 11-threshold U-E sweep [one fused histogram pass].
`value` = MC voxel-samples/s = voxels x T / step time with the subject's images already in HBM (the weight-scaling
forward is executed but not counted).  `e2e` = the same metric through the drop-in steps / hook with
HOST buffers: pinned images are copied in per batch of 32 slices (the reference loader's batch size), the
'probabilities' entry is copied back like loops.py:214-220 does, and the metric tables come back to the host.
Every rank works on its own subject (no data-path collective): "scaling": "weak".
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SLICES, CHANNELS, HEIGHT, WIDTH = 155, 4, 240, 240
MC_STEPS = 20
BATCH = 32  # config/test_brats_baseline_mc.yaml:11 (reference loader batch size)
VOXELS = SLICES * HEIGHT * WIDTH
FLOP_PER_VOXEL_SAMPLE = 518528  # 14.934 GMAC per 4x240x240 slice-forward
METRIC = 'MC-dropout voxel-samples/sec (BraTS T=20 U-Net forward + mean/entropy + ECE/U-E eval)'
UNIT = 'voxel-samples/s'
WORKLOAD = 'brats_baseline_mc: 1 synthetic subject/step/GPU = 155 slices 4x240x240, T=20 (+1 weight-scaling pass), ' \
           'fused mean/entropy/argmax, ECE(mask)+Dice+11-threshold U-E tables'


def measured_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return {'hbm_gbs': float(d['hbm_gbs']), 'tflops_burst': float(d['bf16_tflops']),
                'tflops_sustained': float(d.get('bf16_tflops_sustained', d['bf16_tflops'])), 'source': 'measured'}
    return {'hbm_gbs': 6650.0, 'tflops_burst': 1590.0, 'tflops_sustained': 1400.0, 'source': 'fallback'}


# ------------------------------------------------------------------------------------------------ synthetic subject
def synth_subject(seed):
    """Images: z-scored noise with a zeroed background outside a brain-like ellipse (so a T2>0 mask exists);
    labels: an off-centre blob.  Deterministic in `seed`."""
    import torch
    g = torch.Generator().manual_seed(seed)
    yy, xx = np.mgrid[0:HEIGHT, 0:WIDTH]
    zz = np.arange(SLICES)[:, None, None]
    r = ((yy - 120) / 100.0) ** 2 + ((xx - 120) / 85.0) ** 2 + ((zz - 77) / 75.0) ** 2
    brain = r < 1.0
    images = torch.randn((SLICES, CHANNELS, HEIGHT, WIDTH), generator=g)
    images *= torch.from_numpy(brain[:, None].astype(np.float32))
    blob = (((yy - 100) / 30.0) ** 2 + ((xx - 140) / 25.0) ** 2 + ((zz - 70) / 20.0) ** 2) < 1.0
    return images, blob.astype(np.uint8), brain.astype(np.uint8)


DROPOUT = 0.05        # config/train_brats_baseline.yaml:6-12


def make_state_dict(seed=20):
    """Synthetic weights of the BraTS baseline net (rcu_b200.synth: seeded torch-default init, non-degenerate BN)."""
    import rcu_b200  # noqa: F401
    from rcu_b200 import synth
    return synth.random_unet_state_dict(in_channels=CHANNELS, seed=seed)


# ------------------------------------------------------------------------------------------------ clocks sampler
class ClockSampler:
    Q = 'clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q, '--format=csv,noheader,nounits',
                                          '-lms', '100'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        for ts, line in self.lines:
            if ts < t0 or ts > t1 + 0.15:
                continue
            f = [x.strip() for x in line.split(',')]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        if not sm:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['no samples inside the timed region'], 'samples': 0}
        return {'sm_mhz': float(np.median(sm)), 'sm_max_mhz': float(max(mx)), 'power_w_max': float(max(pw)), 'reasons': sorted(reasons),
                'samples': len(sm)}


# ------------------------------------------------------------------------------------------------ CPU arm (oracle port)
def cpu_sample(n_slices, sd, images, target, mask, threads):
    """The reference's arithmetic for `n_slices` slices: McPredictStep(20) + MultiPredictionSummary on torch CPU, then
    AddBackgroundProbabilities/ToEntropy + EceBinaryNumpy(mask)+Dice+ConfusionMatrix + 11 x UncertaintyAndCorrectionEvalNumpy
    on numpy — the oracle restatement of those functions (oracle/restate.py).  Returns (seconds forward+summary, seconds metrics)."""
    import torch
    from oracle import restate as R
    torch.set_num_threads(threads)
    x = images[:n_slices]
    cfg = R.UNetConfig(in_channels=CHANNELS, dropout=DROPOUT)
    sites = R.dropout_sites(cfg)
    g = torch.Generator().manual_seed(20)
    keep = [[(torch.rand((n_slices, c), generator=g) >= cfg.dropout).float() for (_, c) in sites] for _ in range(MC_STEPS)]
    t0 = time.perf_counter()
    with torch.no_grad():
        out = R.predict_mc(sd, x, cfg, MC_STEPS, keep)
        summ = R.summarize(out['multi_probabilities'])
    t1 = time.perf_counter()
    prob = summ['probabilities'].permute(0, 2, 3, 1).numpy()
    p_fg = np.ascontiguousarray(prob[..., 1])
    pred = np.argmax(prob, -1).astype(np.uint8)
    tgt, msk = target[:n_slices], mask[:n_slices].astype(bool)
    prob2 = R.add_background_probability(p_fg)
    unc = R.normalized_entropy(prob2)
    R.ece_binary(prob2, tgt, mask=msk)
    R.dice(pred, tgt)
    R.confusion(pred, tgt)
    R.sweep(pred, tgt, unc)
    t2 = time.perf_counter()
    return t1 - t0, t2 - t1


def run_reference_arm(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return 0
    import torch
    torch.set_grad_enabled(False)
    cores = os.cpu_count() or 1
    sd = make_state_dict()
    images, target, mask = synth_subject(1000)
    n_slices = args.cpu_slices
    for _ in range(args.warmup):
        cpu_sample(1, sd, images, target, mask, cores)
    times = []
    for _ in range(args.steps):
        fwd, met = cpu_sample(n_slices, sd, images, target, mask, cores)
        times.append(fwd + met)
    ms = 1e3 * float(np.mean(times))
    value = n_slices * HEIGHT * WIDTH * MC_STEPS / (ms / 1e3)
    sample = '%d of 155 slices per step: torch-CPU McPredictStep(20)+MultiPredictionSummary and numpy ECE/Dice/U-E sweep ' \
             '(oracle port of the reference functions), %d torch threads' % (n_slices, cores)
    line = {'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f32', 'data': 'synthetic', 'config': {'workload': WORKLOAD, 'sample': sample},
            'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': sample},
            'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}, 'gpu_launches': 0}
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------------ GPU arm
class _BatchContext:  # same fields as common/trainloop/context.py:334-342
    def __init__(self, batch, batch_index):
        self.input, self.batch_index, self.output, self.metrics, self.score, self.more = batch, batch_index, {}, {}, None, {}


class _Context:
    def __init__(self, net, device, seed=20):
        self.model, self.device, self._seed = net, device, seed

    def get_seed(self):
        return self._seed


def run_gpu_arm(args):
    import torch
    import torch.distributed as dist
    import rcu_b200  # noqa: F401
    from rcu_b200 import model, steps, metrics, hooks, tables

    torch.set_grad_enabled(False)
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device (the product path has no CPU fallback; use --impl reference for the CPU arm)')
    torch.cuda.set_device(local_rank)
    device = torch.device('cuda', local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=device)
    if args.gpus != world and rank == 0:
        sys.stderr.write('bench.py: --gpus %d but WORLD_SIZE=%d (launch N>1 with torch.distributed.run); using %d\n' % (args.gpus, world, world))

    peaks = measured_peaks()
    sd = make_state_dict()
    net = model.B200UNet(sd, in_channels=CHANNELS, dropout=DROPOUT, device=device, seed=20)
    images_h, target_h, mask_h = synth_subject(1000 + rank)
    images_pinned = images_h.pin_memory()
    images_d = images_h.to(device)
    target_d = torch.from_numpy(target_h).to(device).view(-1)
    mask_d = torch.from_numpy(mask_h).to(device).view(-1)
    break_table = tables.uncertainty_break_table(tables.SWEEP_THRESHOLDS)
    stream = torch.cuda.current_stream()
    launches = {'n': 0}

    def device_step(step_index, ev=None):
        """Hot path with the subject resident in HBM.  ev: optional list collecting stage-boundary events."""
        def mark():
            if ev is not None:
                e = torch.cuda.Event(enable_timing=True)
                e.record(stream)
                ev.append(e)
        mark()
        logits = net.forward_samples(images_d, MC_STEPS + 1, dropout_mode=1, det_first=True, slice_index0=step_index * SLICES, sample0=0)
        launches['n'] += net.last_launch_count()
        mark()
        ws = steps.softmax_planar(logits[0])
        out = steps.summarize(steps.LazyMultiProbabilities(logits[1:]), emit_prediction=True, emit_foreground=True)
        launches['n'] += 2
        mark()
        res = metrics.eval_fused(out['foreground'], out['prediction'], target_d, mask_d, 10, tables.SWEEP_THRESHOLDS, sync=False,
                                 break_table=break_table)
        launches['n'] += 1
        mark()
        return ws, out, res

    hook = hooks.DeviceMetricsHook()
    ctx = _Context(net, device)
    h2d = {'n': 0}
    d2h = {'n': 0}

    prob_pinned = torch.empty((SLICES, HEIGHT, WIDTH, 2), dtype=torch.float32).pin_memory()
    target_pinned = torch.from_numpy(target_h).pin_memory()
    mask_pinned = torch.from_numpy(mask_h).pin_memory()

    copy_in = torch.cuda.Stream(device)     # host -> device prefetch of the next batch's images
    copy_out = torch.cuda.Stream(device)    # device -> host copy of the previous batch's probabilities

    def e2e_step(step_index):
        """The call sequence a user of the reference makes (loops.py:204-235) with host buffers on both sides: images come
        from (pinned) host memory per 32-slice batch, the 'probabilities' entry goes back to the host channel-last like
        loops.py:214-220 does (into a pinned subject buffer), the foreground / prediction maps stay in HBM for the in-memory
        metric hook, labels and mask come from the host, the metric tables go back.  All of it inside the timed region;
        the copies run on two side streams (double-buffered like a pin_memory loader) so that they overlap the forward of
        the neighbouring batch instead of serialising with it."""
        mc = steps.McPredictStep(MC_STEPS)
        mc.slices_seen = step_index * SLICES
        summary = steps.MultiPredictionSummary(emit_prediction=True, emit_foreground=True)
        fg, pred = [], []

        def fetch(b0):
            with torch.cuda.stream(copy_in):
                t = images_pinned[b0:b0 + BATCH].to(device, non_blocking=True)
                ready = torch.cuda.Event()
                ready.record(copy_in)
            h2d['n'] += t.numel() * 4
            return t, ready
        nxt = fetch(0)
        for b0 in range(0, SLICES, BATCH):
            images_b, ready = nxt
            if b0 + BATCH < SLICES:
                nxt = fetch(b0 + BATCH)
            stream.wait_event(ready)
            images_b.record_stream(stream)
            bc = _BatchContext({'images': images_b}, b0 // BATCH)
            mc(bc, None, ctx)            # images.float().to(device) inside (customsteps.py:20) finds them resident
            summary(bc, None, ctx)
            n = images_b.shape[0]
            probs = bc.output['probabilities']
            done = torch.cuda.Event()
            done.record(stream)
            with torch.cuda.stream(copy_out):
                copy_out.wait_event(done)
                prob_pinned[b0:b0 + n].copy_(probs.permute(0, 2, 3, 1), non_blocking=True)
            probs.record_stream(copy_out)
            d2h['n'] += n * HEIGHT * WIDTH * 2 * 4
            fg.append(bc.output['foreground'])
            pred.append(bc.output['prediction'])
        target_dev = target_pinned.to(device, non_blocking=True)
        mask_dev = mask_pinned.to(device, non_blocking=True)
        h2d['n'] += target_pinned.numel() + mask_pinned.numel()
        row = hook.evaluate(step_index, torch.cat(fg), torch.cat(pred), target_dev, mask_dev)   # tables come back to the host inside
        d2h['n'] += 8 * (3 * 11 + 4 * 12 + 1)
        copy_out.synchronize()                      # the subject's probabilities are on the host now
        torch.cuda.current_stream().synchronize()
        return row

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---------------- warm-up (also builds the plan / workspaces)
    keep = None
    for i in range(max(args.warmup, 3)):
        keep = device_step(i)   # hold the previous step's outputs like the timed loop does: the caching allocator reaches steady state
    torch.cuda.synchronize()

    # ---------------- timed region: K device-resident steps (no per-launch events: those belong to the second pass below)
    launches['n'] = 0
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    stage_events = []
    barrier()
    w0 = time.time()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(args.steps):
        ev = []
        keep = device_step(100 + i, ev)
        stage_events.append(ev)
    e1.record(stream)
    barrier()
    w1 = time.time()
    clocks = sampler.stop(w0, w1) if rank == 0 else None
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    ms_step = ms_total / args.steps
    n_launches = launches['n']
    stages = np.array([[ev[j].elapsed_time(ev[j + 1]) for j in range(3)] for ev in stage_events]).mean(0)

    # ---------------- the same K steps again with every kernel launch of the U-Net bracketed by CUDA events on the launch
    # stream (rcu_unet_enable_timing): per-kernel durations for the roofline.  Kept out of the region above because ~1300
    # event records per step widen the gaps between launches.
    net.enable_timing(True)
    net.read_timing()
    torch.cuda.synchronize()
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record(stream)
    for i in range(args.steps):
        keep = device_step(100 + i)
    p1.record(stream)
    torch.cuda.synchronize()
    ms_step_op_events = p0.elapsed_time(p1) / args.steps
    op_ms, op_launches = net.read_timing()
    net.enable_timing(False)
    launches['n'] = n_launches

    # results of the last step, on the host (sanity: the work was really done)
    ws, out, res = keep
    count, positives, conf, ue, invalid = [t.cpu().numpy() for t in res[:5]]
    assert int(ue[0].sum()) == VOXELS and int(invalid[0]) == 0 and int(count[0].sum()) == int(mask_h.sum())
    ece = tables.ece_from_tables(count[0, :10], positives[0, :10], conf[0, :10], n_dim=3)
    mean_entropy = float(out['entropy'].mean().item())

    # ---------------- e2e through the drop-in steps with host buffers
    e2e_step(0)
    torch.cuda.synchronize()
    h2d['n'] = d2h['n'] = 0
    barrier()
    e2e_steps = max(1, min(args.steps, 3))
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        row = e2e_step(200 + i)
    barrier()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3) / e2e_steps

    # ---------------- ECE-eval ms (second half of BASELINE's metric): full metric set per subject, device resident
    for _ in range(3):
        metrics.eval_fused(out['foreground'], out['prediction'], target_d, mask_d, 10, tables.SWEEP_THRESHOLDS, sync=False, break_table=break_table)
    # 768 MB: flushes L2 and keeps the device busy long enough for the host to have the call enqueued when the timed
    # region opens (the number is the device time of the call, not the Python launch latency)
    flush = torch.empty(768 << 20, dtype=torch.uint8, device=device)
    ece_ms = []
    for _ in range(10):
        flush.zero_()  # L2 flush between iterations
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        metrics.eval_fused(out['foreground'], out['prediction'], target_d, mask_d, 10, tables.SWEEP_THRESHOLDS, sync=False, break_table=break_table)
        b.record(stream)
        b.synchronize()
        ece_ms.append(a.elapsed_time(b))
    ece_eval_ms = float(np.median(ece_ms))
    agg_ms = []
    lazy = None
    del flush

    # ---------------- roofline of the dominant kernel family (tcgen05 convolutions)
    ops = net.op_table()
    conv_ms = sum(float(op_ms[i]) for i, o in enumerate(ops) if o['kind'] == 'conv')
    conv_launches = int(sum(int(op_launches[i]) for i, o in enumerate(ops) if o['kind'] == 'conv'))
    images_per_step = SLICES * (MC_STEPS + 1)
    conv_flop = 2.0 * sum(o['macs_per_image'] for o in ops if o['kind'] == 'conv') * images_per_step * args.steps
    achieved_tflops = conv_flop / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0
    per_layer = []
    for i, o in enumerate(ops):
        if op_launches[i] == 0:
            continue
        t = float(op_ms[i]) / args.steps
        row_ = {'op': i, 'kind': o['kind'], 'c_in': o['c_in'], 'c_out': o['c_out'], 'h': o['h'], 'w': o['w'], 'ms_per_step': round(t, 4),
                'launches_per_step': int(op_launches[i]) // args.steps}
        if o['macs_per_image']:
            row_['tflops'] = round(2.0 * o['macs_per_image'] * images_per_step / (t * 1e-3) / 1e12, 1)
        per_layer.append(row_)
    traffic = None
    tpath = os.path.join(ROOT, 'profiles', 'conv_traffic.json')
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get('dram_bytes_per_launch')
    roofline = {'bound': 'tensor', 'kernel': 'conv_halo_kernel + conv_wide_kernel (all %d tcgen05 conv launches of a step, aggregated)' % (conv_launches // args.steps),
                'achieved': achieved_tflops, 'peak': peaks['tflops_sustained'], 'unit': 'TFLOP/s',
                'frac': achieved_tflops / peaks['tflops_sustained'], 'traffic': traffic,
                'peak_source': '%s bf16_tflops_sustained (kernel timed inside a long step)' % peaks['source'],
                'share_of_step': conv_ms / args.steps / ms_step_op_events,
                'timed_over': 'a second pass of the same %d steps with per-launch CUDA events (%.1f ms/step there)' % (args.steps, ms_step_op_events)}
    agg_bytes = VOXELS * (8.0 * (MC_STEPS + 1) + 8 + 12 + 4 + 1 + 4)   # logits in (T+1 samples), ws probs, mean, entropy, prediction, foreground
    hist_bytes = VOXELS * 7.0
    roofline_hbm = [
        {'kernel': 'aggregate (softmax+mean+entropy+argmax, 2 launches)', 'bound': 'hbm', 'achieved': agg_bytes / (stages[1] * 1e-3) / 1e9,
         'peak': peaks['hbm_gbs'], 'unit': 'GB/s', 'frac': agg_bytes / (stages[1] * 1e-3) / 1e9 / peaks['hbm_gbs'], 'ms': float(stages[1])},
        {'kernel': 'eval_fused (ECE bins + U-E joint histogram, 7 B/voxel)', 'bound': 'hbm', 'achieved': hist_bytes / (ece_eval_ms * 1e-3) / 1e9,
         'peak': peaks['hbm_gbs'], 'unit': 'GB/s', 'frac': hist_bytes / (ece_eval_ms * 1e-3) / 1e9 / peaks['hbm_gbs'], 'ms': ece_eval_ms}]

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---------------- CPU baseline (rank 0, N=1 only): bounded sample of the same workload on the host cores
    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        cpu_sample(1, sd, images_h, target_h, mask_h, cores)
        fwd_s, met_s = cpu_sample(args.cpu_slices, sd, images_h, target_h, mask_h, cores)
        cpu_baseline = {'value': args.cpu_slices * HEIGHT * WIDTH * MC_STEPS / (fwd_s + met_s), 'unit': UNIT, 'cores': cores, 'kind': 'port',
                        'sample': '%d of 155 slices (T=20 + weight-scaling pass) forward+summary %.2f s on %d torch threads, numpy metric set '
                                  '%.2f s on 1 thread; oracle port of the reference functions' % (args.cpu_slices, fwd_s, cores, met_s),
                        'ece_eval_ms_per_subject_extrapolated': met_s * 1e3 * SLICES / args.cpu_slices}

    value = world * VOXELS * MC_STEPS / (ms_step * 1e-3)
    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3),
        'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'bf16',
        'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'l2_policy': 'inputs larger than L2 (143 MB images, 1.5 GB logits per step, 126 MB L2)',
                   'weights': 'random init seed 20, BN statistics randomised (rcu_b200.synth)',
                   'chunk_images': net.chunk_images, 'parallelism': 'subject-sharded x%d, no data-path collective' % world},
        'e2e': {'value': world * VOXELS * MC_STEPS / (e2e_ms * 1e-3), 'unit': UNIT, 'ms_per_step': e2e_ms,
                'h2d_bytes_per_step': h2d['n'] // e2e_steps, 'd2h_bytes_per_step': d2h['n'] // e2e_steps,
                'api': 'McPredictStep(20)+MultiPredictionSummary per 32-slice batch from pinned host images, probabilities to host, DeviceMetricsHook.evaluate on host maps'},
        'gpu_launches': n_launches,
        'clocks': clocks,
        'roofline': roofline,
        'roofline_hbm': roofline_hbm,
        'cpu_baseline': cpu_baseline,
        'ece_eval_ms': ece_eval_ms,
        'stages_ms': {'unet_forward': float(stages[0]), 'aggregate': float(stages[1]), 'metrics': float(stages[2])},
        'fraction_of_tensor_roofline_whole_step': value / world * FLOP_PER_VOXEL_SAMPLE / 1e12 / peaks['tflops_sustained'],
        'check': {'ece': float(ece), 'mean_entropy': mean_entropy, 'dice': float(row['dice'])},
    }
    print(json.dumps(line), flush=True)
    if args.layers:
        with open(args.layers, 'w') as f:
            json.dump({'ms_per_step': ms_step, 'layers': per_layer}, f, indent=1)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--cpu-slices', type=int, default=2, help='slices per CPU-baseline sample')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--layers', default=None, help='write the per-layer timing table (JSON) here')
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference_arm(args)
    return run_gpu_arm(args)


if __name__ == '__main__':
    sys.exit(main())
