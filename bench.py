#!/usr/bin/env python
"""Benchmark of the stochastic-inference + uncertainty hot path (BASELINE.json's metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one process per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU arithmetic (oracle port) on the host cores
    python bench.py --config isic_mc | ensemble10 | brats50  # the other BASELINE.json configurations (builder / profile runs)

Default workload `brats_mc` (BASELINE.json configs[2], the one the metric is quoted on).  One step = one synthetic BraTS
subject (155 slices of 4x240x240) through
    McPredictStep(mc=20) [T stochastic forwards + the deterministic weight-scaling forward, folded into the batch]
    -> MultiPredictionSummary [fused softmax / mean / entropy / argmax, the weight-scaling softmax in the same launch]
    -> ECE reliability tables (T2>0-style mask) + Dice/confusion + the 11-threshold U-E sweep [one fused histogram pass].
`value` = MC voxel-samples/s = voxels x T / step time with the subject's images already in HBM (the weight-scaling
forward is executed but not counted).  `e2e` = the same metric through the drop-in steps / hook with
HOST buffers: pinned images are copied in per batch of 32 slices (the reference loader's batch size), the
'probabilities' entry is copied back like loops.py:214-220 does, and the metric tables come back to the host.
Every rank works on its own subject (no data-path collective): "scaling": "weak".

  isic_mc     160 synthetic 3x256x256 images per step, MC T=20, summary, per-image ECE / U-E tables (one launch for all images)
  ensemble10  BASELINE config 4: 10 random-init members over one subject, members sharded over the ranks, the 71 MB
              probability sums reduced and finished by rcu_aggregate_finish_peer (--exchange peer, NVLink peer memory),
              the library's NCCL all-reduce (--exchange nccl) or torch.distributed (--exchange torch).  Strong scaling.
  brats50     BASELINE config 5: 50 synthetic subjects, MC T=20 + the full metric set; the 7750 slices are split evenly over
              the ranks (subjects may span ranks), ONE all-reduce of the per-subject count tables at the end.  Strong scaling;
              a step is the whole 50-subject pass.
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
ENSEMBLE_MEMBERS = 10      # config/test_brats_ensemble.yaml:4,9-18
N_SUBJECTS_50 = 50
ISIC_IMAGES, ISIC_SIZE = 160, 256


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


def make_state_dict(seed=20, in_channels=CHANNELS):
    """Synthetic weights of the baseline net (rcu_b200.synth: seeded torch-default init, non-degenerate BN)."""
    import rcu_b200  # noqa: F401
    from rcu_b200 import synth
    return synth.random_unet_state_dict(in_channels=in_channels, seed=seed)


def centre_head_bias(sd, net_factory, images, weight=None, spread=3.0):
    """Random weights predict one class almost everywhere with p ~ 0.5 (Dice against any label is 0, all voxels share one or
    two reliability bins).  One deterministic forward measures the logit difference l1 - l0 (inside `weight`, e.g. the brain
    mask); the 1x1 head is then rescaled so that its spread is `spread` (logits with sigma ~ 3) and its
    foreground bias shifted so that the median is 0 — about half of those voxels are predicted foreground and the
    probabilities fill every reliability bin.  Pure data synthesis, done once before anything is timed."""
    import torch
    net = net_factory(sd)
    n = min(images.shape[0], 16)
    idx = torch.linspace(0, images.shape[0] - 1, n).long().to(images.device)
    logits = net.forward_samples(images[idx], 1, dropout_mode=0)[0]
    d = (logits[..., 1] - logits[..., 0])
    if weight is not None:
        d = d[weight[idx].bool()]
    d = d.float()
    gain = spread / max(float(d.std().item()), 1e-6)
    shift = float(d.median().item()) * gain
    del net
    return shifted_head(sd, shift, gain), (shift, gain)


def shifted_head(sd, shift, gain=1.0):
    sd = dict(sd)
    sd['conv_cls.1.weight'] = sd['conv_cls.1.weight'] * gain
    b = sd['conv_cls.1.bias'] * gain
    b[1] -= shift
    sd['conv_cls.1.bias'] = b
    return sd


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


# ------------------------------------------------------------------------------------------------ GPU arm: shared plumbing
class _BatchContext:  # same fields as common/trainloop/context.py:334-342
    def __init__(self, batch, batch_index):
        self.input, self.batch_index, self.output, self.metrics, self.score, self.more = batch, batch_index, {}, {}, None, {}


class _Context:
    def __init__(self, net, device, seed=20):
        self.model, self.device, self._seed = net, device, seed

    def get_seed(self):
        return self._seed


class Rig:
    """Process-wide set-up of the GPU arm: device, process group, timing helpers."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.args = torch, dist, args
        torch.set_grad_enabled(False)
        self.rank = int(os.environ.get('RANK', '0'))
        self.world = int(os.environ.get('WORLD_SIZE', '1'))
        self.local_rank = int(os.environ.get('LOCAL_RANK', '0'))
        if not torch.cuda.is_available():
            raise SystemExit('bench.py: no CUDA device (the product path has no CPU fallback; use --impl reference for the CPU arm)')
        torch.cuda.set_device(self.local_rank)
        self.device = torch.device('cuda', self.local_rank)
        if self.world > 1:
            os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
            dist.init_process_group('nccl', device_id=self.device)
        if args.gpus != self.world and self.rank == 0:
            sys.stderr.write('bench.py: --gpus %d but WORLD_SIZE=%d (launch N>1 with torch.distributed.run); using %d\n' % (args.gpus, self.world, self.world))
        self.peaks = measured_peaks()
        self.stream = torch.cuda.current_stream()

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def timed(self, fn, steps):
        """K steps of `fn(i)` bracketed by barrier + synchronize on both sides, CUDA events on the launch stream, max over
        ranks; nvidia-smi clocks are sampled during the region on rank 0.  Returns (ms per step, clocks, last result)."""
        torch = self.torch
        sampler = ClockSampler(self.local_rank)
        if self.rank == 0:
            sampler.start()
            time.sleep(0.3)
        self.barrier()
        w0 = time.time()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(self.stream)
        last = None
        for i in range(steps):
            last = fn(i)
        e1.record(self.stream)
        self.barrier()
        w1 = time.time()
        clocks = sampler.stop(w0, w1) if self.rank == 0 else None
        return self.max_over_ranks(e0.elapsed_time(e1)) / steps, clocks, last

    def finish(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def beta_maps(torch, device, n, seed=20):
    """Metric-kernel test data of SURVEY.md §8d: U-shaped p in [0, 1] (Beta(0.3, 0.3)-like: most mass near 0 and 1, every
    reliability bin and uncertainty interval populated), target ~ Bernoulli(p), 25 % mask, iid per voxel — the worst case
    for data-dependent table reads and counter updates."""
    g = torch.Generator(device=device).manual_seed(seed)
    a = 0.3
    u = torch.rand(n, device=device, generator=g).pow_(1.0 / a)
    v = torch.rand(n, device=device, generator=g).pow_(1.0 / a)
    p = (u / (u + v + 1e-30)).clamp_(0, 1)
    del u, v
    target = (torch.rand(n, device=device, generator=g) < p).to(torch.uint8)
    pred = (p > 0.5).to(torch.uint8)
    mask = (torch.rand(n, device=device, generator=g) < 0.25).to(torch.uint8)
    return p, pred, target, mask


def time_device_call(torch, stream, fn, flush, reps=10):
    """Median device time (ms) of one call with an L2 flush (a >L2 buffer rewritten) before every timed repetition."""
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        fn()
        b.record(stream)
        b.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


# ------------------------------------------------------------------------------------------------ brats_mc / isic_mc
def run_mc_config(args, rig, isic=False):
    torch = rig.torch
    import rcu_b200  # noqa: F401
    from rcu_b200 import model, steps, metrics, tables
    device, stream, rank, world = rig.device, rig.stream, rig.rank, rig.world
    if isic:
        n_items, ch, h, w = ISIC_IMAGES, 3, ISIC_SIZE, ISIC_SIZE
        g = torch.Generator().manual_seed(1000 + rank)
        images_h = torch.rand((n_items, ch, h, w), generator=g)            # RGB / 255 (config/test_isic_baseline.yaml:15-20)
        yy, xx = np.mgrid[0:h, 0:w]
        target_h = np.stack([(((yy - 128 - 3 * (i % 7)) / 60.0) ** 2 + ((xx - 120 + 2 * (i % 5)) / 45.0) ** 2 < 1.0) for i in range(n_items)]).astype(np.uint8)
        mask_h = None
        n_subjects = n_items                                               # ISIC: one image = one subject (bin-eval/eval_uncertainty.py:21,25)
        workload = 'isic_baseline_mc: %d synthetic images 3x%dx%d per step/GPU, T=20 (+1 weight-scaling pass), fused ' \
                   'mean/entropy/argmax, per-image ECE (no mask) + Dice + 11-threshold U-E tables in one launch' % (n_items, h, w)
        flop_per_voxel_sample = 517952
    else:
        n_items, ch, h, w = SLICES, CHANNELS, HEIGHT, WIDTH
        images_h, target_h, mask_h = synth_subject(1000 + rank)
        n_subjects = 1
        workload = WORKLOAD
        flop_per_voxel_sample = FLOP_PER_VOXEL_SAMPLE
    voxels = n_items * h * w

    def factory(sd_):
        return model.B200UNet(sd_, in_channels=ch, dropout=DROPOUT, device=device, seed=20)
    images_d = images_h.to(device)
    mask_d = None if mask_h is None else torch.from_numpy(mask_h).to(device)
    sd, head_shift = centre_head_bias(make_state_dict(in_channels=ch), factory, images_d, mask_d)
    net = factory(sd)
    images_pinned = images_h.pin_memory()
    target_d = torch.from_numpy(target_h).to(device).view(-1)
    mask_flat = None if mask_d is None else mask_d.view(-1)
    break_table = tables.uncertainty_break_table(tables.SWEEP_THRESHOLDS)
    launches = {'n': 0}

    def device_step(step_index, ev=None):
        """Hot path with the inputs resident in HBM.  ev: optional list collecting stage-boundary events."""
        def mark():
            if ev is not None:
                e = torch.cuda.Event(enable_timing=True)
                e.record(stream)
                ev.append(e)
        mark()
        # the head writes logit differences l0 - l1 (all a two-class softmax needs): what McPredictStep asks for as well
        logits = net.forward_samples(images_d, MC_STEPS + 1, dropout_mode=1, det_first=True, slice_index0=step_index * n_items, sample0=0, diff=True)
        launches['n'] += net.last_launch_count()
        mark()
        out = steps.summarize(steps.LazyMultiProbabilities(logits[1:], diff=True), emit_prediction=True, emit_foreground=True, ws_logits=logits[0])
        launches['n'] += 1
        mark()
        res = metrics.eval_fused(out['foreground'], out['prediction'], target_d, mask_flat, 10, tables.SWEEP_THRESHOLDS, n_subjects=n_subjects,
                                 sync=False, break_table=break_table)
        launches['n'] += 1
        mark()
        return out, res

    ctx = _Context(net, device)
    h2d, d2h = {'n': 0}, {'n': 0}
    copy_ev = {'h2d': [], 'd2h': []}
    prob_pinned = torch.empty((n_items, h, w, 2), dtype=torch.float32).pin_memory()
    target_pinned = torch.from_numpy(target_h).pin_memory()
    mask_pinned = None if mask_h is None else torch.from_numpy(mask_h).pin_memory()
    copy_in = torch.cuda.Stream(device)     # host -> device prefetch of the next batch's images
    copy_out = torch.cuda.Stream(device)    # device -> host copy of the previous batch's probabilities

    n_tab = (3 * 11 + 4 * 12 + 1) * n_subjects
    tables_pinned = [torch.empty(n_tab, dtype=torch.int64).pin_memory() for _ in range(2)]
    prob_pinned2 = [prob_pinned, torch.empty_like(prob_pinned).pin_memory()]

    def e2e_run(n_steps, first_index):
        """The call sequence a user of the reference makes (loops.py:204-235) with host buffers on both sides, for `n_steps`
        subjects back to back: images come from (pinned) host memory per 32-item batch, the 'probabilities' entry goes back
        to the host channel-last like loops.py:214-220 does (into a pinned buffer), the foreground / prediction maps stay in
        HBM for the in-memory metric pass, labels and mask come from the host, the metric tables go back.  All of it inside
        the timed region.  The copies run on two side streams and the NEXT batch — of this subject or of the next one, the
        way a pin_memory DataLoader runs ahead — is in flight while the current one is computed; nothing on the host waits
        for the device until the end of the run, where the tables of every step are turned into rows."""
        batches = [(st_, b0) for st_ in range(n_steps) for b0 in range(0, n_items, BATCH)]
        summary = steps.MultiPredictionSummary(emit_prediction=True, emit_foreground=True)

        def fetch(k):
            st_, b0 = batches[k]
            with torch.cuda.stream(copy_in):
                a = torch.cuda.Event(enable_timing=True)
                a.record(copy_in)
                t = images_pinned[b0:b0 + BATCH].to(device, non_blocking=True)
                extra = None
                if b0 + BATCH >= n_items:   # last batch of a subject: its labels / mask travel with it
                    extra = (target_pinned.to(device, non_blocking=True), None if mask_pinned is None else mask_pinned.to(device, non_blocking=True))
                    h2d['n'] += target_pinned.numel() + (0 if mask_pinned is None else mask_pinned.numel())
                ready = torch.cuda.Event(enable_timing=True)
                ready.record(copy_in)
            copy_ev['h2d'].append((a, ready))
            h2d['n'] += t.numel() * 4
            return t, extra, ready
        depth = 3                                   # batches in flight ahead of the compute stream, like a loader's prefetch_factor
        queue = [fetch(k) for k in range(min(depth, len(batches)))]
        mc, fg, pred = None, [], []
        for k, (st_, b0) in enumerate(batches):
            if b0 == 0:
                mc = steps.McPredictStep(MC_STEPS)
                mc.slices_seen = (first_index + st_) * n_items
                fg, pred = [], []
            images_b, extra, ready = queue.pop(0)
            if k + depth < len(batches):
                queue.append(fetch(k + depth))
            stream.wait_event(ready)
            images_b.record_stream(stream)
            bc = _BatchContext({'images': images_b}, b0 // BATCH)
            mc(bc, None, ctx)            # images.float().to(device) inside (customsteps.py:20) finds them resident
            summary(bc, None, ctx)
            n = images_b.shape[0]
            probs = bc.output['probabilities']
            fg.append(bc.output['foreground'])
            pred.append(bc.output['prediction'])
            flat = None
            if extra is not None:        # subject complete: the fused metric pass on the maps in HBM
                target_dev, mask_dev = extra
                target_dev.record_stream(stream)
                if mask_dev is not None:
                    mask_dev.record_stream(stream)
                res = metrics.eval_fused(torch.cat(fg), torch.cat(pred), target_dev.view(-1), None if mask_dev is None else mask_dev.view(-1), 10,
                                         tables.SWEEP_THRESHOLDS, n_subjects=n_subjects, sync=False, break_table=break_table)
                flat = torch.cat([res[0].reshape(-1), res[1].reshape(-1), res[2].view(torch.int64).reshape(-1), res[3].reshape(-1), res[4].reshape(-1)])
            done = torch.cuda.Event()
            done.record(stream)
            with torch.cuda.stream(copy_out):
                copy_out.wait_event(done)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(copy_out)
                prob_pinned2[st_ & 1][b0:b0 + n].copy_(probs.permute(0, 2, 3, 1), non_blocking=True)
                if flat is not None:
                    tables_pinned[st_ & 1].copy_(flat, non_blocking=True)
                    flat.record_stream(copy_out)
                    d2h['n'] += 8 * n_tab
                b.record(copy_out)
            copy_ev['d2h'].append((a, b))
            probs.record_stream(copy_out)
            d2h['n'] += n * h * w * 2 * 4
        copy_out.synchronize()                      # probabilities and tables of the last subject are on the host now
        torch.cuda.current_stream().synchronize()
        t = tables_pinned[(n_steps - 1) & 1].numpy()
        nb1 = 11
        cnt, pos = t[:n_subjects * nb1].reshape(n_subjects, nb1), t[n_subjects * nb1:2 * n_subjects * nb1].reshape(n_subjects, nb1)
        conf_ = t[2 * n_subjects * nb1:3 * n_subjects * nb1].view(np.float64).reshape(n_subjects, nb1)
        ue_ = t[3 * n_subjects * nb1:3 * n_subjects * nb1 + n_subjects * 48].reshape(n_subjects, 4, 12)
        return {'ece': float(np.mean([tables.ece_from_tables(cnt[s_, :10], pos[s_, :10], conf_[s_, :10], n_dim=3) for s_ in range(n_subjects)])),
                'dice': tables.dice_from_counts(ue_[:, 0].sum(), ue_[:, 2].sum(), ue_[:, 3].sum())}

    # ---------------- warm-up (also builds the plan / workspaces)
    warm = max(args.warmup, 3)
    keep = None
    for i in range(warm):
        keep = device_step(i)   # hold the previous step's outputs like the timed loop does: the caching allocator reaches steady state
    torch.cuda.synchronize()

    # ---------------- timed region: K device-resident steps (no per-launch events: those belong to the second pass below)
    launches['n'] = 0
    stage_events = []

    def timed_step(i):
        ev = []
        r = device_step(100 + i, ev)
        stage_events.append(ev)
        return r
    keep = None   # the timed loop holds one previous result like the warm-up did: no new block has to be cudaMalloc'ed inside it
    ms_step, clocks, keep = rig.timed(timed_step, args.steps)
    n_launches = launches['n']
    stages_all = np.array([[ev[j].elapsed_time(ev[j + 1]) for j in range(3)] for ev in stage_events])
    stages = stages_all.mean(0)

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

    # results of the last step, on the host (sanity: the work was really done)
    out, res = keep
    count, positives, conf, ue, invalid = [t.cpu().numpy() for t in res[:5]]
    assert int(ue.sum()) == voxels and int(invalid.sum()) == 0
    assert mask_h is None or int(count[0].sum()) == int(mask_h.sum())
    ece = float(np.mean([tables.ece_from_tables(count[s, :10], positives[s, :10], conf[s, :10], n_dim=3) for s in range(n_subjects)]))
    mean_entropy = float(out['entropy'].mean().item())
    bin_occupancy = [round(float(x), 4) for x in (count[:, :10].sum(0) / max(1, count[:, :10].sum()))]
    pred_pos = float(out['prediction'].float().mean().item())

    # ---------------- e2e through the drop-in steps with host buffers: all K steps, one continuous pipeline
    e2e_run(1, 0)
    torch.cuda.synchronize()
    h2d['n'] = d2h['n'] = 0
    copy_ev['h2d'], copy_ev['d2h'] = [], []
    rig.barrier()
    t0 = time.perf_counter()
    row = e2e_run(args.steps, 200)
    rig.barrier()
    e2e_ms = rig.max_over_ranks((time.perf_counter() - t0) * 1e3) / args.steps
    h2d_busy = sum(a.elapsed_time(b) for a, b in copy_ev['h2d']) / args.steps
    d2h_busy = sum(a.elapsed_time(b) for a, b in copy_ev['d2h']) / args.steps

    # ---------------- ECE-eval ms (second half of BASELINE's metric): full metric set per subject, device resident.
    # 768 MB flush: flushes L2 and keeps the device busy long enough for the host to have the call enqueued when the timed
    # region opens (the number is the device time of the call, not the Python launch latency)
    flush = torch.empty(768 << 20, dtype=torch.uint8, device=device)
    ece_eval_ms = time_device_call(torch, stream, lambda: metrics.eval_fused(out['foreground'], out['prediction'], target_d, mask_flat, 10,
                                                                             tables.SWEEP_THRESHOLDS, n_subjects=n_subjects, sync=False,
                                                                             break_table=break_table), flush)
    bp, bpred, btarget, bmask = beta_maps(torch, device, VOXELS)
    ece_eval_beta_ms = time_device_call(torch, stream, lambda: metrics.eval_fused(bp, bpred, btarget, bmask, 10, tables.SWEEP_THRESHOLDS, sync=False,
                                                                                  break_table=break_table), flush)
    bres = metrics.eval_fused(bp, bpred, btarget, bmask, 10, tables.SWEEP_THRESHOLDS, break_table=break_table)
    beta_occupancy = [round(float(x), 4) for x in (bres[0][0, :10] / max(1, bres[0][0, :10].sum()))]
    del bp, bpred, btarget, bmask
    bp, bpred, btarget, bmask = beta_maps(torch, device, 50 * VOXELS, seed=21)
    ece_eval_beta50_ms = time_device_call(torch, stream, lambda: metrics.eval_fused(bp, bpred, btarget, bmask, 10, tables.SWEEP_THRESHOLDS,
                                                                                    n_subjects=50, sync=False, break_table=break_table), flush, reps=5)
    del bp, bpred, btarget, bmask
    # the run's own maps again (after the synthetic ones: same clocks / power state as those), and with the voxel order shuffled
    ece_eval_again_ms = time_device_call(torch, stream, lambda: metrics.eval_fused(out['foreground'], out['prediction'], target_d, mask_flat, 10,
                                                                                   tables.SWEEP_THRESHOLDS, n_subjects=n_subjects, sync=False,
                                                                                   break_table=break_table), flush)
    ece_eval_shuffled_ms = None
    if n_subjects == 1:
        perm = torch.randperm(out['foreground'].numel(), device=device)
        sp, sd_, st_ = out['foreground'].reshape(-1)[perm].contiguous(), out['prediction'].reshape(-1)[perm].contiguous(), target_d[perm].contiguous()
        sm_ = None if mask_flat is None else mask_flat[perm].contiguous()
        del perm
        ece_eval_shuffled_ms = time_device_call(torch, stream, lambda: metrics.eval_fused(sp, sd_, st_, sm_, 10, tables.SWEEP_THRESHOLDS, sync=False,
                                                                                          break_table=break_table), flush)
        del sp, sd_, st_, sm_
    del flush

    # ---------------- roofline of the dominant kernel family (tcgen05 convolutions)
    ops = net.op_table()
    conv_ms = sum(float(op_ms[i]) for i, o in enumerate(ops) if o['kind'] == 'conv')
    conv_launches = int(sum(int(op_launches[i]) for i, o in enumerate(ops) if o['kind'] == 'conv'))
    images_per_step = n_items * (MC_STEPS + 1)
    conv_flop = 2.0 * sum(o['macs_per_image'] for o in ops if o['kind'] == 'conv') * images_per_step * args.steps
    conv_flop_exec = 2.0 * sum(o['executed_macs_per_image'] for o in ops if o['kind'] == 'conv') * images_per_step * args.steps
    achieved_tflops = conv_flop / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0
    executed_tflops = conv_flop_exec / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0
    per_layer = []
    for i, o in enumerate(ops):
        if op_launches[i] == 0:
            continue
        t = float(op_ms[i]) / args.steps
        row_ = {'op': i, 'kind': o['kind'], 'c_in': o['c_in'], 'c_out': o['c_out'], 'h': o['h'], 'w': o['w'], 'ms_per_step': round(t, 4),
                'launches_per_step': int(op_launches[i]) // args.steps}
        if o['macs_per_image']:
            row_['tflops'] = round(2.0 * o['macs_per_image'] * images_per_step / (t * 1e-3) / 1e12, 1)
            row_['tflops_executed'] = round(2.0 * o['executed_macs_per_image'] * images_per_step / (t * 1e-3) / 1e12, 1)
        per_layer.append(row_)
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, 'profiles', 'conv_traffic.json')
    if os.path.exists(tpath) and not isic:
        with open(tpath) as f:
            tj = json.load(f)
        traffic, traffic_src = tj.get('dram_bytes_per_launch'), tj.get('source')
    peaks = rig.peaks
    roofline = {'bound': 'tensor', 'kernel': 'conv_halo_kernel + conv_wide_kernel (all %d tcgen05 conv launches of a step, aggregated)' % (conv_launches // args.steps),
                'achieved': achieved_tflops, 'peak': peaks['tflops_sustained'], 'unit': 'TFLOP/s',
                'frac': achieved_tflops / peaks['tflops_sustained'],
                'achieved_executed': executed_tflops, 'frac_executed': executed_tflops / peaks['tflops_sustained'],
                'accounting': 'achieved = ALGORITHMIC conv FLOPs (3x3 convs at the output resolution, SURVEY.md §8a) / summed launch time; '
                              'achieved_executed counts the four up-path convs as the 2x2-tap phase convolutions that actually run (2.25x fewer MACs)',
                'traffic': traffic, 'traffic_source': traffic_src,
                'peak_source': '%s bf16_tflops_sustained (kernel timed inside a long step)' % peaks['source'],
                'share_of_step': conv_ms / args.steps / ms_step_op_events,
                # the same FLOPs over the convolutions' share of the TIMED region's step: the per-launch events of the second pass sit
                # between the kernels and cost the programmatic-dependent-launch overlap of their prologues, so `frac` is the lower figure
                'frac_in_timed_region': (conv_flop / args.steps) / (conv_ms / args.steps / ms_step_op_events * ms_step * 1e-3) / 1e12 / peaks['tflops_sustained'],
                'timed_over': 'a second pass of the same %d steps with per-launch CUDA events (%.1f ms/step there)' % (args.steps, ms_step_op_events)}
    agg_alg_bytes = voxels * (8.0 * MC_STEPS + 12)                            # T logit pairs in, mean + entropy out
    agg_all_bytes = voxels * (4.0 * (MC_STEPS + 1) + 8 + 12 + 4 + 1)          # what it moves: logit DIFFERENCES in (4 B per voxel-sample incl. the weight-scaling one), its softmax, mean, entropy, foreground, prediction out
    hist_bytes = VOXELS * 7.0

    def hbm(kernel, nbytes, ms, **extra):
        d = {'kernel': kernel, 'bound': 'hbm', 'achieved': nbytes / (ms * 1e-3) / 1e9, 'peak': peaks['hbm_gbs'], 'unit': 'GB/s',
             'frac': nbytes / (ms * 1e-3) / 1e9 / peaks['hbm_gbs'], 'ms': float(ms)}
        d.update(extra)
        return d
    roofline_hbm = [
        hbm('aggregate (softmax + mean + entropy + argmax + the weight-scaling softmax, ONE launch) on the 8T+12 = 172 B/voxel algorithmic '
            'accounting of SURVEY.md (T logit pairs in, mean + entropy out); the head hands over logit differences, so the launch moves '
            '4(T+1)+25 = 109 B/voxel',
            agg_alg_bytes, stages[1], achieved_all_bytes=agg_all_bytes / (stages[1] * 1e-3) / 1e9,
            frac_all_bytes=agg_all_bytes / (stages[1] * 1e-3) / 1e9 / peaks['hbm_gbs'], bytes_per_voxel_moved=4 * (MC_STEPS + 1) + 25),
        hbm('eval_fused (ECE bins + U-E joint histogram, 7 B/voxel), one subject of Beta(0.3,0.3)-shaped iid p, Bernoulli(p) target, 25% mask',
            hist_bytes, ece_eval_beta_ms, bin_occupancy=beta_occupancy),
        hbm('eval_fused, 50 such subjects per launch', 50 * hist_bytes, ece_eval_beta50_ms, ms_per_subject=ece_eval_beta50_ms / 50),
        hbm('eval_fused on the maps this run produced (%d subject%s per launch)' % (n_subjects, '' if n_subjects == 1 else 's'),
            voxels * (7.0 if mask_h is not None else 6.0), ece_eval_ms, bin_occupancy=bin_occupancy, ms_measured_again_after_the_synthetic_maps=ece_eval_again_ms,
            ms_same_values_voxel_order_shuffled=ece_eval_shuffled_ms)]

    if rank != 0:
        return None

    # ---------------- CPU baseline (rank 0, N=1 only): bounded sample of the same workload on the host cores
    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline and not isic:
        cores = os.cpu_count() or 1
        sd_cpu = {k: (v.cpu() if hasattr(v, 'cpu') else v) for k, v in sd.items()}
        cpu_sample(1, sd_cpu, images_h, target_h, mask_h, cores)
        fwd_s, met_s = cpu_sample(args.cpu_slices, sd_cpu, images_h, target_h, mask_h, cores)
        cpu_baseline = {'value': args.cpu_slices * HEIGHT * WIDTH * MC_STEPS / (fwd_s + met_s), 'unit': UNIT, 'cores': cores, 'kind': 'port',
                        'sample': '%d of 155 slices (T=20 + weight-scaling pass) forward+summary %.2f s on %d torch threads, numpy metric set '
                                  '%.2f s on 1 thread; oracle port of the reference functions' % (args.cpu_slices, fwd_s, cores, met_s),
                        'ece_eval_ms_per_subject_extrapolated': met_s * 1e3 * SLICES / args.cpu_slices}

    value = world * voxels * MC_STEPS / (ms_step * 1e-3)
    line = {
        'metric': METRIC if not isic else METRIC.replace('BraTS', 'ISIC 3x256x256'), 'value': value, 'unit': UNIT, 'n_gpus': world,
        'steps': args.steps, 'warmup': warm,
        'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'bf16',
        'data': 'synthetic',
        'config': {'workload': workload, 'name': 'isic_mc' if isic else 'brats_mc',
                   'l2_policy': 'inputs larger than L2 (%d MB images, %.2f GB logit differences per step, 126 MB L2)' % (images_h.numel() * 4 >> 20, voxels * 4 * 21 / 1e9),
                   'weights': 'random init seed 20, BN statistics randomised (rcu_b200.synth), foreground bias of the 1x1 head centred '
                              'and rescaled (shift %.3f, gain %.3f: logit difference median 0, sigma 3 inside the mask)' % head_shift,
                   'chunk_images': net.chunk_images, 'parallelism': 'subject-sharded x%d, no data-path collective' % world},
        'e2e': {'value': world * voxels * MC_STEPS / (e2e_ms * 1e-3), 'unit': UNIT, 'ms_per_step': e2e_ms, 'steps': args.steps,
                'h2d_bytes_per_step': h2d['n'] // args.steps, 'd2h_bytes_per_step': d2h['n'] // args.steps,
                'copy_stream_busy_ms_per_step': {'h2d': h2d_busy, 'd2h': d2h_busy},
                'exposed_ms_per_step': e2e_ms - ms_step,   # everything the host-buffer route adds to the device-resident step: the first
                # batch's copy-in and the last batch's copy-out (nothing to hide behind), 32-slice batches (partial chunks), host syncs
                'copy_hidden_fraction': max(0.0, min(1.0, 1.0 - (e2e_ms - ms_step) / max(1e-9, h2d_busy + d2h_busy))),
                'api': 'McPredictStep(20)+MultiPredictionSummary per 32-item batch from pinned host images (next batch prefetched, also across '
                       'subjects), probabilities to pinned host memory, metrics.eval_fused on the device maps with labels and mask from the '
                       'host, tables to the host; one host synchronisation at the end of the K steps'},
        'gpu_launches': n_launches,
        'clocks': clocks,
        'roofline': roofline,
        'roofline_hbm': roofline_hbm,
        'cpu_baseline': cpu_baseline,
        'ece_eval_ms': ece_eval_ms,
        'stages_ms': {'unet_forward': float(stages[0]), 'aggregate': float(stages[1]), 'metrics': float(stages[2])},
        'stages_ms_per_step': [[round(float(x), 3) for x in row] for row in stages_all],
        'fraction_of_tensor_roofline_whole_step': value / world * flop_per_voxel_sample / 1e12 / peaks['tflops_sustained'],
        'check': {'ece': float(ece), 'mean_entropy': mean_entropy, 'dice': float(row['dice']), 'prediction_positive_fraction': pred_pos,
                  'bin_occupancy': bin_occupancy},
    }
    if args.layers:
        with open(args.layers, 'w') as f:
            json.dump({'ms_per_step': ms_step, 'layers': per_layer}, f, indent=1)
    return line


# ------------------------------------------------------------------------------------------------ ensemble10 (config 4)
def run_ensemble(args, rig):
    """bin-dl/brats_test_ensemble.py:72-94 with the 10 members sharded over the ranks: every rank runs its members over the
    whole subject, accumulates fp32 probability sums, ONE exchange, then mean / entropy / argmax and the metric tables."""
    torch = rig.torch
    import rcu_b200  # noqa: F401
    from rcu_b200 import model, metrics, tables
    from rcu_b200 import distributed as D
    device, stream, rank, world = rig.device, rig.stream, rig.rank, rig.world
    images_h, target_h, mask_h = synth_subject(1000)      # every rank sees the SAME subject
    images_d = images_h.to(device)
    target_d = torch.from_numpy(target_h).to(device).view(-1)
    mask_d = torch.from_numpy(mask_h).to(device).view(-1)
    lo, hi = D.shard_bounds(ENSEMBLE_MEMBERS, world, rank)
    engines = [model.B200UNet(make_state_dict(seed=20 + k), in_channels=CHANNELS, dropout=DROPOUT, device=device, seed=20 + k, chunk_images=SLICES)
               for k in range(lo, hi)]                    # seeds 20 + k mirror config/train_ensemble/*: seed: 20 + k
    route = args.exchange if world > 1 else 'none'
    exchange = D.PeerExchange(SLICES, HEIGHT, WIDTH) if route == 'peer' else None
    comm = D.Comm() if route == 'nccl' else None
    break_table = tables.uncertainty_break_table(tables.SWEEP_THRESHOLDS)
    launches = {'n': 0}
    ev_log = []

    def step(i, log=False):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        ev[0].record(stream)
        logits = torch.stack([e.forward_samples(images_d, 1, dropout_mode=0)[0] for e in engines]) if engines else None
        launches['n'] += sum(e.last_launch_count() for e in engines)
        target = exchange.sums if exchange is not None else None
        if logits is not None:
            sums = D.aggregate_partial(logits, out=target)
            launches['n'] += 1
        else:
            sums = target.zero_() if target is not None else torch.zeros((SLICES, 2, HEIGHT, WIDTH), dtype=torch.float32, device=device)
        ev[1].record(stream)
        if world == 1:
            out = D.aggregate_finish(sums, ENSEMBLE_MEMBERS, emit_prediction=True, emit_foreground=True)
            launches['n'] += 1
        else:
            out = D._reduce_and_finish(sums, ENSEMBLE_MEMBERS, False, False, True, True, None, comm, exchange)
            launches['n'] += 3 if exchange is not None else 1
        ev[2].record(stream)
        res = metrics.eval_fused(out['foreground'], out['prediction'], target_d, mask_d, 10, tables.SWEEP_THRESHOLDS, sync=False, break_table=break_table)
        launches['n'] += 1
        ev[3].record(stream)
        if log:
            ev_log.append(ev)
        return out, res

    keep = None
    for i in range(max(args.warmup, 3)):
        keep = step(i)
    torch.cuda.synchronize()
    launches['n'] = 0
    ms_step, clocks, keep = rig.timed(lambda i: step(i, True), args.steps)
    n_launches = launches['n']
    st = np.array([[e[j].elapsed_time(e[j + 1]) for j in range(3)] for e in ev_log]).mean(0)
    exch_ms = rig.max_over_ranks(float(st[1]))
    exch_min_ms = -rig.max_over_ranks(-float(st[1]))      # the rank that arrives last waits for nobody: its figure is the exchange proper
    out, res = keep
    count, positives, conf, ue, invalid = [t.cpu().numpy() for t in res[:5]]
    assert int(ue[0].sum()) == VOXELS and int(invalid[0]) == 0
    ece = tables.ece_from_tables(count[0, :10], positives[0, :10], conf[0, :10], n_dim=3)
    sums_bytes = VOXELS * 2 * 4
    line = None
    if rank == 0:
        # bytes that must cross NVLink per rank: the peer route reads (N-1)/N of one rank's sums and writes (N-1)/N of the 17 B/voxel
        # outputs; an all-reduce moves 2 (N-1)/N of the buffer
        link_bytes = 0 if world == 1 else ((world - 1) / world * (sums_bytes + 17 * VOXELS) if route == 'peer' else 2 * (world - 1) / world * sums_bytes)
        value = VOXELS * ENSEMBLE_MEMBERS / (ms_step * 1e-3)
        line = {
            'metric': 'ensemble voxel-members/sec (BraTS 10-member ensemble forward + probability-sum exchange + mean/entropy + ECE/U-E eval)',
            'value': value, 'unit': 'voxel-members/s', 'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3), 'ms_per_step': ms_step,
            'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'bf16', 'data': 'synthetic',
            'config': {'workload': 'brats_ensemble10 (BASELINE config 4): one synthetic subject (155 slices 4x240x240), 10 random-init members '
                                   '(seeds 20..29) sharded over the ranks (%s per rank), fp32 probability sums exchanged once, then '
                                   'mean/entropy/argmax + ECE(mask)/Dice/U-E tables on every rank' % D.shard_sizes(ENSEMBLE_MEMBERS, world),
                       'name': 'ensemble10', 'exchange': route, 'l2_policy': 'inputs larger than L2 (143 MB images, 71 MB sums per member set)'},
            'e2e': None, 'gpu_launches': n_launches, 'clocks': clocks,
            'stages_ms': {'members_forward_and_partial_sums': float(st[0]), 'exchange_and_finish': float(st[1]), 'metrics': float(st[2])},
            'exchange': {'route': route, 'ms_slowest_rank': exch_ms, 'ms_last_arriving_rank': exch_min_ms, 'share_of_step': exch_ms / ms_step,
                         'sums_bytes': sums_bytes, 'nvlink_bytes_per_rank': link_bytes,
                         'nvlink_gbs_per_rank': link_bytes / (exch_min_ms * 1e-3) / 1e9 if world > 1 else None,
                         'of_measured_peer_copy_770_gbs': link_bytes / (exch_min_ms * 1e-3) / 1e9 / 770.0 if world > 1 else None,
                         'note': 'ms_slowest_rank includes the wait for the last rank to finish its members (member imbalance); '
                                 'ms_last_arriving_rank is barrier + exchange + finish with nobody to wait for'},
            'strong_scaling_ceiling': ENSEMBLE_MEMBERS / (world * max(D.shard_sizes(ENSEMBLE_MEMBERS, world))),
            'check': {'ece': float(ece), 'mean_entropy': float(out['entropy'].mean().item()),
                      'prediction_positive_fraction': float(out['prediction'].float().mean().item())},
        }
    if exchange is not None:
        exchange.close()
    if comm is not None:
        comm.close()
    return line


# ------------------------------------------------------------------------------------------------ brats50 (config 5)
def run_brats50(args, rig):
    """bin-eval/eval_uncertainty.py:39-50 over a 50-subject test set produced by MC dropout T=20: the 50 x 155 slices are split
    evenly over the ranks (a subject may span two ranks), every rank runs forward + summary + the fused histogram pass on its
    slices, and ONE grouped all-reduce sums the per-subject count tables / confidence sums (exact for the integers)."""
    torch = rig.torch
    import rcu_b200  # noqa: F401
    from rcu_b200 import model, steps, metrics, tables
    from rcu_b200 import distributed as D
    device, stream, rank, world = rig.device, rig.stream, rig.rank, rig.world
    n_sub = args.subjects
    total_slices = n_sub * SLICES
    lo, hi = D.shard_bounds(total_slices, world, rank)
    segments = []   # (subject, first slice, end slice) on this rank — at most two partial subjects
    s = lo
    while s < hi:
        subj = s // SLICES
        e = min(hi, (subj + 1) * SLICES)
        segments.append((subj, s - subj * SLICES, e - subj * SLICES))
        s = e
    yy, xx = np.mgrid[0:HEIGHT, 0:WIDTH]
    zz = np.arange(SLICES)[:, None, None]
    brain = torch.from_numpy((((yy - 120) / 100.0) ** 2 + ((xx - 120) / 85.0) ** 2 + ((zz - 77) / 75.0) ** 2 < 1.0)).to(device)
    blob = torch.from_numpy(((((yy - 100) / 30.0) ** 2 + ((xx - 140) / 25.0) ** 2 + ((zz - 70) / 20.0) ** 2) < 1.0)).to(device)
    # the rank's slices, resident in HBM before anything is timed (generated on the device: 143 MB per subject)
    data = []
    for subj, a, b in segments:
        g = torch.Generator(device=device).manual_seed(5000 + subj)
        img = torch.randn((SLICES, CHANNELS, HEIGHT, WIDTH), device=device, generator=g)[a:b].clone()
        img *= brain[a:b, None].float()
        data.append((subj, a, b, img, blob[a:b].to(torch.uint8).reshape(-1).contiguous(), brain[a:b].to(torch.uint8).reshape(-1).contiguous()))

    def factory(sd_):
        return model.B200UNet(sd_, in_channels=CHANNELS, dropout=DROPOUT, device=device, seed=20)
    gp = torch.Generator(device=device).manual_seed(1)
    probe = torch.randn((16, CHANNELS, HEIGHT, WIDTH), device=device, generator=gp) * brain[70:86, None].float()
    _, head_shift = centre_head_bias(make_state_dict(), factory, probe, brain[70:86])
    if world > 1:   # identical weights everywhere: the shift / gain measured on rank 0
        t = torch.tensor(list(head_shift), dtype=torch.float64, device=device)
        rig.dist.broadcast(t, src=0)
        head_shift = (float(t[0].item()), float(t[1].item()))
    net = factory(shifted_head(make_state_dict(), *head_shift))
    comm = D.Comm() if world > 1 and args.exchange != 'torch' else None
    break_table = tables.uncertainty_break_table(tables.SWEEP_THRESHOLDS)
    nb1, ncls = 11, len(tables.SWEEP_THRESHOLDS) + 1
    width_i = 2 * nb1 + 4 * ncls + 1
    launches = {'n': 0}
    ev_log = []

    def job(i, subset=None, log=False):
        ints = torch.zeros((n_sub, width_i), dtype=torch.int64, device=device)
        conf = torch.zeros((n_sub, nb1), dtype=torch.float64, device=device)
        e0 = torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for (subj, a, b, img, tgt, msk) in (data if subset is None else data[:subset]):
            logits = net.forward_samples(img, MC_STEPS + 1, dropout_mode=1, det_first=True, slice_index0=subj * SLICES + a, sample0=0, diff=True)
            launches['n'] += net.last_launch_count()
            out = steps.summarize(steps.LazyMultiProbabilities(logits[1:], diff=True), emit_prediction=True, emit_foreground=True, ws_logits=logits[0])
            cnt, pos, cf, ue, inv, _ = metrics.eval_fused(out['foreground'], out['prediction'], tgt, msk, 10, tables.SWEEP_THRESHOLDS, sync=False,
                                                          break_table=break_table)
            launches['n'] += 2
            ints[subj] += torch.cat([cnt.reshape(-1), pos.reshape(-1), ue.reshape(-1), inv.reshape(-1)])
            conf[subj] += cf.reshape(-1)
        e1 = torch.cuda.Event(enable_timing=True)
        e1.record(stream)
        if world > 1:
            if comm is not None:
                comm.allreduce_metric_tables_(ints.view(-1), conf.view(-1))
            else:
                rig.dist.all_reduce(ints)
                rig.dist.all_reduce(conf)
        e2 = torch.cuda.Event(enable_timing=True)
        e2.record(stream)
        if log:
            ev_log.append((e0, e1, e2))
        return ints, conf

    for i in range(max(args.warmup, 1)):
        job(i, subset=min(len(data), 2))        # warm-up passes run the rank's first two segments (same kernels, same shapes)
    torch.cuda.synchronize()
    launches['n'] = 0
    ms_step, clocks, keep = rig.timed(lambda i: job(i, log=True), args.steps)
    n_launches = launches['n']
    compute_ms = float(np.mean([a.elapsed_time(b) for a, b, _ in ev_log]))
    reduce_ms = float(np.mean([b.elapsed_time(c) for _, b, c in ev_log]))
    slowest_compute = rig.max_over_ranks(compute_ms)
    fastest_reduce = -rig.max_over_ranks(-reduce_ms)
    ints, conf = keep
    ints_h, conf_h = ints.cpu().numpy(), conf.cpu().numpy()
    line = None
    if rank == 0:
        cnt, pos, ue = ints_h[:, :nb1], ints_h[:, nb1:2 * nb1], ints_h[:, 2 * nb1:2 * nb1 + 4 * ncls].reshape(n_sub, 4, ncls)
        assert int(ints_h[:, -1].sum()) == 0 and all(int(ue[s_].sum()) == VOXELS for s_ in range(n_sub)), 'count conservation over the all-reduce'
        eces = [tables.ece_from_tables(cnt[s_, :10], pos[s_, :10], conf_h[s_, :10], n_dim=3) for s_ in range(n_sub)]
        dices = [tables.dice_from_counts(ue[s_, 0].sum(), ue[s_, 2].sum(), ue[s_, 3].sum()) for s_ in range(n_sub)]
        value = n_sub * VOXELS * MC_STEPS / (ms_step * 1e-3)
        sizes = D.shard_sizes(total_slices, world)
        line = {
            'metric': METRIC + ' over a %d-subject test set' % n_sub, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
            'warmup': max(args.warmup, 1), 'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
            'dtype': 'bf16', 'data': 'synthetic',
            'config': {'workload': 'brats50 (BASELINE config 5): %d synthetic subjects x 155 slices 4x240x240, MC T=20 (+1), fused summary, ECE(mask)+Dice+'
                                   '11-threshold U-E tables per subject; slices split evenly over the ranks (%s), one grouped all-reduce of the '
                                   'per-subject tables; a step is the whole pass' % (n_sub, sizes),
                       'name': 'brats50', 'exchange': 'none' if world == 1 else ('rcu_allreduce_counts (NCCL)' if comm is not None else 'torch.distributed'),
                       'l2_policy': 'inputs larger than L2', 'weights': 'random init seed 20, head centred and rescaled (shift %.3f, gain %.3f)' % head_shift},
            'e2e': None, 'gpu_launches': n_launches, 'clocks': clocks,
            'stages_ms': {'forward_summary_metrics_rank0': compute_ms, 'forward_summary_metrics_slowest_rank': slowest_compute,
                          'table_allreduce_rank0_incl_wait': reduce_ms, 'table_allreduce_last_arriving_rank': fastest_reduce},
            'exchange': {'bytes': int(ints.numel() * 8 + conf.numel() * 8), 'ms': fastest_reduce, 'share_of_step': fastest_reduce / ms_step},
            'ms_per_subject': ms_step / n_sub,
            'check': {'mean_ece': float(np.mean(eces)), 'mean_dice': float(np.mean(dices)), 'subjects': n_sub},
        }
    if comm is not None:
        comm.close()
    return line


def run_gpu_arm(args):
    rig = Rig(args)
    if args.config == 'brats_mc':
        line = run_mc_config(args, rig, isic=False)
    elif args.config == 'isic_mc':
        line = run_mc_config(args, rig, isic=True)
    elif args.config == 'ensemble10':
        line = run_ensemble(args, rig)
    else:
        line = run_brats50(args, rig)
    if rig.rank == 0 and line is not None:
        print(json.dumps(line), flush=True)
    rig.finish()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--config', default='brats_mc', choices=['brats_mc', 'isic_mc', 'ensemble10', 'brats50'])
    ap.add_argument('--exchange', default='peer', choices=['peer', 'nccl', 'torch'], help='probability-sum exchange of the sharded configs')
    ap.add_argument('--subjects', type=int, default=N_SUBJECTS_50, help='subjects of the brats50 config')
    ap.add_argument('--cpu-slices', type=int, default=2, help='slices per CPU-baseline sample')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--layers', default=None, help='write the per-layer timing table (JSON) here')
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference_arm(args)
    return run_gpu_arm(args)


if __name__ == '__main__':
    sys.exit(main())
