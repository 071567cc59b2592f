"""Two-GPU NCCL tests of the sharded paths: MC samples split over ranks with ONE all-reduce of the
probability sums, ensemble members split over ranks, and a subject's metric tables summed across ranks.  Skipped on
boxes with fewer than two GPUs (the single-GPU suites cover the kernels; tests/test_distributed_gloo.py the host logic)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, 'tests'))
    import rcu_b200  # noqa: F401
    from rcu_b200 import distributed as D, metrics, model, steps
    from oracle import restate as R
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    torch.set_grad_enabled(False)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    try:
        cfg = R.UNetConfig(in_channels=4)
        sd = R.randomize_statistics(R.init_state_dict(cfg, 20), 7)
        x = torch.randn(3, 4, 48, 64, generator=torch.Generator().manual_seed(1))
        T = 6
        net = model.B200UNet(sd, in_channels=4, dropout=cfg.dropout, device='cuda:%d' % rank, seed=20)
        # (1) samples sharded over ranks == single-GPU run with the same Philox sample ids
        out = D.mc_predict_sample_sharded(net, x, T, want_mi=True, want_var=True, emit_prediction=True)
        logits = net.forward_samples(x, T + 1, dropout_mode=1, det_first=True)
        ref = steps.summarize(steps.LazyMultiProbabilities(logits[1:]), do_mi=True, do_var=True, emit_prediction=True)
        res = {k: float((out[k].float() - ref[k].float()).abs().max().item()) for k in ('probabilities', 'entropy', 'mutual_info', 'variance')}
        res['pred_mismatch'] = float((out['prediction'] != ref['prediction']).float().mean().item())
        # (1b) the same exchange through the library's own NCCL communicator (rcu_comm_* / rcu_allreduce_probsum) ...
        comm = D.Comm()
        out_c = D.mc_predict_sample_sharded(net, x, T, want_mi=True, want_var=True, emit_prediction=True, comm=comm)
        res['comm_vs_torch'] = max(float((out_c[k].float() - out[k].float()).abs().max().item()) for k in ('probabilities', 'entropy', 'mutual_info', 'variance'))
        # (1c) ... and through the fused peer-memory reduce + finish kernel (rcu_aggregate_finish_peer): twice, the regions are reused
        ex = D.PeerExchange(x.shape[0], x.shape[2], x.shape[3], want_mi=True, want_var=True)
        for rep in range(2):
            out_p = D.mc_predict_sample_sharded(net, x, T, want_mi=True, want_var=True, emit_prediction=True, emit_foreground=True, exchange=ex)
            res['peer_rep%d' % rep] = max(float((out_p[k].float() - ref[k].float()).abs().max().item()) for k in ('probabilities', 'entropy', 'mutual_info', 'variance'))
        res['peer_pred_mismatch'] = float((out_p['prediction'] != ref['prediction']).float().mean().item())
        res['peer_fg'] = float((out_p['foreground'] - out_p['probabilities'][:, 1]).abs().max().item())
        # every rank must hold bit-identical outputs: gather rank 0's probabilities and compare
        probe = out_p['probabilities'].clone()
        dist.broadcast(probe, src=0)
        res['peer_ranks_identical'] = float(torch.equal(probe, out_p['probabilities']))
        # (2) ensemble members sharded over ranks == all members on one GPU
        sds = [R.randomize_statistics(R.init_state_dict(cfg, 20 + k), 7 + k) for k in range(3)]
        lo, hi = D.shard_bounds(3, world, rank)
        local = [model.B200UNet(sds[k], in_channels=4, dropout=cfg.dropout, device='cuda:%d' % rank) for k in range(lo, hi)]
        ens = D.ensemble_member_sharded(local, 3, x)
        ex2 = D.PeerExchange(x.shape[0], x.shape[2], x.shape[3])
        ens_p = D.ensemble_member_sharded(local, 3, x, exchange=ex2)
        all_nets = [model.B200UNet(s, in_channels=4, dropout=cfg.dropout, device='cuda:%d' % rank) for s in sds]
        lg = torch.stack([n.forward_samples(x, 1, dropout_mode=0)[0] for n in all_nets])
        ens_ref = steps.summarize(steps.LazyMultiProbabilities(lg))
        res['ens_prob'] = float((ens['probabilities'] - ens_ref['probabilities']).abs().max().item())
        res['ens_entropy'] = float((ens['entropy'] - ens_ref['entropy']).abs().max().item())
        res['ens_peer_prob'] = float((ens_p['probabilities'] - ens_ref['probabilities']).abs().max().item())
        # (3) a subject whose voxels span ranks: per-rank tables, exact integer all-reduce
        n = 200000
        rng = np.random.default_rng(5)
        p = rng.random(n).astype(np.float32)
        target = (rng.random(n) < p).astype(np.uint8)
        pred = (p > 0.5).astype(np.uint8)
        mask = rng.random(n) < 0.5
        a, b = D.shard_bounds(n, world, rank)
        cnt, pos, conf, ue, inv, order = metrics.eval_fused(p[a:b], pred[a:b], target[a:b], mask[a:b], sync=False)
        # (3b) the library's grouped NCCL call on copies of the same per-rank tables
        ints = torch.cat([cnt.reshape(-1), pos.reshape(-1), ue.reshape(-1)]).contiguous()
        conf_c = conf.clone()
        comm.allreduce_metric_tables_(ints, conf_c)
        D.allreduce_metric_tables_(cnt, pos, conf, ue)
        res['counts_comm_equal'] = float(torch.equal(ints, torch.cat([cnt.reshape(-1), pos.reshape(-1), ue.reshape(-1)])) and
                                         torch.equal(conf_c, conf))
        full = metrics.eval_fused(p, pred, target, mask)
        res['tables_equal'] = float(np.array_equal(cnt.cpu().numpy(), full[0]) and np.array_equal(pos.cpu().numpy(), full[1]) and
                                    np.array_equal(ue.cpu().numpy(), full[3]))
        res['conf_rel'] = float(np.abs(conf.cpu().numpy()[0, :10] - full[2][0, :10]).max() / np.abs(full[2][0, :10]).max())
        np.save(os.path.join(out_dir, 'rank%d.npy' % rank), np.array([res[k] for k in sorted(res)]))
        if rank == 0:
            with open(os.path.join(out_dir, 'keys.txt'), 'w') as f:
                f.write(','.join(sorted(res)))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs')
def test_two_gpu_sharded_paths(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    keys = open(os.path.join(str(tmp_path), 'keys.txt')).read().split(',')
    for rank in range(world):
        r = dict(zip(keys, np.load(os.path.join(str(tmp_path), 'rank%d.npy' % rank))))
        # fp32 sums re-associated by the all-reduce: ulp-level differences only
        # variance: raw second moments (float32 sums, float64 difference) against the single-GPU Welford pass
        assert r['probabilities'] <= 1e-6 and r['entropy'] <= 2e-6 and r['mutual_info'] <= 2e-6 and r['variance'] <= 5e-6, r
        assert r['pred_mismatch'] <= 1e-3, r
        assert r['comm_vs_torch'] <= 5e-6, r
        assert r['peer_rep0'] <= 5e-6 and r['peer_rep1'] <= 5e-6 and r['peer_pred_mismatch'] <= 1e-3 and r['peer_fg'] == 0.0, r
        assert r['peer_ranks_identical'] == 1.0, r
        assert r['ens_prob'] <= 1e-6 and r['ens_entropy'] <= 2e-6 and r['ens_peer_prob'] <= 1e-6, r
        assert r['counts_comm_equal'] == 1.0, r
        assert r['tables_equal'] == 1.0 and r['conf_rel'] <= 1e-12, r
