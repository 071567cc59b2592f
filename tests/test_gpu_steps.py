"""GPU tests of the drop-in BatchSteps: same output keys / shapes / dtypes / devices as the reference steps, values
within the stated bf16 tolerance of the reference golden outputs."""
import pytest
import torch

from rcu_b200 import model, steps
from oracle import restate as R
from helpers import GOLDEN_CONFIGS

pytestmark = pytest.mark.gpu
torch.set_grad_enabled(False)
P_MAX = 1.2e-2


class BatchContext:  # same fields as common/trainloop/context.py:334-342
    def __init__(self, batch, batch_index):
        self.input, self.batch_index, self.output, self.metrics, self.score, self.more = batch, batch_index, {}, {}, None, {}


class Context:
    def __init__(self, net, seed=20):
        self.model, self.device, self._seed = net, torch.device('cuda'), seed

    def get_seed(self):
        return self._seed


def _net(name, **extra):
    cfg = R.UNetConfig(**GOLDEN_CONFIGS[name])
    sd = R.randomize_statistics(R.init_state_dict(cfg, 20), 7)
    return cfg, sd, model.B200UNet(sd, in_channels=cfg.in_channels, dropout=cfg.dropout, dropout_center=cfg.dropout_center, **extra)


@pytest.mark.parametrize('name', sorted(GOLDEN_CONFIGS))
def test_mc_predict_and_summary_steps(golden_unet, name):
    cfg, sd, net = _net(name)
    x = torch.from_numpy(golden_unet[name + '/input'])
    T = golden_unet[name + '/multi_probabilities'].shape[0]
    ctx = Context(net)
    bc = BatchContext({'images': x.double()}, 0)   # any dtype on the host, like the reference
    steps.McPredictStep(T)(bc, None, ctx)
    assert bc.input['images'].dtype == torch.float32 and bc.input['images'].is_cuda
    assert set(bc.output) == {'ws_probabilities', 'multi_probabilities'}
    assert tuple(bc.output['multi_probabilities'].shape) == golden_unet[name + '/multi_probabilities'].shape
    steps.MultiPredictionSummary(do_mi=True, do_var=True)(bc, None, ctx)
    assert set(bc.output) == {'ws_probabilities', 'probabilities', 'entropy', 'mutual_info', 'variance'}
    for key, gold, tol in (('ws_probabilities', 'ws_probabilities', P_MAX), ('probabilities', 'summary_probabilities', P_MAX),
                           ('entropy', 'summary_entropy', 4e-2), ('mutual_info', 'summary_mutual_info', 2e-2),
                           ('variance', 'summary_variance', 5e-3)):
        v, g = bc.output[key], torch.from_numpy(golden_unet[name + '/' + gold])
        assert v.shape == g.shape and v.dtype == torch.float32 and v.is_cuda
        assert (v.cpu() - g).abs().max().item() <= tol, key
    # keep the stacked samples when asked to
    bc = BatchContext({'images': x}, 1)
    mc = steps.McPredictStep(T)
    mc.slices_seen = 0
    mc(bc, None, ctx)
    steps.MultiPredictionSummary(remove_multi_probs=False)(bc, None, ctx)
    multi = bc.output['multi_probabilities']
    assert torch.is_tensor(multi) and (multi.cpu() - torch.from_numpy(golden_unet[name + '/multi_probabilities'])).abs().max().item() <= P_MAX
    assert torch.allclose(multi.mean(0), bc.output['probabilities'], atol=1e-6)


def test_segmentation_predict_step(golden_unet):
    cfg, sd, net = _net('isic')
    x = torch.from_numpy(golden_unet['isic/input'])
    bc = BatchContext({'images': x}, 0)
    steps.SegmentationPredictStep(do_probs=True)(bc, None, Context(net))
    assert set(bc.output) == {'logits', 'probabilities'}
    assert tuple(bc.output['logits'].shape) == (2, 2, 48, 64)
    assert (bc.output['probabilities'].cpu() - torch.from_numpy(golden_unet['isic/probabilities'])).abs().max().item() <= P_MAX
    assert torch.allclose(torch.softmax(bc.output['logits'], 1), bc.output['probabilities'], atol=1e-6)
    # the loop's post-processing (loops.py:214-220): channel_to_end + .cpu().numpy()
    arr = bc.output['probabilities'].permute(0, 2, 3, 1).cpu().numpy()
    assert arr.shape == (2, 48, 64, 2)
    with pytest.raises(ValueError):
        steps.SegmentationPredictStep()(bc, None, object())


def test_ensemble_step_matches_oracle():
    cfg = R.UNetConfig()
    sds = [R.randomize_statistics(R.init_state_dict(cfg, 20 + k), 7 + k) for k in range(3)]   # seeds 20+k like config/train_ensemble
    nets = [model.B200UNet(sd) for sd in sds]
    x = torch.randn(2, 4, 32, 48)
    bc = BatchContext({'images': x}, 0)
    steps.EnsemblePredictionStep(nets[1:])(bc, None, Context(nets[0]))
    assert tuple(bc.output['multi_probabilities'].shape) == (3, 2, 2, 32, 48)
    steps.MultiPredictionSummary()(bc, None, Context(nets[0]))
    ref = R.summarize(R.predict_ensemble(sds, x, cfg)['multi_probabilities'])
    assert (bc.output['probabilities'].cpu() - ref['probabilities']).abs().max().item() <= P_MAX
    assert (bc.output['entropy'].cpu() - ref['entropy']).abs().max().item() <= 4e-2


def test_reference_style_sequential_loop_equals_folded_step():
    """The unmodified McPredictStep loop (T sequential model() calls with the dropout modules flipped) driven
    through B200UNet gives exactly the folded step's samples."""
    cfg, sd, net = _net('brats')
    x = torch.randn(2, 4, 32, 32)
    T = 3
    bc = BatchContext({'images': x}, 0)
    steps.McPredictStep(T)(bc, None, Context(net))
    folded = bc.output['multi_probabilities'].materialize()
    net.reset_stream(seed=20, slice_index=0, sample=0)
    for m in net.modules():
        if isinstance(m, torch.nn.Dropout2d):
            m.train()
    seq = torch.stack([torch.softmax(net(x.cuda()), 1) for _ in range(T)])
    assert torch.allclose(seq, folded, atol=1e-6)
