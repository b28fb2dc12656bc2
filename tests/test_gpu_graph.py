"""The replayed CUDA graph of the whole training iteration (trainer.GraphedStep: three streams -- compute, weight-gradient side
stream, generator branch stream -- both Adam updates inside) must reproduce the eager iteration: same losses and same parameters
after several iterations from the same start, on the pose workload with the face discriminator (tiny widths), and constructing
the graph must not advance the training state."""
import copy
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'baseline'))
import synth   # noqa: E402

pytestmark = pytest.mark.gpu


def _build(seed):
    import bench
    from fsv import model, trainer
    opt = bench.make_opt('tiny')
    opt.gpu_ids = [0]
    torch.manual_seed(seed)
    step = model.Vid2VidStep(opt)
    for m in [step.netG] + step.d_modules():
        m.train()
    return opt, step, trainer


@pytest.mark.parametrize('dstep_stream', [False, True])
def test_graphed_iteration_equals_eager_iteration(dstep_stream):
    from fsv import trainer as tr
    old = tr.DSTEP_STREAM
    tr.DSTEP_STREAM = dstep_stream
    try:
        opt, step_e, trainer = _build(0)
        _, step_g, _ = _build(1)
        for a, b in zip([step_e.netG] + step_e.d_modules(), [step_g.netG] + step_g.d_modules()):
            b.load_state_dict(copy.deepcopy(a.state_dict()))
        batches = [{k: v.cuda() for k, v in synth.make('pose', 2, 64, 64, seed=s).items()} for s in (1, 2, 3)]
        cur = torch.cuda.Stream()
        cur.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(cur):
            oge, ode = trainer.make_step_optimizers(opt, step_e)
            ogg, odg = trainer.make_step_optimizers(opt, step_g, capturable=True)
            before = {k: v.detach().clone() for k, v in step_g.netG.state_dict().items()}
            graphed = trainer.GraphedStep(step_g, ogg, odg, batches[0])
            for k, v in step_g.netG.state_dict().items():            # construction (3 warm-up iterations + capture) left the state untouched
                assert torch.equal(v, before[k]), k
            for b in batches:
                de, ge, fe, _ = trainer.train_iteration(step_e, oge, ode, b)
                dg, gg, fg, _ = graphed(b)
                torch.cuda.synchronize()
                for n in de:
                    assert abs(float(de[n]) - float(dg[n])) < 2e-3 * max(1.0, abs(float(de[n]))), ('D', n, float(de[n]), float(dg[n]))
                for n in ge:
                    assert abs(float(ge[n]) - float(gg[n])) < 2e-3 * max(1.0, abs(float(ge[n]))), ('G', n, float(ge[n]), float(gg[n]))
                assert float((fe - fg).abs().max()) < 2e-3
        torch.cuda.synchronize()
        # parameters after three updates: Adam's first steps are sign-like (|update| = lr for every element whose gradient is not
        # exactly zero), so compare through the mean absolute difference relative to lr
        lr = opt.lr
        for (n, p), (_, q) in zip(step_e.netG.named_parameters(), step_g.netG.named_parameters()):
            assert float((p - q).abs().mean()) < 0.5 * lr, n
    finally:
        tr.DSTEP_STREAM = old


def test_graphed_generator_equals_eager_eval_forward():
    """fsv.infer.GraphedGenerator (steady-state frame, eval mode, cached hyper-weights, previous-frame branch) vs the eager call."""
    import bench
    from fsv import networks
    from fsv.infer import GraphedGenerator
    opt = bench.make_opt('tiny')
    opt.gpu_ids = [0]
    opt.isTrain = False
    torch.manual_seed(0)
    G = networks.define_G(opt)
    G.init_temporal_network()
    G.cuda().eval()
    b = {k: v.cuda() for k, v in synth.make('pose', 1, 64, 64, seed=4).items()}
    label, lref, iref = b['tgt_label'][:, 0], b['ref_label'], b['ref_image']
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s), torch.no_grad():
        o0 = G(label, lref, iref, [None, None], t=0)
        prev = [label, o0[0].contiguous()]
        e1 = G(label, lref, iref, prev, t=1)
        gg = GraphedGenerator(G, label, lref, iref, prev)
        g1 = gg(label, prev)
        torch.cuda.synchronize()
        assert float((g1[0] - e1[0]).abs().max()) < 1e-5
        prev2 = [label, e1[0].contiguous()]
        e2 = G(label, lref, iref, prev2, t=2)
        g2 = gg(label, prev2)
        torch.cuda.synchronize()
        assert float((g2[0] - e2[0]).abs().max()) < 1e-5
