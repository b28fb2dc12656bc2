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
    """Three identical copies of the tiny pose model: two run the eager iteration (their mutual deviation is the noise floor of the
    path: the split-K weight gradients add with fp32 atomics, and Adam's first steps turn a rounding-level sign change of a tiny
    gradient into a +-lr step), the third replays the captured graph.  Iteration 1 starts from bit-identical states, so everything
    that does not pass through an optimizer step (the D losses) must agree tightly; for the rest the graph may deviate from eager
    copy A by at most a small multiple of what eager copy B does."""
    from fsv import trainer as tr
    old = tr.DSTEP_STREAM
    tr.DSTEP_STREAM = dstep_stream
    try:
        opt, step_a, trainer = _build(0)
        _, step_b, _ = _build(1)
        _, step_g, _ = _build(2)
        for other in (step_b, step_g):
            for a, b in zip([step_a.netG] + step_a.d_modules(), [other.netG] + other.d_modules()):
                b.load_state_dict(copy.deepcopy(a.state_dict()))
        batches = [{k: v.cuda() for k, v in synth.make('pose', 2, 64, 64, seed=s).items()} for s in (1, 2, 3)]
        cur = torch.cuda.Stream()
        cur.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(cur):
            oga, oda = trainer.make_step_optimizers(opt, step_a)
            ogb, odb = trainer.make_step_optimizers(opt, step_b)
            ogg, odg = trainer.make_step_optimizers(opt, step_g, capturable=True)
            before = {k: v.detach().clone() for k, v in step_g.netG.state_dict().items()}
            graphed = trainer.GraphedStep(step_g, ogg, odg, batches[0])
            for k, v in step_g.netG.state_dict().items():            # construction (3 warm-up iterations + capture) left the state untouched
                assert torch.equal(v, before[k]), k
            report = []
            for it, b in enumerate(batches):
                da, ga, fa, _ = trainer.train_iteration(step_a, oga, oda, b)
                db, gb, fb, _ = trainer.train_iteration(step_b, ogb, odb, b)
                dg, gg, fg, _ = graphed(b)
                torch.cuda.synchronize()
                va = {n: float(v) for n, v in list(da.items()) + list(ga.items())}
                vb = {n: float(v) for n, v in list(db.items()) + list(gb.items())}
                vg = {n: float(v) for n, v in list(dg.items()) + list(gg.items())}
                assert all(v == v and abs(v) < 1e4 for v in vg.values()), vg
                for n in va:
                    twin = abs(va[n] - vb[n]) / max(1.0, abs(va[n]))
                    dev = abs(va[n] - vg[n]) / max(1.0, abs(va[n]))
                    report.append((it, n, dev, twin))
                    if it == 0 and n in da:
                        assert dev < 1e-4, ('iteration 1 D-step losses (no optimizer step upstream)', n, va[n], vg[n])
                    # iteration 1: one (sign-like) Adam step of D lies between identical states and the G-step losses; later iterations
                    # compound the amplification (measured on the B200: eager twins drift apart by up to 13 % by iteration 3)
                    bound = max(1e-2, 10.0 * twin) if it == 0 else max(0.3, 10.0 * twin)
                    assert dev < bound, (it, n, va[n], vb[n], vg[n], report)
                ftwin, fdev = float((fa - fb).abs().max()), float((fa - fg).abs().max())
                assert fdev < (max(1e-2, 10.0 * ftwin) if it == 0 else max(0.3, 10.0 * ftwin)), (it, 'frame', fdev, ftwin)
        torch.cuda.synchronize()
        print('graph vs eager (dstep_stream=%s): max deviation %.2e, eager twin %.2e' %
              (dstep_stream, max(r[2] for r in report), max(r[3] for r in report)))
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
    G.cuda()
    b = {k: v.cuda() for k, v in synth.make('pose', 1, 64, 64, seed=4).items()}
    label, lref, iref = b['tgt_label'][:, 0], b['ref_label'], b['ref_image']
    G.train()                       # a freshly initialised network has BatchNorm running statistics (0, 1): in eval mode its activations overflow.
    with torch.no_grad():           # A dozen training-mode forwards give it the statistics a checkpoint would carry.
        for _ in range(12):
            G(label, lref, iref, [label, b['tgt_image'][:, 0]])
    G.eval()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s), torch.no_grad():
        o0 = G(label, lref, iref, [None, None], t=0)
        prev = [label, o0[0].contiguous()]
        e1 = G(label, lref, iref, prev, t=1)
        gg = GraphedGenerator(G, label, lref, iref, prev)
        g1 = gg(label, prev)
        torch.cuda.synchronize()
        assert torch.isfinite(e1[0]).all() and float(e1[0].abs().max()) > 0
        assert float((g1[0] - e1[0]).abs().max()) < 1e-5
        prev2 = [label, e1[0].contiguous()]
        e2 = G(label, lref, iref, prev2, t=2)
        g2 = gg(label, prev2)
        torch.cuda.synchronize()
        assert float((g2[0] - e2[0]).abs().max()) < 1e-5
