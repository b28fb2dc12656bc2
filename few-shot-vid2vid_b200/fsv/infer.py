"""Steady-state frame synthesis replayed from a CUDA graph (the inference path of models/vid2vid_model.py:179-205: per frame one
``netG(label, ref_labels, ref_images, prevs, t)`` call in eval mode, with the hyper-network weights cached after frame 0,
generator.py:403-418).  An eager call is ~1500 C-ABI launches issued from Python, i.e. host-bound at small frames; the graph
replays them as one launch.  The reference cannot be captured as is (its ``resample`` builds the sampling grid on the CPU and
copies it to the device on every call, base_network.py:13-37)."""
import torch


class GraphedGenerator:
    """``g = GraphedGenerator(netG, label, ref_labels, ref_images, prev)`` after frame 0 was synthesised eagerly (so that the eval-mode
    weight cache is filled); then ``frame = g(label, prev)`` per frame.  ``prev`` = [prev_labels, prev_images] as netG takes them."""

    def __init__(self, netG, label, label_refs, img_refs, prev, t=1, warmup=2):
        if netG.training:
            raise ValueError('GraphedGenerator replays the eval-mode forward: call netG.eval() first')
        self.netG = netG
        self.static = dict(label=label.clone(), lref=label_refs.clone(), iref=img_refs.clone(),
                           prev=[None if p is None else p.clone() for p in prev])
        st = self.static
        cur = torch.cuda.current_stream()
        if cur == torch.cuda.default_stream():
            raise RuntimeError('GraphedGenerator: run under a non-default stream (torch.cuda.set_stream)')

        def run():
            with torch.no_grad():
                return netG(st['label'], st['lref'], st['iref'], st['prev'], t=t)
        for _ in range(warmup):
            run()
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, stream=cur):
            self.out = run()

    def __call__(self, label, prev):
        self.static['label'].copy_(label, non_blocking=True)
        for s, p in zip(self.static['prev'], prev):
            if s is not None:
                s.copy_(p, non_blocking=True)
        self.graph.replay()
        return self.out
