"""CUDA-graph capture of fixed-shape launch sequences.

Every kernel of this backend is launched through the C ABI on torch's current stream, with tensor maps passed by value and
no hidden synchronisation, so a forward whose shapes and device buffers are fixed -- e.g. BASELINE configs[1]: given
heat-maps -> un-projection -> V2VNet -> NMS / top-K, ~40 launches of a few microseconds each -- can be captured once and
replayed as ONE graph launch; the host side (argument structs, Python) disappears from the step.  The reference has no
counterpart (it relies on cudnn + eager launches).

What may be inside ``fn``: kernel launches, ``torch.empty`` and device-side tensor ops.  What may not: host<->device
copies of pageable memory, ``.item()`` / ``.cpu()``, anything data-dependent in shape (the person-cube stage of the full
pipeline compacts the valid proposals on the host and therefore stays eager).
"""
from __future__ import annotations

import torch


class GraphedCall:
    """``g = GraphedCall(fn, *example_inputs)``; ``out = g(*inputs)`` copies ``inputs`` into the captured input buffers,
    replays the graph and returns the captured output tensors (valid until the next call; clone to keep).

    ``fn`` takes CUDA tensors (other arguments: close over them) and returns a tensor or a tuple / list of tensors."""

    def __init__(self, fn, *example_inputs, warmup=2):
        if not example_inputs or not all(isinstance(t, torch.Tensor) and t.is_cuda for t in example_inputs):
            raise ValueError("GraphedCall captures functions of CUDA tensors")
        self._inputs = [t.clone() for t in example_inputs]
        self._fn = fn       # the graph reads every tensor fn closes over (camera tables, ...) by ADDRESS: keep them alive
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), torch.no_grad():      # warm-up off the default stream: lazy one-time work (function
            for _ in range(max(int(warmup), 1)):           # attributes, packed weights, caches) must not be captured
                fn(*self._inputs)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph), torch.no_grad():
            out = fn(*self._inputs)
        self._single = isinstance(out, torch.Tensor)
        self._outputs = [out] if self._single else list(out)

    def __call__(self, *inputs):
        if len(inputs) != len(self._inputs):
            raise ValueError("expected %d inputs" % len(self._inputs))
        for dst, src in zip(self._inputs, inputs):
            if dst.shape != src.shape or dst.dtype != src.dtype:
                raise ValueError("GraphedCall inputs must keep the captured shapes and dtypes")
            if dst.data_ptr() != src.data_ptr():
                dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return self._outputs[0] if self._single else tuple(self._outputs)


def graphed_proposal_net(root_net, heatmaps, meta, flip_xcoords=None):
    """``CuboidProposalNet`` (evaluation mode) for a fixed rig and heat-map shape as one CUDA graph:
    ``g(*heatmaps) -> (root_cubes, grid_centers)``.  The packed camera table of ``meta`` is built once, outside."""
    from . import ops
    if root_net.training:
        raise ValueError("graph capture is for the inference path (call .eval())")
    device = heatmaps[0].device
    cams = ops.pack_cameras(meta, root_net.project_layer.img_size, flip_xcoords).to(device)

    def fn(*hms):
        root = root_net.root_volume(list(hms), None, cams=cams)
        return root, root_net.proposal_layer(root, None)

    return GraphedCall(fn, *[h.contiguous() for h in heatmaps])


def graphed_backbone(backbone, images):
    """``PoseResNet`` (evaluation mode) on a fixed image-batch shape as one CUDA graph: ``g(images) -> heat-maps``
    (the zero-copy channel-last view ``backbone(images)`` returns; valid until the next call)."""
    if backbone.training:
        raise ValueError("graph capture is for the inference path (call .eval())")
    return GraphedCall(lambda x: backbone(x), images.float().contiguous())
