"""GPU: CUDA-graph capture of the fixed-shape proposal stage (selfpose3d_b200/graphs.py): replays reproduce the eager
launches bit for bit, for new heat-map contents in the captured buffers."""
import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]

from selfpose3d_b200 import graphs, ops, synthetic  # noqa: E402
from selfpose3d_b200.config import default_config  # noqa: E402
from selfpose3d_b200.models import cuboid_proposal_net  # noqa: E402

DEV = "cuda:0"


@pytest.mark.parametrize("mode", ["bf16x3", "bf16"])
def test_graphed_proposal_net_equals_eager(mode):
    prev_dtype, prev_conv = ops.volume_dtype(), ops.float32_conv()
    try:
        ops.set_volume_dtype(torch.bfloat16 if mode == "bf16" else torch.float32)
        if mode != "bf16":
            ops.set_float32_conv(mode)
        cfg = default_config()
        cfg.NETWORK.NUM_JOINTS = 4
        cfg.NETWORK.IMAGE_SIZE, cfg.NETWORK.HEATMAP_SIZE = [96, 128], [24, 32]
        cfg.MULTI_PERSON.INITIAL_CUBE_SIZE = [24, 24, 8]
        cfg.MULTI_PERSON.MAX_PEOPLE_NUM = 3
        cfg.MULTI_PERSON.THRESHOLD = -1.0
        cfg.NETWORK.ROOTNET_ROOTHM = False
        net = cuboid_proposal_net.CuboidProposalNet(cfg)
        net.load_state_dict(synthetic.trained_like_state_dict(net, seed=2), strict=True)
        net = net.to(DEV).eval()
        B, V = 2, 5
        meta = synthetic.make_meta(synthetic.ring_cameras(V, seed=0), B, (96, 128))
        g = torch.Generator().manual_seed(3)
        sets = [[torch.rand(B, 4, 32, 24, generator=g).to(DEV) for _ in range(V)] for _ in range(3)]
        graphed = graphs.graphed_proposal_net(net, sets[0], meta)
        # unrelated allocations after the capture must not disturb the graph (it keeps everything it reads alive: the
        # camera table its function closes over was once freed with the closure and overwritten)
        junk = [torch.randn(1 << 18, device=DEV) for _ in range(32)]
        for hms in sets:
            before = ops._lib.launch_count
            root_g, gc_g = [t.clone() for t in graphed(*hms)]
            assert ops._lib.launch_count == before          # a replay launches nothing through the C ABI from the host
            with torch.no_grad():
                root_e, gc_e = net(hms, meta)
            assert ops._lib.launch_count > before
            assert torch.equal(root_g, root_e) and torch.equal(gc_g, gc_e)
            root_g2 = graphed(*hms)[0]                      # and again after the eager forward's allocations
            assert torch.equal(root_g2, root_e)
        del junk
    finally:
        ops.set_volume_dtype(prev_dtype)
        ops.set_float32_conv(prev_conv)


def test_graphed_backbone_equals_eager():
    from selfpose3d_b200.models import pose_resnet
    cfg = default_config()
    cfg.NETWORK.NUM_JOINTS = 15
    cfg.NETWORK.IMAGE_SIZE, cfg.NETWORK.HEATMAP_SIZE = [96, 128], [24, 32]
    net = pose_resnet.get_pose_net(cfg, is_train=False)
    net.load_state_dict(synthetic.trained_like_state_dict(net, seed=4), strict=True)
    net = net.to(DEV).eval()
    g = torch.Generator().manual_seed(5)
    batches = [torch.rand(3, 3, 128, 96, generator=g).to(DEV) for _ in range(3)]
    graphed = graphs.graphed_backbone(net, batches[0])
    junk = [torch.randn(1 << 18, device=DEV) for _ in range(16)]
    for x in batches:
        before = ops._lib.launch_count
        out = graphed(x)
        assert ops._lib.launch_count == before
        got = out.contiguous().clone()
        with torch.no_grad():
            want = net(x)
        assert out.shape == want.shape and out.stride() == want.stride()      # the zero-copy channel-last view
        assert torch.equal(got, want.contiguous()), float((got - want).abs().max())
    del junk
