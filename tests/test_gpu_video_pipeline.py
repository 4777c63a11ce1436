"""GPU: FrameBatchPipeline (the role of FrameCallbackPool, nunif/utils/video.py:1622-1757): ticket order, ragged last batch,
callbacks that emit fewer / later frames, and the iw3 SBS callback bit-identical to the unpipelined path."""
import pytest
import torch

from nunif_b200 import synth

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _frames(n, h=48, w=80):
    g = torch.Generator().manual_seed(5)
    return [torch.randint(0, 256, (h, w, 3), dtype=torch.uint8, generator=g) for _ in range(n)]


@pytest.mark.parametrize("batch,depth,n", [(4, 3, 23), (1, 2, 7), (8, 3, 8), (3, 4, 40)])
def test_pipeline_returns_every_frame_in_order(batch, depth, n):
    from nunif_b200.nunif.video import FrameBatchPipeline
    frames = _frames(n)

    def cb(x):                                     # flip + darken: order-sensitive, exactly representable
        return torch.flip(x, dims=[3]) * 0.5
    pipe = FrameBatchPipeline(cb, batch, DEV, depth=depth)
    out = []
    for f in frames:
        out += pipe(f)
    out += pipe(None)
    assert len(out) == n and pipe.submitted == pipe.returned
    for f, o in zip(frames, out):
        want = (torch.flip(f.permute(2, 0, 1).float() / 255.0, dims=[2]) * 0.5 * 255.0).round().to(torch.uint8).permute(1, 2, 0)
        assert not o.is_cuda and torch.equal(o, want)


def test_pipeline_with_lookahead_callback():
    """A callback that holds frames back (EMA look-ahead) and emits them later / at flush."""
    from nunif_b200.nunif.video import FrameBatchPipeline
    held = []

    def cb(x):
        held.append(x)
        if len(held) < 3:
            return None
        return held.pop(0)
    frames = _frames(12)
    pipe = FrameBatchPipeline(cb, 2, DEV)
    out = []
    for f in frames:
        out += pipe(f)
    out += pipe.finish()
    assert len(out) == 8                      # 6 batches in, the first 2 held back by the callback
    for f, o in zip(frames, out):
        assert torch.equal(o, f)


def test_pipeline_iw3_sbs_matches_direct_path():
    from nunif_b200.nunif.video import FrameBatchPipeline
    from nunif_b200.iw3 import stereo_sbs, hwc_to_chw_float, chw_float_to_hwc
    frames = [(synth.synth_image(30 + i, 3, 96, 160, smooth=False).permute(1, 2, 0) * 255).round().to(torch.uint8) for i in range(6)]
    depth = synth.synth_depth(3, 6, 36, 60).to(DEV)
    k = [0]

    def cb(x):
        d = depth[k[0]:k[0] + x.shape[0]]
        k[0] += x.shape[0]
        return stereo_sbs(x, d, 2.0, 0.5, method="forward_fill", edge_dilation=[2, 1])
    pipe = FrameBatchPipeline(cb, 2, DEV)
    out = []
    for f in frames:
        out += pipe(f)
    out += pipe(None)
    x = hwc_to_chw_float(torch.stack(frames).to(DEV))
    want = chw_float_to_hwc(stereo_sbs(x, depth, 2.0, 0.5, method="forward_fill", edge_dilation=[2, 1])).cpu()
    assert len(out) == 6
    for o, w in zip(out, want):
        assert torch.equal(o, w)
