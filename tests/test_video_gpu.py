"""Streaming pipeline on the GPU: conversions bit-exact against numpy, frame gather, and the batched / frame-reusing
restore against the reference's one-window-per-call loop (oracle/video_oracle.py) — uint8 output must be identical."""
import numpy as np
import pytest
import torch

from oracle import video_oracle as VO

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def _ops():
    from pgtformer_b200 import ops
    return ops


def test_u8_to_f32_is_numpy_exact():
    o = _ops()
    fr = np.arange(256, dtype=np.uint8).repeat(3).reshape(2, 16, 8, 3)[:, :, ::-1].copy()
    fr[1] = np.random.RandomState(0).randint(0, 256, size=fr[1].shape, dtype=np.uint8)
    out = torch.empty(2, 3, 16, 8, dtype=torch.float32, device=DEV)
    o.u8hwc_to_f32nchw(torch.from_numpy(fr).to(DEV), out)
    assert np.array_equal(out.cpu().numpy(), VO.rgbnp2tensor(list(fr)))


def test_f32_to_u8_is_numpy_exact():
    o = _ops()
    g = torch.Generator().manual_seed(3)
    x = torch.rand(6, 3, 8, 16, generator=g) * 1.4 - 0.2
    x[0, 0, 0, :6] = torch.tensor([0.0, 1.0, 254.999 / 255, 0.5, 1.0 / 255, 0.99999994])
    out = torch.empty(2, 8, 16, 3, dtype=torch.uint8, device=DEV)
    o.f32nchw_to_u8hwc(x.to(DEV), out, first=1, step=3)
    ref = np.stack([VO.tensor2rgb(x[1].numpy()), VO.tensor2rgb(x[4].numpy())])
    assert np.array_equal(out.cpu().numpy(), ref)
    o.f32nchw_to_u8hwc(x.to(DEV), out, first=0, step=1)
    assert np.array_equal(out.cpu().numpy()[0], VO.tensor2rgb(x[0].numpy()))


def test_gather_frames():
    o = _ops()
    x = torch.randn(5, 7, 8, 16, device=DEV).bfloat16()
    idx = torch.tensor([4, 4, 0, 2, 1, 2, 3], dtype=torch.int32, device=DEV)
    out = torch.empty(7, 7, 8, 16, dtype=torch.bfloat16, device=DEV)
    o.gather_frames(x, idx, out)
    assert torch.equal(out, x[idx.long()])


@pytest.fixture(scope='module')
def model(network_g):
    from archs.pgtformer_arch import PGTFormer
    kw = dict(network_g)
    kw.pop('type', None)
    m = PGTFormer(**kw).cuda()
    m.eval()
    return m


def _frames(n, size, seed):
    return np.random.RandomState(seed).randint(0, 256, size=(n, size, size, 3), dtype=np.uint8)


def _reference_loop(model, frames, w=1.0):
    def apply_window(win):                       # apply_net_to_frames, inference.py:12-19, on this repo's model
        x = torch.from_numpy(VO.rgbnp2tensor(win)).cuda()
        with torch.no_grad():
            mid = model(x, w=w, adain=True)[0][1]
        return VO.tensor2rgb(mid.float().cpu().numpy())
    return np.stack(VO.restore_frames(list(frames), apply_window))


@pytest.mark.parametrize('n,batch,reuse', [(1, 4, True), (2, 4, True), (5, 4, True), (11, 4, True), (11, 4, False), (9, 16, True)])
def test_streaming_restore_equals_window_by_window_loop(model, n, batch, reuse):
    """Batching windows and computing per-frame work once per distinct frame must not change a single output byte."""
    from pgtformer_b200.video import VideoRestorer
    frames = _frames(n, 64, 100 + n)
    got = VideoRestorer(model, w=1.0, adain=True, clips_per_batch=batch, reuse_frames=reuse).restore(frames)
    ref = _reference_loop(model, frames)
    assert got.shape == ref.shape and got.dtype == np.uint8
    assert np.array_equal(got, ref), 'max |d| = %d' % np.abs(got.astype(int) - ref.astype(int)).max()


def test_stream_iterator_equals_restore(model):
    from pgtformer_b200.video import VideoRestorer
    frames = _frames(10, 64, 7)
    vr = VideoRestorer(model, clips_per_batch=3)
    got = np.stack(list(vr.stream(iter(frames))))
    assert np.array_equal(got, vr.restore(frames))
    assert list(vr.stream(iter([]))) == []
