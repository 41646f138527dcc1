"""Host logic of the streaming pipeline (pgtformer_b200/video.py) against the restated reference loop."""
import numpy as np
import pytest

from oracle import video_oracle as VO
from pgtformer_b200.video import plan_batches, window_indices


@pytest.mark.parametrize('n', [0, 1, 2, 3, 4, 7, 33])
def test_window_indices_match_reference_buffer_loop(n):
    seen = []
    VO.restore_frames(list(range(n)), lambda win: seen.append(tuple(win)) or 0)
    assert window_indices(n) == seen


@pytest.mark.parametrize('n,b', [(1, 16), (2, 16), (5, 2), (16, 16), (17, 16), (40, 7)])
def test_plan_batches_covers_every_window_once(n, b):
    wins = window_indices(n)
    covered = []
    for first, cnt, lo, hi in plan_batches(n, b):
        assert 1 <= cnt <= b
        need = {j for i in range(first, first + cnt) for j in wins[i]}
        assert min(need) == lo and max(need) == hi, 'distinct-frame range of the batch'
        covered += list(range(first, first + cnt))
    assert covered == list(range(n))


def test_conversion_oracles_follow_numpy_semantics():
    fr = (np.arange(2 * 4 * 8 * 3, dtype=np.int64) * 7 % 256).astype(np.uint8).reshape(2, 4, 8, 3)
    t = VO.rgbnp2tensor(list(fr))
    assert t.dtype == np.float32 and t.shape == (2, 3, 4, 8)
    assert np.array_equal(t[1, 2], (fr[1, :, :, 2] / 255.0).astype(np.float32))
    x = np.array([[-0.2, 0.0, 0.5, 0.9999, 1.0, 1.7]], np.float32).repeat(3, 0).reshape(3, 1, 6)
    assert VO.tensor2rgb(x)[0, :, 0].tolist() == [0, 0, 127, 254, 255, 255]


class _FakeRestorer:
    """VideoRestorer.stream() with the device work replaced by a window function that identifies its three frames."""

    def __new__(cls, batch):
        from pgtformer_b200.video import VideoRestorer
        vr = VideoRestorer(model=None, clips_per_batch=batch)
        vr.batches = []

        def run_batch(frames_u8, local_windows):
            vr.batches.append((frames_u8.shape[0], len(local_windows)))
            return np.stack([_mix([frames_u8[a], frames_u8[b], frames_u8[c]]) for a, b, c in local_windows])
        vr._run_batch = run_batch
        return vr


def _mix(win):
    return (win[0].astype(np.int64) * 5 + win[1].astype(np.int64) * 11 + win[2].astype(np.int64) * 17).astype(np.int64)


@pytest.mark.parametrize('n', [0, 1, 2, 3, 5, 16, 17, 18, 41])
@pytest.mark.parametrize('batch', [1, 2, 16])
def test_stream_bookkeeping_matches_reference_loop(n, batch):
    """Iterator in, iterator out: every window is emitted once, in order, from exactly the frames the reference's
    three-slot buffer would hold — across batch boundaries, buffer trimming and the duplicated end frames."""
    frames = [np.full((2, 2, 3), (7 * i + 1) % 251, np.uint8) for i in range(n)]
    vr = _FakeRestorer(batch)
    got = list(vr.stream(iter(frames)))
    want = VO.restore_frames(frames, _mix)
    assert len(got) == len(want) == n
    for g, w in zip(got, want):
        assert np.array_equal(g, w)
    # bounded memory: a batch never ships more than its windows + the two neighbours
    assert all(fd <= cnt + 2 and cnt <= batch for fd, cnt in vr.batches)
