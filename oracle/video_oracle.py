"""TEST INFRASTRUCTURE — CPU restatement of the reference's streaming frame loop, used only by tests/ as the checker
for pgtformer_b200/video.py.  Not a product path.

Restates `/root/reference/inference.py`:
  * `rgbnp2tensor` (:6-10)            rgb24 frames -> float32 [t,3,h,w] via numpy's double division
  * `apply_net_to_frames` (:12-19)    model(window)[0][1] -> clamp -> * 255 -> uint8 (truncation)
  * `process_video_ffmpeg` (:37-76)   the three-frame buffer: first frame duplicated, one window per new frame, last
                                      frame duplicated for the final window
Parity is pinned by construction (the loop is restated line by line, the arithmetic is numpy's own); the model inside
the window is whatever callable the test passes.
"""
import numpy as np


def rgbnp2tensor(rgbnplist):
    """inference.py:6-10 without the .cuda(): float32 [t,3,h,w]."""
    rgbnps = np.array(rgbnplist).copy()
    lqinput = np.array(np.array(rgbnps) / 255.0, np.float32)          # [t,h,w,c]
    return np.ascontiguousarray(lqinput.transpose(0, 3, 1, 2))


def tensor2rgb(restored_middle):
    """inference.py:15-18: restored_middle float32 [3,h,w] -> uint8 [h,w,3]."""
    restored = np.clip(np.asarray(restored_middle, np.float32), 0, 1)
    return np.array(restored.transpose(1, 2, 0) * 255, np.uint8)


def restore_frames(frames, apply_window):
    """inference.py:37-76 over an in-memory / iterable frame sequence.  apply_window(list of 3 frames) -> frame."""
    out = []
    frame_buffer = []
    it = iter(frames)
    first = next(it, None)
    if first is not None:
        frame_buffer.append(first)
        frame_buffer.append(first)                 # pad the previous frame (duplicate the first frame)
    for frame in it:
        frame_buffer.append(frame)
        if len(frame_buffer) == 3:
            out.append(apply_window(list(frame_buffer)))
            frame_buffer.pop(0)
    if len(frame_buffer) == 2:
        frame_buffer.append(frame_buffer[-1])      # pad the last frame
        out.append(apply_window(list(frame_buffer)))
    return out
