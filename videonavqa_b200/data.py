"""Data path, first slice: from decoded frames to the video tensor the drivers consume.

The reference's ``VNQADataset.__getitem__`` (eval/dataset.py:57-93) decodes a clip with OpenCV into a
``np.empty((3, 160, 208, 400))`` float64 array, keeps one random frame out of every four (``DROP_EVERY_N_FRAMES``), and
returns ``(3, 160, 208, 35)`` float64 ``/ 255`` per clip; the driver casts the collated batch to fp32 and moves it to
the GPU (q_and_v_eval.py:93, 98).  Video decoding stays on the CPU here too (OpenCV; no decoder bindings in this
image), but everything after it changes: frames stay uint8, the selected ones are gathered into ONE pinned staging
buffer per batch, a single byte per sample crosses PCIe, and ``vnqa_frames_u8_to_video`` (csrc/data.cu) produces the
``(B, 3, H, W, T)`` fp32 tensor on the device, bit-identical to the reference's arithmetic.
"""
from __future__ import annotations

import ctypes
import random as _random

import numpy as np
import torch

from . import _lib

DROP_EVERY_N_FRAMES = 4                      # eval/utils.py:19
MAX_FRAMES = 35                              # eval/utils.py:20 MAX_ALLOWED_NUM_FRAMES_DROPPING
VID_HEIGHT, VID_WIDTH = 160, 208             # eval/utils.py:24-25


def subsample_indices(vid_len, rng=_random, every=DROP_EVERY_N_FRAMES):
    """The frame subsampling of eval/dataset.py:81-86: one uniformly chosen frame out of every group of ``every``
    consecutive frames (``random.randint`` on the module-level generator, inclusive bounds)."""
    return [rng.randint(i, min(i + every, vid_len) - 1) for i in range(0, vid_len, every)]


def decode_clip_uint8(path):
    """All frames of a clip as (n, H, W, 3) uint8 (BGR, as ``cv2.VideoCapture.read`` returns them, dataset.py:66-73).
    The reference writes each frame into a float64 array here; nothing is converted on the host in this path."""
    import cv2
    cap = cv2.VideoCapture(path)
    frames = []
    while True:
        ok, image = cap.read()
        if not ok:
            break
        frames.append(image)
    cap.release()
    if not frames:
        return np.zeros((0, VID_HEIGHT, VID_WIDTH, 3), dtype=np.uint8)
    return np.stack(frames)


class VideoBatchStager:
    """Pinned uint8 staging buffer [B][T][H][W][3] for one batch + the device-side conversion.

    ``put(b, frames, indices)`` copies the selected frames of clip ``b`` (zero frames behind its end, as the zero-filled
    ``X_vid_final`` of dataset.py:79-80) and returns its ``v_len``; ``to_device()`` uploads the buffer and returns the
    ``(B, 3, H, W, T)`` fp32 tensor = ``frames / 255`` in the layout the drivers slice per frame
    (``v_inputs[:, :, :, :, j]``, q_and_v_eval.py:106)."""

    def __init__(self, batch_size, device, max_frames=MAX_FRAMES, height=VID_HEIGHT, width=VID_WIDTH):
        self.B, self.T, self.H, self.W = batch_size, max_frames, height, width
        self.device = torch.device(device)
        self.host = torch.zeros(batch_size, max_frames, height, width, 3, dtype=torch.uint8).pin_memory() \
            if torch.cuda.is_available() else torch.zeros(batch_size, max_frames, height, width, 3, dtype=torch.uint8)
        self._host_np = self.host.numpy()
        # fp32(double(i) / 255.0): the reference divides in float64 (dataset.py:90) and casts afterwards (:93 of the driver)
        self._lut_host = torch.from_numpy((np.arange(256, dtype=np.float64) / 255.0).astype(np.float32))
        self._lut = None
        self._dev_u8 = None

    def put(self, b, frames_u8, indices=None, rng=_random):
        n = int(frames_u8.shape[0])
        if frames_u8.dtype != np.uint8 or tuple(frames_u8.shape[1:]) != (self.H, self.W, 3):
            raise ValueError(f"expected (n, {self.H}, {self.W}, 3) uint8 frames, got {frames_u8.dtype} {frames_u8.shape}")
        if indices is None:
            indices = subsample_indices(n, rng)
        if len(indices) > self.T:
            raise ValueError(f"{len(indices)} frames after subsampling exceed max_frames = {self.T}")
        slot = self._host_np[b]
        for k, i in enumerate(indices):
            slot[k] = frames_u8[i]
        slot[len(indices):] = 0
        return len(indices)

    def to_device(self, out=None):
        if self.device.type != "cuda":
            raise _lib.VnqaError("videonavqa_b200.data needs a CUDA device (no CPU fallback)")
        if self._lut is None:
            self._lut = self._lut_host.to(self.device)
            self._dev_u8 = torch.empty(self.host.shape, dtype=torch.uint8, device=self.device)
        self._dev_u8.copy_(self.host, non_blocking=True)
        if out is None:
            out = torch.empty(self.B, 3, self.H, self.W, self.T, dtype=torch.float32, device=self.device)
        st = ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        P = lambda t: ctypes.c_void_p(t.data_ptr())     # noqa: E731
        _lib.check(_lib.lib().vnqa_frames_u8_to_video(P(self._dev_u8), P(out), P(self._lut), self.B, self.T, self.H,
                                                      self.W, st), "vnqa_frames_u8_to_video")
        return out

    @property
    def h2d_bytes(self):
        return self.host.numel()
