"""CPU oracle of the data path's first slice - TEST INFRASTRUCTURE ONLY.

Restatement of what ``VNQADataset.__getitem__`` (eval/dataset.py:57-93) does to the DECODED frames of one clip, given
the frame indices its ``random.randint`` calls chose: float64 container, frame.transpose(2, 0, 1) into the last axis,
zero frames behind the clip's end, ``/ 255.0`` in float64; the driver's ``.float()`` (eval/q_and_v_eval.py:93) is
applied by the caller.  Pinned against the unmodified reference class by ``oracle/gen_golden_data.py`` ->
``tests/golden/data_small.npz``.  Only tests/ may import this module.
"""
import numpy as np
import torch

MAX_FRAMES = 35          # eval/utils.py:20


def clip_tensor(frames_u8, indices, max_frames=MAX_FRAMES):
    """frames_u8 (n, H, W, 3) uint8 as decoded; indices: the kept frames (dataset.py:81-86).  Returns
    ((3, H, W, max_frames) float64 tensor = X['video'], v_len)."""
    n, H, W, _ = frames_u8.shape
    x = np.zeros((3, H, W, max_frames))                                  # :79-80
    for count, i in enumerate(indices):
        x[:, :, :, count] = frames_u8[i].transpose(2, 0, 1)              # :71, :85
    return torch.from_numpy(x) / 255.0, len(indices)                     # :87, :90
