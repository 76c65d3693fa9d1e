"""Golden vector of the data path from the UNMODIFIED reference ``VNQADataset`` (eval/dataset.py).

Build container only (needs /root/reference and OpenCV):    python oracle/gen_golden_data.py

Writes a small synthetic clip (tests/golden/data_clip.mp4, 37 frames of 160 x 208), runs the reference's
``VNQADataset.__getitem__`` on it with ``random.seed(7)`` and stores what identifies its output: v_len, the frame indices
its random.randint calls chose, the sha256 of the float32 video tensor and of the float64 one, and a few sampled values.
"""
import hashlib
import json
import os
import random
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"


def write_clip(path, n_frames=37):
    import cv2
    rng = np.random.RandomState(3)
    w = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"mp4v"), 10.0, (208, 160))
    yy, xx = np.mgrid[0:160, 0:208]
    for f in range(n_frames):
        img = np.stack([(xx + 3 * f) % 256, (yy * 2 + f) % 256, (xx + yy + 5 * f) % 256], axis=-1).astype(np.uint8)
        img[20:60, 30 + 2 * f:70 + 2 * f] = rng.randint(0, 255, size=3)
        w.write(img)
    w.release()


def main():
    gold = os.path.join(ROOT, "tests", "golden")
    clip = os.path.join(gold, "data_clip.mp4")
    write_clip(clip)
    tmp = tempfile.mkdtemp()
    os.makedirs(os.path.join(tmp, "videos")); os.makedirs(os.path.join(tmp, "q"))
    os.symlink(clip, os.path.join(tmp, "videos", "clip0.mp4"))
    np.save(os.path.join(tmp, "q", "clip0.npy"), np.array([5, 9, 2], dtype=np.int64))
    import cv2
    cv2.destroyAllWindows = lambda *a, **k: None          # headless OpenCV raises in dataset.py:77
    sys.path.insert(0, os.path.join(REF, "eval")); sys.path.insert(0, REF)
    cwd = os.getcwd()
    os.chdir(os.path.join(REF, "eval"))
    try:
        from dataset import VNQADataset
    finally:
        os.chdir(cwd)
    ds = VNQADataset(q_dir=os.path.join(tmp, "q"), v_dir=os.path.join(tmp, "videos"), filenames=["clip0"],
                     labels={"clip0": 3})
    random.seed(7)
    X, y = ds[0]
    v64 = X["video"].numpy()
    v32 = X["video"].float().numpy()                   # what the driver feeds the GPU (q_and_v_eval.py:93)
    random.seed(7)                                     # replay the index choices (dataset.py:81-83)
    n_raw = 37
    idx = [random.randint(i, min(i + 4, n_raw) - 1) for i in range(0, n_raw, 4)]
    out = dict(v_len=np.array(X["v_len"]), indices=np.array(idx), n_raw=np.array(n_raw),
               sha256_f32=np.frombuffer(hashlib.sha256(np.ascontiguousarray(v32).tobytes()).digest(), dtype=np.uint8),
               sha256_f64=np.frombuffer(hashlib.sha256(np.ascontiguousarray(v64).tobytes()).digest(), dtype=np.uint8),
               sample_idx=np.array([[0, 0, 0, 0], [1, 80, 100, 5], [2, 159, 207, 9], [0, 33, 47, 34]]),
               clip_sha256=np.frombuffer(hashlib.sha256(open(clip, "rb").read()).digest(), dtype=np.uint8))
    out["sample_val"] = np.array([v32[tuple(i)] for i in out["sample_idx"]], dtype=np.float32)
    np.savez_compressed(os.path.join(gold, "data_small.npz"), **out)
    print("v_len", int(X["v_len"]), "indices", idx, "clip bytes", os.path.getsize(clip), json.dumps(out["sample_val"].tolist()))


if __name__ == "__main__":
    main()
