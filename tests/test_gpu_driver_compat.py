"""The UNMODIFIED reference drivers run end to end against the drop-in package (north_star: "eval/q_and_v_eval.py and
q_and_v_test.py run unchanged"; SURVEY.md §4 "Driver compat") - ``-m gpu``.

oracle/stage_ref.py copies the drivers byte for byte to oracle/_ref/reference (test infrastructure, git-ignored, ships
to the GPU box).  This test builds a tiny synthetic ``../data`` tree (mp4 clips written with OpenCV, encoded questions,
labels / split / q_ids json, a random-init ObjDetectCNN checkpoint), a ``demo`` stub for the un-vendored Faster R-CNN
feature extractor, puts ``dropin/`` first on PYTHONPATH and executes

  * eval/q_and_v_eval.py : 2 training iterations (forward, CrossEntropy(sum), backward, clip_grad_norm, Adam, checkpoint
    envelope written) + 1 validation iteration, for the attention, global-pooling and multi-hop models;
  * eval/q_and_v_test.py : restores that checkpoint and runs the test split, whose last batch is short and therefore
    zero-padded with v_len = 1, q_len = 1 clips (eval/q_and_v_test.py:80-87).

The drivers' ``from models.obj_detector import ObjDetectCNN`` (eval/utils.py:4) resolves to the B200 frozen stem too
(dropin/models/obj_detector.py): it is built by the unmodified ``get_object_detector()``, loads the checkpoint written
here by the REFERENCE's class and runs once per frame (q_and_v_eval.py:104-110).
"""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "reference")

DEMO_STUB = '''
"""Stub of the un-vendored faster-rcnn.pytorch fork's demo.get_frcnn_feature_extractor (eval/q_and_v_eval.py:17):
a frozen strided convolution producing the 128 x 40 x 52 map the reference's ObjDetectCNN expects."""
import torch
import torch.nn as nn


def get_frcnn_feature_extractor(path):
    torch.manual_seed(7)
    net = nn.Sequential(nn.Conv2d(3, 128, kernel_size=4, stride=4), nn.ReLU())
    for p in net.parameters():
        p.requires_grad_(False)
    return net.eval()
'''


SITE_STUB = '''
"""Environment shim (not a driver change): this image ships headless OpenCV, whose destroyAllWindows() raises; the
reference's dataset.py calls it after every clip (eval/dataset.py:77)."""
try:
    import cv2
    cv2.destroyAllWindows = lambda *a, **k: None
except Exception:
    pass
# torch 1.4 (the reference's pin) allowed indexing a CPU tensor with a CUDA index tensor; the test driver does it
# (eval/q_and_v_test.py:115-117, q_ids is never moved to the GPU).  PyTorch 2 raises - restore the old behaviour.
import torch
_getitem = torch.Tensor.__getitem__


def _getitem_compat(self, idx):
    if isinstance(idx, torch.Tensor) and idx.is_cuda and not self.is_cuda:
        idx = idx.cpu()
    return _getitem(self, idx)


torch.Tensor.__getitem__ = _getitem_compat
'''


def _make_data(tmp, n_train=8, n_val=4, n_test=6):
    import cv2
    sys.path.insert(0, os.path.join(REF, "models"))
    try:
        from obj_detector import ObjDetectCNN          # the reference's own class, for a loadable checkpoint
    finally:
        sys.path.pop(0)
    data = os.path.join(tmp, "data")
    os.makedirs(os.path.join(data, "videos"))
    os.makedirs(os.path.join(data, "encoded_questions"))
    os.makedirs(os.path.join(tmp, "eval"))
    rng = np.random.RandomState(0)
    ids = {"train": [f"tr{i}" for i in range(n_train)], "val": [f"va{i}" for i in range(n_val)],
           "test": [f"te{i}" for i in range(n_test)]}
    labels, q_ids = {}, {}
    fourcc = cv2.VideoWriter_fourcc(*"mp4v")
    for split in ids.values():
        for name in split:
            n_frames = int(rng.randint(12, 141))                       # 3 .. 35 frames after 1-of-4 subsampling
            w = cv2.VideoWriter(os.path.join(data, "videos", name + ".mp4"), fourcc, 10.0, (208, 160))
            base = rng.randint(0, 255, size=(160, 208, 3)).astype(np.uint8)
            for f in range(n_frames):
                w.write(np.roll(base, 3 * f, axis=1))
            w.release()
            q = rng.randint(1, 134, size=int(rng.randint(5, 31))).astype(np.int64)
            np.save(os.path.join(data, "encoded_questions", name + ".npy"), q)
            labels[name] = int(rng.randint(0, 70))
            q_ids[name] = int(rng.randint(0, 35))
    json.dump(labels, open(os.path.join(data, "labels.json"), "w"))
    json.dump(ids, open(os.path.join(data, "split.json"), "w"))
    json.dump(q_ids, open(os.path.join(data, "q_ids.json"), "w"))
    torch.manual_seed(3)
    det = ObjDetectCNN(nb_classes=27, num_filters=512, tail_hidden_dim=1024, tail_dropout_p=0, logits=True,
                       pretrained_features=True)
    torch.save({"state_dict": det.state_dict()}, os.path.join(data, "obj_detect.pt"))
    os.makedirs(os.path.join(tmp, "stubs"))
    open(os.path.join(tmp, "stubs", "demo.py"), "w").write(DEMO_STUB)
    open(os.path.join(tmp, "stubs", "sitecustomize.py"), "w").write(SITE_STUB)
    return ids


def _run(tmp, script, *argv):
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([os.path.join(tmp, "stubs"), os.path.join(ROOT, "dropin"), ROOT,
                                         os.path.join(REF, "eval")])
    env["VIDEONAVQA_REFERENCE"] = REF
    env["TORCH_FORCE_NO_WEIGHTS_ONLY_LOAD"] = "1"      # the reference's checkpoints hold numpy scalars (torch 1.4 era)
    r = subprocess.run([sys.executable, os.path.join(REF, "eval", script), *argv], cwd=os.path.join(tmp, "eval"), env=env,
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, f"{script} failed:\n{r.stdout[-3000:]}\n{r.stderr[-3000:]}"
    return r.stdout


@pytest.mark.parametrize("model,extra", [
    ("film_attn_pt", ["--at_hidden_size", "64"]),
    ("film_gp_pt", ["--num_tail_channels", "16"]),
    ("time_multi_hop", ["--num_tail_channels", "16"]),
])
def test_unmodified_reference_drivers_run_against_the_dropin(tmp_path, model, extra):
    assert os.path.isdir(REF), "oracle/_ref/reference is missing: run __graft_entry__.build() (oracle/stage_ref.py)"
    tmp = str(tmp_path)
    _make_data(tmp)
    common = ["--model", model, "--batch_size", "4", "--loss_reduction", "sum", "--num_res_blocks", "2",
              "--num_res_block_channels", "64", "--num_input_channels", "512", "--hidden_size", "64", "--embed_size",
              "32", "--checkpoint_path", "ckpt.pt", "--num_workers", "0", *extra]
    # which classes will the drivers' own import statements get?  (same environment as the driver runs below)
    probe = os.path.join(tmp, "probe.py")
    open(probe, "w").write("import sys; sys.path.insert(0, '.')\nfrom models.obj_detector import ObjDetectCNN as O\n"
                           "from models.film_attn_pt_stem import FiLMAttnPretrainedStem as A\n"
                           "from models.mac import MACNetwork as M\nprint(O.__module__, A.__module__, M.__module__)\n")
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(ROOT, "dropin"), ROOT]), VIDEONAVQA_REFERENCE=REF)
    mods = subprocess.run([sys.executable, probe], env=env, capture_output=True, text=True, check=True).stdout.split()
    assert mods == ["videonavqa_b200.models.obj_detector", "videonavqa_b200.models.film_attn_pt_stem", "models.mac"], mods
    out = _run(tmp, "q_and_v_eval.py", *common, "--num_epochs", "1", "--stats_after_every", "1")
    assert "8 train examples, 4 validation examples" in out
    assert "Average loss after 2 iterations in epoch 1" in out           # two training iterations ran
    assert "Train Epoch: 0" in out and "Validation:" in out
    ck = os.path.join(tmp, "eval", "e0_ckpt.pt")
    assert os.path.exists(ck)
    saved = torch.load(ck, map_location="cpu", weights_only=False)
    assert set(saved) == {"epoch", "model", "state_dict", "train_f1w", "train_f1micro", "optimizer"}
    assert not any(k.startswith("conv1x1_layers") for k in saved["state_dict"])           # quirk Q1
    if model != "time_multi_hop":
        assert "film_layer.0.weight_ih_l0" in saved["state_dict"]                          # quirk Q2 (CUDA branch)
    else:
        assert "q_encoder.weight_ih_l0" in saved["state_dict"] and "decoder_norm.weight" in saved["state_dict"]
    assert all(np.isfinite(v.float().numpy()).all() for v in saved["state_dict"].values())
    # the test driver restores that checkpoint; 6 test clips at batch 4 -> the second batch is padded with 2 dummies
    os.replace(ck, os.path.join(tmp, "eval", "ckpt.pt"))
    out = _run(tmp, "q_and_v_test.py", *common)
    assert "Restored checkpoint from epoch 1" in out and "Testing:" in out
    t = np.load(os.path.join(tmp, "eval", "t_ckpt.pt.npy"))
    p = np.load(os.path.join(tmp, "eval", "p_ckpt.pt.npy"))
    assert t.shape == (6,) and p.shape == (6,)                                             # pads were sliced off
