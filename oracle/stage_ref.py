"""Stages the UNMODIFIED reference drivers (and the out-of-scope `models.*` modules they import) under oracle/_ref/ so
that the driver-compatibility test can execute them on the GPU box, where /root/reference does not exist.

TEST INFRASTRUCTURE ONLY.  Nothing is edited: files are copied byte for byte from the reference tree into
oracle/_ref/reference/ (git-ignored, travels to the GPU box with the snapshot; never part of the product, never
imported by videonavqa_b200).  Run by __graft_entry__.build() when /root/reference is present:

    python oracle/stage_ref.py [/root/reference]

Staged:
  eval/q_and_v_eval.py, eval/q_and_v_test.py   the train / validate and test drivers
  eval/utils.py, eval/dataset.py               constants, per-class accuracy, the mp4 / question loader
  models/__init__.py, mac.py, q_concat_*.py, obj_detector.py   imported by the drivers, out of scope for the product
  models_hot/film_attn_pt_stem.py, film_global_pooling_pt_stem.py, time_multi_hop_pt_stem.py
                                               the reference's own hot path, for the CPU baseline arm of bench.py only
The drivers get the three hot-path modules from dropin/models (the B200 implementations), never from models_hot/.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
FILES = ["eval/q_and_v_eval.py", "eval/q_and_v_test.py", "eval/utils.py", "eval/dataset.py", "models/__init__.py",
         "models/mac.py", "models/q_concat_cnn2d_lstm.py", "models/q_concat_cnn3d.py", "models/obj_detector.py"]
# the reference's OWN implementation of the hot path, staged under a different directory name (models_hot/) so that it
# can never shadow dropin/models: bench.py's CPU arm (--impl reference / cpu_baseline, kind "reference") loads these
# files by path and times them unmodified on the host cores
HOT = {"models/film_attn_pt_stem.py": "models_hot/film_attn_pt_stem.py",
       "models/film_global_pooling_pt_stem.py": "models_hot/film_global_pooling_pt_stem.py",
       "models/time_multi_hop_pt_stem.py": "models_hot/time_multi_hop_pt_stem.py"}


def stage(ref="/root/reference"):
    dst_root = os.path.join(HERE, "_ref", "reference")
    manifest = {}
    for rel, rel_dst in list(zip(FILES, FILES)) + list(HOT.items()):
        src = os.path.join(ref, rel)
        dst = os.path.join(dst_root, rel_dst)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        manifest[rel_dst] = hashlib.sha256(open(dst, "rb").read()).hexdigest()
    with open(os.path.join(dst_root, "MANIFEST.json"), "w") as f:
        json.dump({"source": ref, "sha256": manifest}, f, indent=1)
    return dst_root


if __name__ == "__main__":
    print(stage(sys.argv[1] if len(sys.argv) > 1 else "/root/reference"))
