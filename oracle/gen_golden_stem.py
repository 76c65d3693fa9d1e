"""Golden vector of the frozen stem from the UNMODIFIED reference module (models/obj_detector.py).

Build container only (needs /root/reference):    python oracle/gen_golden_stem.py

Builds the reference ``ObjDetectCNN`` the way eval/utils.py:42-51 does (pretrained_features=True, eval()), loads the
seeded parameters of ``oracle.stem_oracle.synthetic_stem_case`` into it, runs its forward on the seeded input and stores
the output together with the parameters (the input is regenerated from its seed; a sha256 guards against RNG drift).
"""
import hashlib
import importlib.util
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.stem_oracle import synthetic_stem_case  # noqa: E402

REF = "/root/reference/models/obj_detector.py"
CASE = dict(n=3, h=16, w=52, filters=32, seed=21)


def main():
    spec = importlib.util.spec_from_file_location("_ref_obj_detector", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    P, x = synthetic_stem_case(**CASE)
    m = mod.ObjDetectCNN(nb_classes=27, num_filters=CASE["filters"], tail_hidden_dim=64, tail_dropout_p=0, logits=True,
                         pretrained_features=True)
    missing = m.load_state_dict(P, strict=False)
    assert not missing.unexpected_keys, missing
    assert all(k.startswith(("fc_tail", "bn_tail")) or k.endswith("num_batches_tracked") for k in missing.missing_keys)
    m.eval()
    with torch.no_grad():
        y = m(x)
        y64 = m.double()(x.double())
    out = {"p." + k: v.numpy() for k, v in P.items()}
    out["y"] = y.numpy()
    out["y64"] = y64.numpy()
    out["x_sha256"] = np.frombuffer(hashlib.sha256(x.numpy().tobytes()).digest(), dtype=np.uint8)
    out["case"] = np.array([CASE[k] for k in ("n", "h", "w", "filters", "seed")])
    out["state_dict_keys"] = np.array(list(m.state_dict().keys()))
    path = os.path.join(ROOT, "tests", "golden", "stem_small.npz")
    np.savez_compressed(path, **out)
    print(path, "%.2f MB" % (os.path.getsize(path) / 1e6), "out", tuple(y.shape), "max|y|", float(y.abs().max()),
          "fp32 vs fp64", float((y.double() - y64).abs().max()))


if __name__ == "__main__":
    main()
