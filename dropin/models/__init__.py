"""`models` package shim: put this directory's parent (`dropin/`) FIRST on PYTHONPATH and the reference drivers
(eval/q_and_v_eval.py:19-24, eval/q_and_v_test.py:19-24) import the B200 implementations of the three hot-path
classes from `models.*` unchanged.  Every other `models.<name>` (mac, q_concat_*, obj_detector, ...) is out of scope
and falls through to the reference tree when VIDEONAVQA_REFERENCE points at it."""
import os

_ref = os.environ.get("VIDEONAVQA_REFERENCE")
if _ref and os.path.isdir(os.path.join(_ref, "models")):
    __path__.append(os.path.join(_ref, "models"))
