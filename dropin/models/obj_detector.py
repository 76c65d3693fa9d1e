"""Drop-in for the reference module models/obj_detector.py (frozen stem) - see videonavqa_b200.models.obj_detector."""
from videonavqa_b200.models.obj_detector import ObjDetectCNN  # noqa: F401
