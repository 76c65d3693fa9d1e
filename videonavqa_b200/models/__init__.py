"""Drop-in replacements for the reference's ``models`` package (the three hot-path classes and the frozen stem)."""
from .film_attn_pt_stem import FiLMAttnPretrainedStem
from .film_global_pooling_pt_stem import FiLMGlobalPoolingPretrainedStem
from .obj_detector import ObjDetectCNN
from .time_multi_hop_pt_stem import TimeMultiHopFiLMPretrainedStem

__all__ = ["FiLMAttnPretrainedStem", "FiLMGlobalPoolingPretrainedStem", "TimeMultiHopFiLMPretrainedStem",
           "ObjDetectCNN"]
