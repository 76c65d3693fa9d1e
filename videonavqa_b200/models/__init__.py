"""Drop-in replacements for the reference's ``models`` package (hot-path classes only)."""
from .film_attn_pt_stem import FiLMAttnPretrainedStem
from .film_global_pooling_pt_stem import FiLMGlobalPoolingPretrainedStem
from .time_multi_hop_pt_stem import TimeMultiHopFiLMPretrainedStem

__all__ = ["FiLMAttnPretrainedStem", "FiLMGlobalPoolingPretrainedStem", "TimeMultiHopFiLMPretrainedStem"]
