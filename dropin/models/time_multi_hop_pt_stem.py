"""Drop-in for the reference module models/time_multi_hop_pt_stem.py - see videonavqa_b200.models.time_multi_hop_pt_stem."""
from videonavqa_b200.models.time_multi_hop_pt_stem import TimeMultiHopFiLMPretrainedStem  # noqa: F401
