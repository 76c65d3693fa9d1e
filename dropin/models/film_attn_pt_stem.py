"""Drop-in for the reference module models/film_attn_pt_stem.py - see videonavqa_b200.models.film_attn_pt_stem."""
from videonavqa_b200.models.film_attn_pt_stem import FiLMAttnPretrainedStem  # noqa: F401
