"""Drop-in for the reference module models/film_global_pooling_pt_stem.py - see videonavqa_b200.models.film_global_pooling_pt_stem."""
from videonavqa_b200.models.film_global_pooling_pt_stem import FiLMGlobalPoolingPretrainedStem  # noqa: F401
