"""videonavqa_b200 - B200-native (sm_100a) implementation of VideoNavQA's FiLM video-QA hot path.

Drop-in modules live in ``videonavqa_b200.models`` (same class names, constructor arguments, ``forward`` /
``init_hidden`` and ``state_dict`` layout as the reference's ``models`` package); all arithmetic runs in
``libvnqa_b200.so`` (hand-written CUDA, C ABI in ``include/vnqa_b200.h``).
"""
__version__ = "0.1.0"
