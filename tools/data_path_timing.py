"""Timing of the data path's device slice at the benchmark batch (64 clips x 35 frames of 160 x 208): H2D of the uint8
staging buffer + vnqa_frames_u8_to_video, against uploading the reference's fp32 tensor (what the driver does after
its float64 -> fp32 cast on the host)."""
import sys
import torch
sys.path.insert(0, ".")
from videonavqa_b200 import data

B = 64
st = data.VideoBatchStager(B, "cuda")
st.host.random_(0, 256)
out = st.to_device()
ref_host = torch.empty(B, 3, 160, 208, 35, dtype=torch.float32).pin_memory()
ref_dev = torch.empty_like(out)


def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


a = t(lambda: st.to_device(out))
st._dev_u8.copy_(st.host)
import ctypes
from videonavqa_b200 import _lib
P = lambda x: ctypes.c_void_p(x.data_ptr())     # noqa: E731
k = t(lambda: _lib.lib().vnqa_frames_u8_to_video(P(st._dev_u8), P(out), P(st._lut), B, 35, 160, 208,
                                                  ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
b = t(lambda: ref_dev.copy_(ref_host, non_blocking=True))
print(f"uint8 staging: H2D {st.h2d_bytes / 1e6:.0f} MB + kernel = {a:.2f} ms (kernel alone {k:.3f} ms, "
      f"{(st.h2d_bytes + out.numel() * 4) / k / 1e6:.0f} GB/s)")
print(f"reference layout: H2D of the fp32 tensor {ref_host.numel() * 4 / 1e6:.0f} MB = {b:.2f} ms (pinned; the reference's "
      f"DataLoader hands over pageable float64, 2x that size, and casts on the host first)")
