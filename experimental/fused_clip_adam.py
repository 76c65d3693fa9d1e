"""EXPERIMENTAL (not validated on hardware yet): clip_grad_norm_ + Adam as two launches (fused_clip_adam.cu).

Subclasses torch.optim.Adam so that ``state_dict()`` / ``load_state_dict()`` keep the reference's checkpoint envelope
(eval/q_and_v_eval.py:148-156, 336-346): the state entries are the standard ``step`` / ``exp_avg`` / ``exp_avg_sq``.
Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -shared -Xcompiler -fPIC fused_clip_adam.cu -o libvnqa_exp.so
"""
import ctypes
import math
import os

import torch

CHUNK = 1 << 16


class FusedClipAdam(torch.optim.Adam):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, max_norm=1.0, lib_path=None):
        super().__init__(params, lr=lr, betas=betas, eps=eps)
        self.max_norm = max_norm
        path = lib_path or os.path.join(os.path.dirname(os.path.abspath(__file__)), "libvnqa_exp.so")
        self._lib = ctypes.CDLL(path)
        self._tables = None

    def _build_tables(self, params):
        dev = params[0].device
        tensor_of_chunk, chunk_start = [], []
        for i, p in enumerate(params):
            for s in range(0, p.numel(), CHUNK):
                tensor_of_chunk.append(i)
                chunk_start.append(s)
        def ptrs(ts):
            return torch.tensor([t.data_ptr() for t in ts], dtype=torch.int64, device=dev)
        self._tables = dict(
            toc=torch.tensor(tensor_of_chunk, dtype=torch.int32, device=dev),
            cs=torch.tensor(chunk_start, dtype=torch.int64, device=dev),
            numel=torch.tensor([p.numel() for p in params], dtype=torch.int64, device=dev),
            norm2=torch.zeros(1, dtype=torch.float64, device=dev), n_chunks=len(tensor_of_chunk), ptrs=ptrs)

    @torch.no_grad()
    def step(self, closure=None):
        assert closure is None
        for group in self.param_groups:
            params = [p for p in group["params"] if p.grad is not None]
            if not params:
                continue
            for p in params:
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = torch.tensor(0.0)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["step"] += 1
            step = float(self.state[params[0]]["step"])
            beta1, beta2 = group["betas"]
            if self._tables is None:
                self._build_tables(params)
            t = self._tables
            ptrs = t["ptrs"]
            # gradient tensors are re-allocated every step (zero_grad(set_to_none=True)): refresh their pointer table
            pp, gp = ptrs(params), ptrs([p.grad for p in params])
            mp, vp = ptrs([self.state[p]["exp_avg"] for p in params]), ptrs([self.state[p]["exp_avg_sq"] for p in params])
            rc = self._lib.vnqa_exp_fused_clip_adam(
                ctypes.c_void_p(t["toc"].data_ptr()), ctypes.c_void_p(t["cs"].data_ptr()),
                ctypes.c_void_p(pp.data_ptr()), ctypes.c_void_p(gp.data_ptr()), ctypes.c_void_p(mp.data_ptr()),
                ctypes.c_void_p(vp.data_ptr()), ctypes.c_void_p(t["numel"].data_ptr()), ctypes.c_int(t["n_chunks"]),
                ctypes.c_void_p(t["norm2"].data_ptr()), ctypes.c_float(self.max_norm), ctypes.c_float(group["lr"]),
                ctypes.c_float(beta1), ctypes.c_float(beta2), ctypes.c_float(group["eps"]),
                ctypes.c_float(1.0 - beta1 ** step), ctypes.c_float(math.sqrt(1.0 - beta2 ** step)),
                ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
            if rc:
                raise RuntimeError("vnqa_exp_fused_clip_adam failed")
        return None
