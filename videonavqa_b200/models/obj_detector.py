"""Drop-in for the reference's frozen stem ``models.obj_detector.ObjDetectCNN`` (models/obj_detector.py:5-104) -
SURVEY.md 8(f) rank 2, the step directly before the FiLM path.

The drivers build it with ``get_object_detector()`` (eval/utils.py:42-51: 512 filters, ``pretrained_features=True``,
``eval()``), load a checkpoint into it and call it under ``torch.no_grad()`` once per frame on the VGG front's
``(B, 128, 40, 52)`` maps (eval/q_and_v_eval.py:104-110); what comes back, ``(B, 512, 10, 13)``, is the ``v_input`` of
the FiLM modules.  That inference path is what runs here, in hand-written CUDA:

    bn_input (eval affine) + NCHW -> NHWC + operand rounding          csrc/stem.cu        (HBM-bound)
    conv11 -> conv12 -> bn1 -> relu -> pool1      at 40 x 52          csrc/conv_gemm.cu   EPI_BIAS, EPI_BIAS_AFFINE_RELU + pool
    conv21 -> conv22 -> bn2 -> relu -> pool2      at 20 x 26          (same tcgen05 implicit GEMM, 208-pixel tiles =
    conv31 -> conv32 -> bn3 -> relu               at 10 x 13           2 / 4 / 8 images x 2 rows x W)
    NHWC -> NCHW fp32                                                 csrc/stem.cu

BatchNorm (eval), ReLU and the 2 x 2 max-pool are fused into the producing convolution's epilogue.  The classifier tail
(pool3 / fc_tail1 / bn_tail1 / fc_tail2, obj_detector.py:88-101) and training of the detector are outside the path: the
parameters are registered (same names, shapes and order, so the reference's checkpoints load) but ``forward`` raises
for ``pretrained_features=False`` or ``train()`` mode - there is no PyTorch fallback.
"""
from __future__ import annotations

import os

import torch
import torch.nn as nn

from .. import _lib, engine, ops
from ._base import DEFAULT_OPERANDS


class ObjDetectCNN(nn.Module):
    IN_CHANNELS = 128                      # VGG-16 conv2_2 maps (obj_detector.py:21)

    def __init__(self, nb_classes, num_filters=128, tail_hidden_dim=256, tail_dropout_p=0.5, logits=False,
                 pretrained_features=False):
        super().__init__()
        if num_filters % 32:
            raise ValueError(f"videonavqa_b200 needs num_filters to be a multiple of 32 (got {num_filters})")
        self.logits = logits
        self.pretrained_features = pretrained_features
        # Same attribute names, shapes and REGISTRATION ORDER as obj_detector.py:21-44 (the reference's checkpoints and
        # optimizer states are keyed by them): input BatchNorm, then three stages of conv-conv-BatchNorm-pool at
        # 40 x 52 / 20 x 26 / 10 x 13, then the classifier tail (registered for state_dict parity, not run here).
        F = num_filters
        self.bn_input = nn.BatchNorm2d(self.IN_CHANNELS)
        for stage, cin in ((1, self.IN_CHANNELS), (2, F), (3, F)):
            setattr(self, f"conv{stage}1", nn.Conv2d(cin, F, kernel_size=3, padding=1))
            setattr(self, f"conv{stage}2", nn.Conv2d(F, F, kernel_size=3, padding=1))
            setattr(self, f"bn{stage}", nn.BatchNorm2d(F))
            setattr(self, f"pool{stage}", nn.MaxPool2d(kernel_size=2, stride=2))
        self.fc_tail1 = nn.Linear(F * 6 * 5, tail_hidden_dim)
        self.bn_tail1 = nn.BatchNorm1d(tail_hidden_dim)
        self.fc_tail2 = nn.Linear(tail_hidden_dim, nb_classes)
        self.dropout = nn.Dropout(p=tail_dropout_p)
        self.relu = nn.ReLU(inplace=True)
        self.num_filters = num_filters
        for module in self.modules():
            self.weights_init(module)
        self.__dict__["_weight_cache"] = engine.WeightCache()

    def weights_init(self, m):
        """obj_detector.py:50-53 (the LSTM branch of the reference's copy never fires: the stem has none)."""
        if isinstance(m, (nn.Linear, nn.Conv2d)):
            nn.init.xavier_uniform_(m.weight.data)
            m.bias.data.fill_(0.0)

    # ---- precision / caches ------------------------------------------------------------------------------------
    @property
    def precision(self):
        """Operand storage of the convolutions ('fp16' / 'tf32'), as for the FiLM modules: VNQA_OPERANDS or
        ``set_operands()``."""
        p = self.__dict__.get("_precision")
        if p is None:
            p = engine.Precision(os.environ.get("VNQA_OPERANDS", DEFAULT_OPERANDS))
            self.__dict__["_precision"] = p
        return p

    def set_operands(self, operands):
        self.__dict__["_precision"] = engine.Precision(operands)
        return self

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        if "_weight_cache" in self.__dict__:
            self._weight_cache.invalidate()
        return out

    def load_state_dict(self, *args, **kwargs):
        out = super().load_state_dict(*args, **kwargs)
        self._weight_cache.invalidate()
        return out

    def _packed(self, conv, prec):
        return self._weight_cache.get(prec.key + str(id(conv)), conv.weight,
                                      lambda w: ops.pack_conv_weight(w.detach(), act=prec.act))

    def _affine(self, name, bn):
        def build(_):
            return ops.bn_eval_affine(bn.weight.detach(), bn.bias.detach(), bn.running_mean, bn.running_var, bn.eps)
        # rebuilt when the scale parameter OR the running statistics change (both are written by load_state_dict)
        key = (name, bn.bias._version, bn.running_mean._version, bn.running_var._version, bn.running_mean.data_ptr())
        return self._weight_cache.get(key, bn.weight, build)

    # ---- forward -----------------------------------------------------------------------------------------------
    def forward(self, inputs):
        if not self.pretrained_features:
            raise NotImplementedError(
                "videonavqa_b200.ObjDetectCNN implements the frozen-stem path (pretrained_features=True, "
                "models/obj_detector.py:84-85); the object-classification tail is outside the accelerated path")
        if self.training:
            raise NotImplementedError(
                "videonavqa_b200.ObjDetectCNN runs the frozen stem in eval() mode only, as the reference's drivers do "
                "(eval/utils.py:50); training the detector is outside the accelerated path")
        return self.features(inputs)

    @torch.no_grad()
    def features(self, inputs, nhwc=False):
        """(N, 128, 40, 52) fp32 -> (N, F, 10, 13) fp32.  With ``nhwc=True`` the result stays in the kernels' own layout
        and operand type, [N][130][F], ready to be the FiLM trunk's input without another relayout."""
        if not inputs.is_cuda:
            raise _lib.VnqaError("videonavqa_b200 ops need CUDA tensors (no CPU fallback)")
        if inputs.dim() != 4 or inputs.shape[1] != self.IN_CHANNELS or inputs.shape[3] != 52 or inputs.shape[2] % 8:
            raise ValueError("ObjDetectCNN expects (N, 128, H, 52) maps with H a multiple of 8 (the reference feeds "
                             f"(N, 128, 40, 52)); got {tuple(inputs.shape)}")
        prec = self.precision
        act, F = prec.act, self.num_filters
        N, _, H, W = inputs.shape
        x = inputs.float().contiguous()
        s0, t0 = self._affine("bn_input", self.bn_input)
        h = ops.nchw_to_nhwc_affine(x, s0, t0, act=act)                       # obj_detector.py:70
        stages = ((self.conv11, self.conv12, self.bn1, "bn1", True),          # :72-75
                  (self.conv21, self.conv22, self.bn2, "bn2", True),          # :77-80
                  (self.conv31, self.conv32, self.bn3, "bn3", False))         # :82-85 (pool3 is not reached)
        for ca, cb, bn, name, pool in stages:
            cin = ca.weight.shape[1]
            mid = torch.empty(N, H * W, F, device=x.device, dtype=prec.dtype)
            ops.conv_gemm(self._packed(ca, prec), h, taps=9, cin=cin, cout=F, n_img=N, epi=_lib.EPI_BIAS, out=mid,
                          bias=ca.bias.detach(), act=act, img_hw=(H, W))
            sc, sh = self._affine(name, bn)
            last = not pool
            Ho, Wo = (H // 2, W // 2) if pool else (H, W)
            f32_out = last and not nhwc
            out = torch.empty(N, Ho * Wo, F, device=x.device, dtype=torch.float32 if f32_out else prec.dtype)
            ops.conv_gemm(self._packed(cb, prec), mid, taps=9, cin=F, cout=F, n_img=N, epi=_lib.EPI_BIAS_AFFINE_RELU,
                          out=out, bias=cb.bias.detach(), scale=sc, shift=sh, act=act, img_hw=(H, W), pool=pool,
                          out_f32=f32_out and prec.fp16, no_round=f32_out)
            h, H, W = out, Ho, Wo
        if nhwc:
            return h
        return ops.nhwc_to_nchw(h, (H, W))
