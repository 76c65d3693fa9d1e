"""CPU-side checks (no GPU): the C-ABI library loads and exports every symbol include/vnqa_b200.h declares, the
Python binding declares the same set, the drop-in modules keep the reference's state_dict layout, and the ragged
bookkeeping matches the reference loop."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "vnqa_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vnqa_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from videonavqa_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "build the library first (__graft_entry__.build())"
    handle = ctypes.CDLL(_lib.LIB_PATH)
    syms = header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(handle, s), f"libvnqa_b200.so does not export {s}"
    assert sorted(_lib.SIGNATURES) == syms, set(_lib.SIGNATURES) ^ set(syms)
    assert handle.vnqa_abi_version() == 3


def test_abi_rejects_bad_arguments_before_touching_the_device():
    """Error behaviour of the C ABI: argument checks come first, return -1 and leave a thread-local message
    (include/vnqa_b200.h); none of these calls reaches a CUDA API, so they run on the CPU-only build container."""
    from videonavqa_b200 import _lib
    from videonavqa_b200._lib import ConvArgs
    L = _lib.lib()
    msg = lambda: L.vnqa_last_error().decode()     # noqa: E731
    assert L.vnqa_conv_gemm(None, None) == -1 and "null args" in msg()
    a = ConvArgs()
    a.taps, a.rows_pad, a.cout, a.n_img = 5, 128, 128, 1
    assert L.vnqa_conv_gemm(ctypes.byref(a), None) == -1 and "taps must be 1 or 9" in msg()
    a.taps, a.rows_pad = 9, 100
    assert L.vnqa_conv_gemm(ctypes.byref(a), None) == -1 and "rows_pad" in msg()
    a.rows_pad, a.n_img = 128, 0
    assert L.vnqa_conv_gemm(ctypes.byref(a), None) == -1 and "empty batch" in msg()
    a.n_img = 1
    assert L.vnqa_conv_gemm(ctypes.byref(a), None) == -1 and "null operand" in msg()
    buf = (ctypes.c_float * 64)()
    pb = ctypes.cast(buf, ctypes.c_void_p)
    a.w_packed = a.x = a.out = pb.value
    a.act_dtype = 7
    assert L.vnqa_conv_gemm(ctypes.byref(a), None) == -1 and "act_dtype" in msg()
    a.act_dtype, a.cin, a.img_h, a.img_w = 1, 64, 10, 50
    assert L.vnqa_conv_gemm(ctypes.byref(a), None) == -1 and "img_w must be 13, 26 or 52" in msg()
    a.img_w, a.img_h = 26, 7
    assert L.vnqa_conv_gemm(ctypes.byref(a), None) == -1 and "img_h must be even" in msg()
    a.img_h, a.pool, a.epi = 20, 1, 0
    assert L.vnqa_conv_gemm(ctypes.byref(a), None) == -1 and "pool needs BIAS_AFFINE_RELU" in msg()
    assert L.vnqa_frames_u8_to_video(None, None, None, 1, 1, 1, 1, None) == -1 and "null pointer" in msg()
    assert L.vnqa_frames_u8_to_video(pb, pb, pb, 0, 35, 160, 208, None) == -1 and "bad sizes" in msg()
    assert L.vnqa_nchw_to_nhwc_affine(pb, pb, pb, None, 1, 2, 4, 1, None) == -1 and "scale and shift go together" in msg()
    assert L.vnqa_nhwc_to_nchw(None, pb, 1, 2, 4, 0, None) == -1
    assert L.vnqa_frames_blocked_to_nhwc(pb, pb, pb, 1, 33, 35, 0, 0, 1, None) == -1 and "multiple of 32" in msg()
    assert L.vnqa_sumsq_partials(pb, 3, pb, None) == -1 and "4k floats" in msg()
    assert L.vnqa_clip_adam(pb, pb, pb, pb, 64, pb, 1, 1.0, 1.0, 1e-3, 0.9, 0.999, 1e-8, 0, None, None) == -1 \
        and "step counts from 1" in msg()
    assert L.vnqa_clip_adam_dev(pb, pb, pb, pb, 64, pb, 1, 1.0, 1.0, 1e-3, 0.9, 0.999, 1e-8, None, None, None) == -1
    assert L.vnqa_conv_wgrad_ws_bytes(100, 128, 9, None) < 0 or True      # (returns a size or a negative code; no device)


def test_conv_args_struct_matches_header_field_order():
    from videonavqa_b200._lib import ConvArgs
    text = open(os.path.join(ROOT, "include", "vnqa_b200.h")).read()
    body = text[text.index("typedef struct {"):text.index("} vnqa_conv_args;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = re.findall(r"\b(\w+);", body)
    assert fields == [f[0] for f in ConvArgs._fields_]


def test_missing_library_fails_loudly(monkeypatch):
    from videonavqa_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libvnqa_b200.so")
    with pytest.raises(_lib.VnqaError):
        _lib.lib()


def test_state_dict_layout_matches_reference_golden():
    """Keys/shapes/order of the reference state_dict (golden file written by the reference on a CPU host, where
    film_layer is a plain list - Q2; the B200 build always has the CUDA-branch keys in between)."""
    from videonavqa_b200.models import FiLMAttnPretrainedStem, FiLMGlobalPoolingPretrainedStem
    from oracle.gen_golden import CASES
    for name, cls in (("attn_small", FiLMAttnPretrainedStem), ("gp_small", FiLMGlobalPoolingPretrainedStem)):
        z = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
        ref_keys = [str(k) for k in z["state_dict_keys"]]
        m = cls(**CASES[name][1])
        sd = m.state_dict()
        mine = [k for k in sd if not k.startswith("film_layer.")]
        assert mine == ref_keys
        for k in ref_keys:
            if "p." + k in z.files:
                assert tuple(sd[k].shape) == z["p." + k].shape, k
        film_keys = [k for k in sd if k.startswith("film_layer.")]
        assert film_keys == ["film_layer.0.weight_ih_l0", "film_layer.0.weight_hh_l0", "film_layer.0.bias_ih_l0",
                             "film_layer.0.bias_hh_l0", "film_layer.1.weight", "film_layer.1.bias"]
        assert list(sd).index("film_layer.0.weight_ih_l0") == list(sd).index("bn_init.num_batches_tracked") + 1
        assert not any(k.startswith("conv1x1_layers") for k in sd)            # Q1
        assert len(m.conv1x1_layers) == CASES[name][1]["num_res_blocks"]
        # parameters() order is the Adam-state order of the checkpoint envelope
        assert [n for n, _ in m.named_parameters()] == [k for k in sd if "running" not in k and "tracked" not in k]


def test_param_count_eval_sh_config():
    """70 868 424 registered parameters for the shipped attention configuration."""
    from videonavqa_b200.models import FiLMAttnPretrainedStem
    m = FiLMAttnPretrainedStem(batch_size=2, q_embedding_size=128, nb_classes=70, num_input_channels=512,
                               num_res_block_channels=1024, num_res_blocks=5, hidden_size=128, at_hidden_size=128,
                               max_num_frames=35, vocab_size=134)
    assert sum(p.numel() for p in m.parameters()) == 70868424
    assert sum(p.numel() for c in m.conv1x1_layers for p in c.parameters()) == 5248000
    # LSTM init of the reference: forget-gate bias slice = 1, b_ih = 0
    lstm = m.film_layer[0]
    assert torch.all(lstm.bias_hh_l0[128:256] == 1) and torch.all(lstm.bias_ih_l0 == 0)
    assert tuple(m.film_hidden[0].shape) == (1, 2, 128)


def test_forward_refuses_cpu_tensors():
    from videonavqa_b200.models import FiLMGlobalPoolingPretrainedStem
    m = FiLMGlobalPoolingPretrainedStem(2, 16, 7, 32, 32, 8, 1, 32, vocab_size=20)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(2, 32, 10, 13, 35), torch.ones(2, 5, dtype=torch.long), torch.tensor([3, 2]), torch.tensor([2, 2]))


def test_plan_matches_reference_frame_loop():
    from oracle.film_oracle import effective_batch_sizes
    from videonavqa_b200.engine import Plan
    v_lens = [30, 17, 17, 9, 1]
    cts = effective_batch_sizes(torch.tensor(v_lens), 35)
    # Plan needs a device for its index tensors; build the host part only
    class _CpuPlan(Plan):
        def __init__(self, v, B, T):
            try:
                super().__init__(v, B, T, torch.device("cpu"))
            except RuntimeError:
                pass
    try:
        p = Plan(v_lens, 5, 35, torch.device("cpu"))
    except RuntimeError:
        pytest.skip("pinned memory unavailable on this host")
    assert p.n_frames == 30 and p.n_img == sum(v_lens)
    assert p.frame_ct_host.tolist() == cts
    img_of = p.img_of.view(5, 35)
    for b, vl in enumerate(v_lens):
        assert (img_of[b, :vl] >= 0).all() and (img_of[b, vl:] == -1).all()
    # frame-major ordering and the two row maps
    assert p.img_frame.tolist() == sorted(p.img_frame.tolist())
    assert int(p.film_row[img_of[1, 3]]) == 3 * 5 + 1
    assert int(p.feat_row[img_of[1, 3]]) == 1 * 35 + 3
    with pytest.raises(ValueError):
        Plan([3, 5], 2, 35, torch.device("cpu"))


def test_product_and_oracle_input_generators_agree():
    from oracle.film_oracle import synthetic_batch as a
    from videonavqa_b200.synth import synthetic_batch as b
    for kw in (dict(B=3, c_in=32, seed=4), dict(B=5, c_in=32, seed=9, ragged=True)):
        for x, y in zip(a(**kw), b(**kw)):
            assert torch.equal(x, y)


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under videonavqa_b200/ may import it."""
    import glob
    for f in glob.glob(os.path.join(ROOT, "videonavqa_b200", "**", "*.py"), recursive=True):
        src = open(f).read()
        assert "import oracle" not in src and "from oracle" not in src, f


def test_dropin_models_package_shim():
    """With dropin/ first on the path the reference drivers' imports resolve to the B200 classes."""
    import importlib
    import sys
    sys.path.insert(0, os.path.join(ROOT, "dropin"))
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k == "models" or k.startswith("models.")}
    try:
        a = importlib.import_module("models.film_attn_pt_stem")
        g = importlib.import_module("models.film_global_pooling_pt_stem")
        t = importlib.import_module("models.time_multi_hop_pt_stem")
        import videonavqa_b200.models as real
        assert a.FiLMAttnPretrainedStem is real.FiLMAttnPretrainedStem
        assert g.FiLMGlobalPoolingPretrainedStem is real.FiLMGlobalPoolingPretrainedStem
        assert t.TimeMultiHopFiLMPretrainedStem is real.TimeMultiHopFiLMPretrainedStem
        assert importlib.import_module("models.obj_detector").ObjDetectCNN is real.ObjDetectCNN      # frozen stem
    finally:
        sys.path.remove(os.path.join(ROOT, "dropin"))
        for k in [k for k in sys.modules if k == "models" or k.startswith("models.")]:
            del sys.modules[k]
        sys.modules.update(saved)


def test_frozen_stem_module_surface(golden_dir):
    """ObjDetectCNN: constructor of models/obj_detector.py:11-17, state_dict keys / shapes / order of the reference
    module (recorded in the golden file), weights_init, and no CPU path."""
    import numpy as np
    from videonavqa_b200._lib import VnqaError
    from videonavqa_b200.models import ObjDetectCNN
    g = np.load(os.path.join(golden_dir, "stem_small.npz"))
    m = ObjDetectCNN(nb_classes=27, num_filters=int(g["case"][3]), tail_hidden_dim=64, tail_dropout_p=0, logits=True,
                     pretrained_features=True)
    assert list(m.state_dict().keys()) == list(g["state_dict_keys"])
    for k in g.files:
        if k.startswith("p."):
            assert tuple(m.state_dict()[k[2:]].shape) == g[k].shape, k
    assert float(m.conv12.bias.abs().max()) == 0.0 and float(m.fc_tail1.bias.abs().max()) == 0.0   # weights_init
    big = ObjDetectCNN(nb_classes=27, num_filters=512, tail_hidden_dim=1024, tail_dropout_p=0, logits=True,
                       pretrained_features=True)                                                   # eval/utils.py:43-48
    assert big.fc_tail1.weight.shape == (1024, 512 * 30) and big.conv11.weight.shape == (512, 128, 3, 3)
    with pytest.raises(ValueError, match="multiple of 32"):
        ObjDetectCNN(27, num_filters=100)
    with pytest.raises(NotImplementedError):
        m.train()(torch.zeros(1, 128, 40, 52))
    with pytest.raises(VnqaError, match="no CPU fallback"):
        m.eval()(torch.zeros(1, 128, 40, 52))


def test_q8_embedding_padding_idx_matches_the_reference():
    """Quirk Q8: the attention model's embedding has NO padding_idx (film_attn_pt_stem.py:37: row 0 is an ordinary,
    randomly initialised row), the global-pooling and multi-hop models use padding_idx=0 (row 0 zero, no gradient)
    (film_global_pooling_pt_stem.py:34, time_multi_hop_pt_stem.py:30)."""
    from videonavqa_b200.models import (FiLMAttnPretrainedStem, FiLMGlobalPoolingPretrainedStem,
                                        TimeMultiHopFiLMPretrainedStem)
    kw = dict(num_input_channels=32, num_res_block_channels=32, hidden_size=32, vocab_size=20)
    a = FiLMAttnPretrainedStem(2, 16, 7, at_hidden_size=32, **kw)
    g = FiLMGlobalPoolingPretrainedStem(2, 16, 7, num_tail_channels=8, **kw)
    t = TimeMultiHopFiLMPretrainedStem(2, 16, 7, num_tail_channels=8, **kw)
    assert a.embed.padding_idx is None and float(a.embed.weight[0].abs().sum()) > 0
    for m in (g, t):
        assert m.embed.padding_idx == 0 and float(m.embed.weight[0].abs().sum()) == 0.0


def test_constructor_validates_kernel_constraints():
    from videonavqa_b200.models import FiLMAttnPretrainedStem
    with pytest.raises(ValueError, match="multiples of 32"):
        FiLMAttnPretrainedStem(2, 16, 7, num_input_channels=48, num_res_block_channels=50)
    with pytest.raises(ValueError, match="hidden_size"):
        FiLMAttnPretrainedStem(2, 16, 7, num_input_channels=32, num_res_block_channels=32, hidden_size=100)
    with pytest.raises(NotImplementedError):
        FiLMAttnPretrainedStem(2, 16, 7, num_input_channels=32, num_res_block_channels=32, q_encoder="bow")


def test_bench_reference_arm_json_contract():
    """`bench.py --impl reference` (the staged reference module when oracle/_ref is there, else the oracle port, on a
    bounded sample, GPUs hidden) prints ONE JSON line with the contract keys."""
    import json
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0", "--cpu-clips", "1", "--workload", "eval"], capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
              "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in line, k
    staged = os.path.exists(os.path.join(ROOT, "oracle", "_ref", "reference", "models_hot", "film_attn_pt_stem.py"))
    assert line["impl"] == "reference" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == ("reference" if staged else "port")
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["vs_baseline"] is None


def test_bench_reference_arm_stem_workload():
    """The same contract for the frozen-stem workload (`--workload stem`): one call of the reference's ObjDetectCNN on
    64 frame images per step, reported in clips/s at 35 frames per clip."""
    import json
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "stem",
                        "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and "frozen stem" in line["metric"] and line["unit"] == "clips/s"
    assert line["value"] > 0 and "64 frame images" in line["cpu_baseline"]["sample"]
    assert "models/obj_detector.py" in line["config"]["workload"]


def test_header_is_valid_c_and_struct_matches_ctypes(tmp_path):
    """include/vnqa_b200.h must compile as plain C (the ABI has no C++ / torch types) and vnqa_conv_args must have
    the size and field offsets the ctypes mirror assumes."""
    import ctypes
    import subprocess
    from videonavqa_b200._lib import ConvArgs
    src = tmp_path / "abi.c"
    fields = [f[0] for f in ConvArgs._fields_]
    body = "\n".join(f'  printf("%zu\\n", offsetof(vnqa_conv_args, {f}));' for f in fields)
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "vnqa_b200.h"\nint main(void) {\n'
                   '  printf("%zu\\n", sizeof(vnqa_conv_args));\n' + body + "\n  return 0;\n}\n")
    exe = tmp_path / "abi"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)],
                   check=True)
    out = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    assert out[0] == ctypes.sizeof(ConvArgs)
    assert out[1:] == [getattr(ConvArgs, f).offset for f in fields]


def test_bench_kernel_summary_helpers():
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    klass = {"conv_gemm.9tap.epi3.k1024.n1024": [62.0, 10, 5.496e12, [6.0] * 5 + [6.4] * 5],
             "hbm.film_bwd": [5.5, 10, 3.578e9, [0.55] * 10], "empty": [0.0, 0, 1.0, []]}
    k = bench.summarise_kernels(klass, steps=2, ms_per_step=140.0)
    assert abs(k["conv_gemm.9tap.epi3.k1024.n1024"]["tflops"] - 886.5) < 1 and "empty" not in k
    assert abs(k["hbm.film_bwd"]["gb_per_s"] - 6505.5) < 1 and k["hbm.film_bwd"]["launches_per_step"] == 5
    assert k["conv_gemm.9tap.epi3.k1024.n1024"]["ms_min_med_max"] == [6.0, 6.4, 6.4]
    r = bench.hbm_roofline(k, {"hbm_gbs": 6547.8})
    assert r["bound"] == "hbm" and 0.98 < r["frac"] < 1.0
    assert bench.hbm_roofline({}, {}) is None
