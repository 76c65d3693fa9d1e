"""Generate golden vectors under tests/golden/ from the UNMODIFIED reference modules.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python oracle/gen_golden.py

The reference ships no tests or fixtures; these files are the pins.  For each
case we build the reference nn.Module (models/film_attn_pt_stem.py, film_global_pooling_pt_stem.py,
time_multi_hop_pt_stem.py), dump ALL of its weights (state_dict + the CPU-branch ``film_layer``
list + the unregistered ``conv1x1_layers``), run an eval forward and a
train-mode forward+backward with CrossEntropyLoss(reduction='sum') on a seeded synthetic batch
and save logits, loss, every gradient, the BatchNorm running statistics and the carried
question-LSTM state.  The video tensor is regenerated from its numpy seed by
``oracle.film_oracle.synthetic_batch`` (a sha256 of it is stored to detect RNG drift).
"""
import hashlib
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.film_oracle import synthetic_batch  # noqa: E402

REF = "/root/reference"

CASES = {
    # name: (kind, ctor kwargs, batch kwargs, v_lens, q_lens)
    "attn_small": ("attn", dict(batch_size=3, q_embedding_size=16, nb_classes=7, num_input_channels=32,
                                num_res_block_channels=32, num_res_blocks=2, hidden_size=32,
                                at_hidden_size=32, max_num_frames=35, vocab_size=20),
                   dict(seed=11, Lq=12), [30, 17, 9], [12, 3, 7]),
    "attn_h128": ("attn", dict(batch_size=2, q_embedding_size=128, nb_classes=10, num_input_channels=32,
                               num_res_block_channels=32, num_res_blocks=1, hidden_size=128,
                               at_hidden_size=128, max_num_frames=35, vocab_size=134),
                  dict(seed=12, Lq=56), [35, 20], [9, 30]),
    "gp_small": ("gp", dict(batch_size=3, q_embedding_size=16, nb_classes=7, num_input_channels=32,
                            num_res_block_channels=32, num_tail_channels=8, num_res_blocks=2,
                            hidden_size=32, vocab_size=20),
                 dict(seed=13, Lq=12), [35, 35, 4], [5, 12, 1]),
    "tmh_small": ("tmh", dict(batch_size=3, q_embedding_size=16, nb_classes=7, num_input_channels=32,
                              num_res_block_channels=32, num_res_blocks=2, num_tail_channels=8,
                              hidden_size=32, vocab_size=20),
                  dict(seed=14, Lq=12), [28, 11, 11], [10, 12, 4]),
}


def build(kind, kw):
    sys.path.insert(0, REF)
    if kind == "attn":
        from models.film_attn_pt_stem import FiLMAttnPretrainedStem as M
    elif kind == "gp":
        from models.film_global_pooling_pt_stem import FiLMGlobalPoolingPretrainedStem as M
    else:
        from models.time_multi_hop_pt_stem import TimeMultiHopFiLMPretrainedStem as M
    return M(**kw)


def all_params(kind, m):
    """name -> Parameter for everything that takes part in forward."""
    out = dict(m.named_parameters())
    if kind != "tmh" and not isinstance(m.film_layer, torch.nn.ModuleList):
        # CPU branch: film_layer is a plain list (Q2) - export under the CUDA-branch keys
        for n, t in m.film_layer[0].named_parameters():
            out["film_layer.0." + n] = t
        for n, t in m.film_layer[1].named_parameters():
            out["film_layer.1." + n] = t
    for k, c in enumerate(m.conv1x1_layers):
        out[f"conv1x1_layers.{k}.weight"] = c.weight
        out[f"conv1x1_layers.{k}.bias"] = c.bias
    return out


def main():
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    for name, (kind, kw, bkw, v_lens, q_lens) in CASES.items():
        torch.manual_seed(100 + len(name))
        m = build(kind, kw)
        B = kw["batch_size"]
        v, q, _, _, labels = synthetic_batch(B, c_in=kw["num_input_channels"], vocab=kw["vocab_size"],
                                             classes=kw["nb_classes"], **bkw)
        v_lens_t = torch.tensor(v_lens, dtype=torch.int64)
        q_lens_t = torch.tensor(q_lens, dtype=torch.int64)
        q = q.clone()
        q[torch.arange(q.shape[1])[None, :] >= q_lens_t[:, None]] = 0
        # make sure every question has non-pad tokens up to its (overridden) length
        rng = np.random.default_rng(bkw["seed"] + 1000)
        fill = torch.from_numpy(rng.integers(1, kw["vocab_size"], size=tuple(q.shape)).astype(np.int64))
        q = torch.where((torch.arange(q.shape[1])[None, :] < q_lens_t[:, None]) & (q == 0), fill, q)

        params = all_params(kind, m)
        out = {"p." + k: t.detach().numpy().copy() for k, t in params.items()}
        out["q_input"] = q.numpy()
        out["v_lens"] = v_lens_t.numpy()
        out["q_lens"] = q_lens_t.numpy()
        out["labels"] = labels.numpy()
        out["v_sha256"] = np.frombuffer(hashlib.sha256(v.numpy().tobytes()).digest(), dtype=np.uint8)

        # eval forward from fresh BN statistics / fresh hidden state
        m.eval()
        m.init_hidden()
        with torch.no_grad():
            out["logits_eval"] = m(v, q, v_lens_t, q_lens_t).numpy().copy()
        out["film_h_eval"] = m.film_hidden[0].detach().numpy().copy()[0]
        out["film_c_eval"] = m.film_hidden[1].detach().numpy().copy()[0]

        # train forward + backward
        m.train()
        m.init_hidden()
        logits = m(v, q, v_lens_t, q_lens_t)
        loss = F.cross_entropy(logits, labels, reduction="sum")
        loss.backward()
        out["logits_train"] = logits.detach().numpy().copy()
        out["loss_train"] = np.array(loss.item(), dtype=np.float64)
        for k, t in params.items():
            if t.grad is not None:
                out["g." + k] = t.grad.numpy().copy()
        out["bn_running_mean"] = m.bn_init.running_mean.numpy().copy()
        out["bn_running_var"] = m.bn_init.running_var.numpy().copy()
        out["bn_num_batches_tracked"] = np.array(int(m.bn_init.num_batches_tracked))
        out["film_h_train"] = m.film_hidden[0].detach().numpy().copy()[0]
        out["film_c_train"] = m.film_hidden[1].detach().numpy().copy()[0]
        out["state_dict_keys"] = np.array(list(m.state_dict().keys()))

        path = os.path.join(ROOT, "tests", "golden", name + ".npz")
        np.savez_compressed(path, **out)
        print(name, "->", path, "%.2f MB" % (os.path.getsize(path) / 1e6),
              "loss", float(out["loss_train"]))


if __name__ == "__main__":
    main()
