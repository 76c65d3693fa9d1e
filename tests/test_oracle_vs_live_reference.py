"""Extra pin of the oracle at the SHIPPED hyper-parameters (eval.sh: C = 1024, 5 / 4 / 3 blocks) against the live,
unmodified reference modules.  Runs only where /root/reference exists (the build container); the GPU box relies on the
committed golden vectors instead."""
import os
import sys

import pytest
import torch

from oracle import film_oracle as O
from oracle.gen_golden import all_params, build

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "models")), reason="reference tree not present")

CONFIGS = {
    "attn": dict(batch_size=2, q_embedding_size=128, nb_classes=70, num_input_channels=512,
                 num_res_block_channels=1024, num_res_blocks=5, hidden_size=128, at_hidden_size=128,
                 max_num_frames=35, vocab_size=134),
    "gp": dict(batch_size=2, q_embedding_size=128, nb_classes=70, num_input_channels=512, num_res_block_channels=1024,
               num_tail_channels=32, num_res_blocks=4, hidden_size=128, vocab_size=134),
    "tmh": dict(batch_size=2, q_embedding_size=128, nb_classes=70, num_input_channels=512,
                num_res_block_channels=1024, num_res_blocks=3, num_tail_channels=64, hidden_size=128, vocab_size=134),
}


@pytest.mark.parametrize("kind", ["attn", "gp", "tmh"])
def test_oracle_matches_live_reference_at_eval_sh_size(kind):
    torch.manual_seed(7)
    kw = CONFIGS[kind]
    m = build(kind, kw)
    v, q, _, q_lens, labels = O.synthetic_batch(2, seed=31)
    v_lens = torch.tensor([3, 2])
    params = {k: t.detach().clone() for k, t in all_params(kind, m).items()}
    m.train()
    m.init_hidden()
    logits = m(v, q, v_lens, q_lens)
    loss = torch.nn.functional.cross_entropy(logits, labels, reduction="sum")
    loss.backward()
    ref_grads = {k: t.grad for k, t in all_params(kind, m).items() if t.grad is not None}
    lo, lo_loss, g, _ = O.loss_and_grads(kind, params, O.new_bn_state(1024), O.new_film_state(2, 128), v, q, v_lens,
                                         q_lens, labels, training=True)
    scale = float(logits.detach().abs().max())
    assert float((lo - logits.detach()).abs().max()) <= 2e-5 * scale
    assert abs(float(lo_loss) - float(loss)) <= 1e-4 * abs(float(loss))
    for k in ("conv_init.weight", "film_pipeline.0.weight", "conv1x1_layers.0.weight", "embed.weight",
              "out_linear.weight"):
        a, b = g[k].double().flatten(), ref_grads[k].double().flatten()
        assert float((a - b).norm() / b.norm().clamp_min(1e-30)) <= 2e-3, k      # fp32 summation-order noise through BatchNorm over 5 frame images
    if "/root/reference" in sys.path:
        sys.path.remove("/root/reference")
