"""Argmax-agreement report (north-star gate: >= 99.9 % over >= 4096 synthetic clips) - run on the B200 box:

    python tests/tools/argmax_agreement.py [--clips 4096] [--batch 64] [--blocks 5] [--channels 1024]

The checker is the CPU oracle's functional code (oracle/film_oracle.py) executed on CUDA tensors in strict fp32
(TF32 disabled in cuDNN/cuBLAS) - test infrastructure only; the product path under test is
videonavqa_b200.models.FiLMAttnPretrainedStem.  Prints the normalised max logit error and the fraction of clips whose
argmax answer matches, plus the margin statistics of the clips that disagree (near-ties of random-init logits).
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from oracle import film_oracle as O  # noqa: E402
from videonavqa_b200.models import FiLMAttnPretrainedStem  # noqa: E402
from videonavqa_b200.synth import synthetic_batch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--clips", type=int, default=4096)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--blocks", type=int, default=5)
    ap.add_argument("--channels", type=int, default=1024)
    ap.add_argument("--ragged", action="store_true")
    ap.add_argument("--operands", choices=["fp16", "tf32"], default=None)
    a = ap.parse_args()
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    dev = torch.device("cuda")
    torch.manual_seed(1)
    m = FiLMAttnPretrainedStem(batch_size=a.batch, q_embedding_size=128, nb_classes=70, num_input_channels=512,
                               num_res_block_channels=a.channels, num_res_blocks=a.blocks, hidden_size=128,
                               at_hidden_size=128, max_num_frames=35, vocab_size=134).to(dev).eval()
    if a.operands:
        m.set_operands(a.operands)
    params = {k: p.detach() for k, p in m.named_parameters()}
    for k, conv in enumerate(m.conv1x1_layers):
        params[f"conv1x1_layers.{k}.weight"] = conv.weight.detach()
        params[f"conv1x1_layers.{k}.bias"] = conv.bias.detach()
    bn = {"running_mean": m.bn_init.running_mean.clone(), "running_var": m.bn_init.running_var.clone(),
          "num_batches_tracked": m.bn_init.num_batches_tracked.clone()}
    agree = total = 0
    worst = 0.0
    margins = []
    for i in range((a.clips + a.batch - 1) // a.batch):
        v, q, v_lens, q_lens, _ = synthetic_batch(a.batch, seed=1000 + i, ragged=a.ragged)
        v, q, v_lens, q_lens = v.to(dev), q.to(dev), v_lens.to(dev), q_lens.to(dev)
        with torch.no_grad():
            m.init_hidden()
            lo = m(v, q, v_lens, q_lens)
            st = (torch.zeros(a.batch, 128, device=dev), torch.zeros(a.batch, 128, device=dev))
            ref, _ = O.film_attn_forward(params, bn, st, v, q, v_lens, q_lens, training=False)
        worst = max(worst, float((lo - ref).abs().max() / ref.abs().max()))
        same = lo.argmax(1) == ref.argmax(1)
        agree += int(same.sum()); total += a.batch
        top2 = ref.topk(2, dim=1).values
        margins += ((top2[:, 0] - top2[:, 1]) / ref.abs().max())[~same].tolist()
    print(f"operands {m.precision.name}  clips {total}  argmax agreement {100.0 * agree / total:.3f} %  normalised max logit error {worst:.3e}")
    if margins:
        print(f"disagreeing clips: {len(margins)}; their reference top-2 margins (normalised): max {max(margins):.2e}")


if __name__ == "__main__":
    main()
