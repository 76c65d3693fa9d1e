"""Drop-in for the reference's models/time_multi_hop_pt_stem.py (TimeMultiHopFiLMPretrainedStem, :8-250)."""
import torch
import torch.nn as nn

from .. import engine
from ._base import FilmModuleBase, check_channels, reference_init


class TimeMultiHopFiLMPretrainedStem(FilmModuleBase):
    kind = "tmh"

    def __init__(self, batch_size, q_embedding_size, nb_classes, num_input_channels=512,
                 num_res_block_channels=512, num_res_blocks=1, num_tail_channels=32, hidden_size=128,
                 vocab_size=134):
        super().__init__()
        check_channels(num_input_channels, num_res_block_channels, hidden_size)
        self.nb_classes = nb_classes
        self.batch_size = batch_size
        self.q_embedding_size = q_embedding_size
        self.hidden_size = hidden_size
        C = num_res_block_channels
        total_out_size = 2 * C * num_res_blocks
        # registration order of the reference (:30-60)
        self.embed = nn.Embedding(vocab_size, q_embedding_size, padding_idx=0)
        self.relu = nn.ReLU(inplace=True)
        self.conv_init = nn.Conv2d(num_input_channels, C, kernel_size=3, padding=1)
        self.bn_init = nn.BatchNorm2d(C)
        self.conv1x1_layers = []
        self.q_encoder = nn.LSTM(q_embedding_size, hidden_size)
        self.encoder_norm = nn.LayerNorm(hidden_size)
        self.h = None
        self.fc_hidden_attn = nn.Linear(hidden_size, 1)
        self.fc_attn_out = nn.Linear(hidden_size, total_out_size)
        self.decoder_norm = nn.LayerNorm(total_out_size)
        convs = []
        for _ in range(num_res_blocks):
            convs.append(nn.Conv2d(C, C, kernel_size=3, padding=1))
            self.conv1x1_layers.append(nn.Conv2d(C, C, kernel_size=1))
        self.film_pipeline = nn.ModuleList(convs)
        self.c1x1_tail = nn.Conv2d(C, num_tail_channels, kernel_size=1)
        self.out_linear = nn.Linear(130 * num_tail_channels, nb_classes)
        for module in self.modules():
            reference_init(module)
        for module in self.conv1x1_layers:
            reference_init(module)
        if torch.cuda.is_available():
            for conv in self.conv1x1_layers:
                conv.cuda()
        self.init_hidden()

    def init_hidden(self):
        super().init_hidden()
        self.h = None                      # models/time_multi_hop_pt_stem.py:116

    def _forward_impl(self, P, v_input, q_input, v_lens, q_lens, save):
        v_host, q_lens32, perm, h, c, q_in = self._prepare(v_input, q_input, v_lens, q_lens)
        q_host = self._last_q_host
        logits, ctx = engine.tmh_forward(P, self._bn_state(), self._cache, v_input, q_in, v_host, q_host, q_lens32, h, c,
                                         self.training, save, prec=self.precision)
        self._finish(perm, h, c)
        ctx["q_in"], ctx["q_lens32"] = q_in, q_lens32
        return logits, ctx

    def _backward_impl(self, P, ctx, dlogits):
        if "y0" not in ctx:
            raise RuntimeError("backward through the B200 FiLM path needs train() mode (BatchNorm batch statistics)")
        return engine.tmh_backward(P, self._cache, ctx, dlogits, ctx["q_in"], ctx["q_lens32"],
                                   grad_hook=getattr(self, "_grad_hook", None),
                                    grad_bucket=getattr(self, "_grad_bucket", None))
