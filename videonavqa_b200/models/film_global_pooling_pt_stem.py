"""Drop-in for the reference's models/film_global_pooling_pt_stem.py (FiLMGlobalPoolingPretrainedStem, :8-238)."""
import torch
import torch.nn as nn

from .. import engine
from ._base import FilmModuleBase, check_channels, reference_init


class FiLMGlobalPoolingPretrainedStem(FilmModuleBase):
    kind = "gp"

    def __init__(self, batch_size, q_embedding_size, nb_classes, num_input_channels=512,
                 num_res_block_channels=512, num_tail_channels=16, num_res_blocks=1, hidden_size=128,
                 q_encoder='lstm', vocab_size=134):
        super().__init__()
        check_channels(num_input_channels, num_res_block_channels, hidden_size)
        assert q_encoder.lower() in ['lstm', 'bow'], "Invalid question encoder! ('lstm', 'bow')"
        if q_encoder.lower() != 'lstm':
            raise NotImplementedError("q_encoder='bow' is not supported by the B200 path; use 'lstm'")
        self.q_encoder = q_encoder
        self.nb_classes = nb_classes
        self.batch_size = batch_size
        self.q_embedding_size = q_embedding_size
        self.hidden_size = hidden_size
        C = num_res_block_channels
        # registration order of the reference (:34-56)
        self.embed = nn.Embedding(vocab_size, q_embedding_size, padding_idx=0)        # Q8
        self.relu = nn.ReLU(inplace=True)
        self.conv1x1_layers = []
        self.conv_init = nn.Conv2d(num_input_channels, C, kernel_size=3, padding=1)
        self.bn_init = nn.BatchNorm2d(C)
        self.film_layer = nn.ModuleList([nn.LSTM(q_embedding_size, hidden_size),
                                         nn.Linear(hidden_size, 2 * C * num_res_blocks),
                                         nn.ReLU(inplace=True)])
        convs = []
        for _ in range(num_res_blocks):
            convs.append(nn.Conv2d(C, C, kernel_size=3, padding=1))
            self.conv1x1_layers.append(nn.Conv2d(C, C, kernel_size=1))
        self.film_pipeline = nn.ModuleList(convs)
        self.c1x1_tail = nn.Conv2d(C, num_tail_channels, kernel_size=1)
        self.out_linear = nn.Linear(130 * num_tail_channels, nb_classes)
        for module in self.modules():
            reference_init(module)
        for module in self.conv1x1_layers:   # re-initialised in this model (:60-61), unlike the attention one
            reference_init(module)
        if torch.cuda.is_available():
            self.film_layer.cuda()
            for conv in self.conv1x1_layers:
                conv.cuda()
        self.init_hidden()

    def _forward_impl(self, P, v_input, q_input, v_lens, q_lens, save):
        v_host, q_lens32, perm, h, c, q_in = self._prepare(v_input, q_input, v_lens, q_lens)
        logits, ctx = engine.gp_forward(P, self._bn_state(), self._cache, v_input, q_in, v_host, q_lens32, h, c,
                                        self.training, save, prec=self.precision)
        self._finish(perm, h, c)
        ctx["q_in"], ctx["q_lens32"] = q_in, q_lens32
        return logits, ctx

    def _backward_impl(self, P, ctx, dlogits):
        if "y0" not in ctx:
            raise RuntimeError("backward through the B200 FiLM path needs train() mode (BatchNorm batch statistics)")
        return engine.gp_backward(P, self._cache, ctx, dlogits, ctx["q_in"], ctx["q_lens32"],
                                    grad_hook=getattr(self, "_grad_hook", None),
                                    grad_bucket=getattr(self, "_grad_bucket", None))
