"""Drop-in for the reference's models/film_attn_pt_stem.py (class FiLMAttnPretrainedStem, :9-301).

Same constructor arguments (:14-25), ``forward(v_input, q_input, v_lens, q_lens)`` (:188), ``init_hidden()`` (:133)
and ``state_dict`` keys/shapes/order as the reference's CUDA branch; the arithmetic runs in libvnqa_b200.so.
"""
import torch
import torch.nn as nn

from .. import engine
from ._base import FilmModuleBase, check_channels, reference_init


class FiLMAttnPretrainedStem(FilmModuleBase):
    kind = "attn"

    def __init__(self, batch_size, q_embedding_size, nb_classes, num_input_channels=512,
                 num_res_block_channels=512, num_res_blocks=1, hidden_size=128, at_hidden_size=128,
                 max_num_frames=35, q_encoder='lstm', vocab_size=134):
        super().__init__()
        check_channels(num_input_channels, num_res_block_channels, hidden_size)
        if at_hidden_size % 32 or not 32 <= at_hidden_size <= 256:
            raise ValueError(f"videonavqa_b200 supports at_hidden_size in 32..256, multiple of 32 (got {at_hidden_size})")
        assert q_encoder.lower() in ['lstm', 'bow'], "Invalid question encoder! ('lstm', 'bow')"
        if q_encoder.lower() != 'lstm':
            # the reference's 'bow' branch discards its own averaging and hard-codes
            # torch.cuda.FloatTensor; only 'lstm' is ever launched (eval/q_and_v_eval.py:35).
            raise NotImplementedError("q_encoder='bow' is not supported by the B200 path; use 'lstm'")
        self.q_encoder = q_encoder
        self.nb_classes = nb_classes
        self.batch_size = batch_size
        self.q_embedding_size = q_embedding_size
        self.at_hidden_size = at_hidden_size
        self.hidden_size = hidden_size
        self.max_num_frames = max_num_frames
        C = num_res_block_channels
        # registration order mirrors the reference (:37-63): it is the state_dict / optimizer-state order
        self.embed = nn.Embedding(vocab_size, q_embedding_size)                      # no padding_idx (Q8)
        self.relu = nn.ReLU(inplace=True)
        self.conv_init = nn.Conv2d(num_input_channels, C, kernel_size=3, padding=1)
        self.bn_init = nn.BatchNorm2d(C)
        self.conv1x1_layers = []            # plain list on purpose (Q1): not in state_dict / parameters()
        self.film_layer = nn.ModuleList([nn.LSTM(q_embedding_size, hidden_size),
                                         nn.Linear(hidden_size, 2 * C * num_res_blocks),
                                         nn.ReLU(inplace=True)])                      # CUDA-branch layout (Q2)
        convs = []
        for _ in range(num_res_blocks):
            convs.append(nn.Conv2d(C, C, kernel_size=3, padding=1))
            self.conv1x1_layers.append(nn.Conv2d(C, C, kernel_size=1))
        self.film_pipeline = nn.ModuleList(convs)
        self.fc_embed_attn = nn.Linear(130 * C, at_hidden_size)
        self.fc_attn_1 = nn.Linear(at_hidden_size, 1)
        self.fc_hidden_attn = nn.Linear(at_hidden_size, 1)
        self.lstm_attn = nn.LSTMCell(at_hidden_size, at_hidden_size)
        self.out_linear = nn.Linear(max_num_frames * at_hidden_size, nb_classes)
        for module in self.modules():
            reference_init(module)          # conv1x1_layers keep PyTorch's default init here (:67-68)
        if torch.cuda.is_available():
            self.film_layer.cuda()
            for conv in self.conv1x1_layers:
                conv.cuda()
        self.init_hidden()

    def _forward_impl(self, P, v_input, q_input, v_lens, q_lens, save):
        if v_input.shape[-1] != self.max_num_frames:
            raise RuntimeError(f"v_input has {v_input.shape[-1]} frames, the module is locked to "
                               f"max_num_frames={self.max_num_frames} (out_linear width, SURVEY Q7)")
        v_host, q_lens32, perm, h, c, q_in = self._prepare(v_input, q_input, v_lens, q_lens)
        logits, ctx = engine.attn_forward(P, self._bn_state(), self._cache, v_input, q_in, v_host, q_lens32, h, c,
                                          self.max_num_frames, self.training, save, prec=self.precision)
        self._finish(perm, h, c)
        ctx["q_in"], ctx["q_lens32"] = q_in, q_lens32
        return logits, ctx

    def _backward_impl(self, P, ctx, dlogits):
        if "y0" not in ctx:
            raise RuntimeError("backward through the B200 FiLM path needs train() mode (BatchNorm batch statistics)")
        return engine.attn_backward(P, self._cache, ctx, dlogits, ctx["q_in"], ctx["q_lens32"],
                                    grad_hook=getattr(self, "_grad_hook", None),
                                    grad_bucket=getattr(self, "_grad_bucket", None))
