"""Teacher ("language-vision") attention variants — SURVEY.md §8a row A7.

API mirror of defrcn/modeling/roi_heads/attentive_modules.py: `LV_attention` (:297-437), `LV_attention_VKV`
(:440-487), `LV_attention_textDomination` (:490-634), `LV_attention_textDomination_VKV` (:637-687); same parameter
names (`attention.*`, `proj_k`, `proj2`, `proj_visual`, `proj_value`, `w_bg`).

These modules are *teachers*: they consume ground-truth labels (the text key of ROI i is the embedding of its GT
class), so they only run while training or in the reference's test-with-GT mode, and their key set is per-ROI: the
attention matrix is (R, R+1) over the ROIs of the local batch.  With autograd on (the teacher itself is being trained)
they run as differentiable torch expressions on the GPU, sharing `SingleHeadSiameseAttention` with the student path.
Without autograd — the frozen teacher of student training, test-with-GT — all four variants take a fused path
(`ops.teacher_attention_forward`, `ops.text_domination_forward`): ROIs of one class share their key, so the dense
attention collapses to a (K+2)-key attention (logit + log n_c, per-class mean values) and runs on the tcgen05 GEMM /
fused kernels; the 300-d `textDomination` variants keep their operands in 304-wide buffers (TMA row pitch), the products
still contract over exactly 300 columns.

Where the checked-in reference is broken (`LV_attention_VKV.forward` calls
`forward_language_model(visual_feat, text)` against a one-argument signature and concatenates a 2-d with a 3-d
tensor) the *intended* computation is implemented: Q = V = relu(proj([x ‖ t])), K = relu(t).
"""
import torch
import torch.nn.functional as F
from torch import nn

from ... import ops
from ...utils.class_embedding import get_class_embed, get_class_name
from .attentive_modules import SingleHeadSiameseAttention, _init_parameters


class LV_attention(nn.Module):
    text_dim = 300

    def __init__(self, input_size, cfg=None, is_multi=False, output_size=0, dropout=0, class_embed=None):
        super().__init__()
        self.is_multi, self.dropout = is_multi, dropout
        self.output_size = output_size if output_size else input_size
        self.__init_language_model__(cfg, class_embed)
        self.__init_attention_layer__(input_size)

    def __init_language_model__(self, cfg, class_embed=None):
        self.classes = get_class_name(cfg)
        from ...config import b200_opt
        # the reference sums GloVe-6B word vectors of the (possibly two-word) class name; a (K,300) table is the input
        self.embed = (class_embed if class_embed is not None else
                      get_class_embed(self.classes, "glove", root=b200_opt(cfg, "EMBED_DIR", "datasets"))).float()
        self.class_id = torch.arange(len(self.classes) + 1)
        self.w_bg_init = torch.randn(1, self.text_dim)
        self.w_bg = nn.Parameter(self.w_bg_init.clone(), requires_grad=True)

    def __init_attention_layer__(self, input_size):
        self.attention = SingleHeadSiameseAttention(input_size)
        self.proj_k = nn.Linear(input_size * 2, input_size)
        self.proj2 = nn.Linear(self.text_dim, input_size)
        with torch.no_grad():
            _init_parameters(self.attention, 0.02)

    def _apply(self, fn, *a, **k):
        nn.Module._apply(self, fn, *a, **k)          # explicit base: LV_attention_textDomination borrows this method
        self.embed = fn(self.embed)
        return self

    def class_table(self):
        """(K+1, d) projected class embeddings, background row last."""
        return self.proj2(torch.cat([self.embed, self.w_bg], dim=0))

    def forward_language_model(self, label):
        # one_hot(label) @ table == table[label]
        return {}, {"text_feat": self.class_table()[label]}

    _vkv = False

    def _frozen_forward(self, visual_feat, text):
        """Forward without autograd (frozen teacher while the student trains, test-with-GT): the class-collapsed fused
        path, ops.teacher_attention_forward."""
        if not hasattr(self, "_plan"):
            self._plan = ops.TeacherFusionWeights()
        w = self._plan.refresh({k: v for k, v in self.named_parameters()}, (self.embed,))
        z, zb = ops.teacher_attention_forward(visual_feat, text, w, self._vkv)
        t = w["table"][text]
        return {}, {"text_feat": t[None, :] if self._vkv else t, "sim2stext": z[None, :], "sim2stext_bf16": zb}

    def forward(self, visual_feat, text, num_preds_per_image=None):
        if visual_feat.is_cuda and not torch.is_grad_enabled():
            return self._frozen_forward(visual_feat, text)
        loss, output = self.forward_language_model(text)
        t = output["text_feat"]
        value = F.relu(self.proj_k(torch.cat([visual_feat, t], dim=-1)))
        sim = self.attention(q=visual_feat[None, :], k=F.relu(t)[None, :], v=value[None, :])[0]
        output["sim2stext"] = F.relu(sim)
        return loss, output


class LV_attention_VKV(LV_attention):
    """Query and value are both the fused [visual ‖ text] projection; keys are the GT-class text features."""
    _vkv = True

    def forward(self, visual_feat, text, num_preds_per_image=None):
        if visual_feat.is_cuda and not torch.is_grad_enabled():
            return self._frozen_forward(visual_feat, text)
        loss, output = self.forward_language_model(text)
        t = output["text_feat"][None, :]
        value = F.relu(self.proj_k(torch.cat([visual_feat[None, :], t], dim=2)))
        output["text_feat"] = t
        output["sim2stext"] = F.relu(self.attention(q=value, k=F.relu(t), v=value)[0])
        return loss, output


class LV_attention_textDomination(nn.Module):
    """Attention runs in the 300-d text space: visual features are projected down, the result projected back up."""
    text_dim = 300

    def __init__(self, input_size, cfg=None, is_multi=False, output_size=0, dropout=0, class_embed=None,
                 student_training=False):
        super().__init__()
        self.is_multi, self.dropout = is_multi, dropout
        self.output_size = output_size if output_size else input_size
        self.student_training = student_training
        LV_attention.__init_language_model__(self, cfg, class_embed)
        self.attention = SingleHeadSiameseAttention(self.text_dim)
        self.proj_k = nn.Linear(input_size * 2, input_size)
        self.proj2 = nn.Linear(self.text_dim, input_size)
        self.proj_visual = nn.Linear(input_size, self.text_dim)
        self.proj_value = nn.Linear(self.text_dim * 2, self.text_dim)
        with torch.no_grad():
            _init_parameters(self.attention, 0.02)
        if student_training:
            mlp = lambda: nn.Sequential(nn.Linear(input_size, input_size), nn.ReLU(), nn.Linear(input_size, input_size), nn.ReLU())
            self.mlp_adapter, self.mlp_adapter2 = mlp(), mlp()

    _apply = LV_attention._apply

    def forward_language_model(self, visual_feat, label):
        table = torch.cat([self.embed, self.w_bg], dim=0)
        return {}, {"text_feat": table[label][None, :]}

    def _qkv(self, visual_feat, text):
        v300 = self.proj_visual(visual_feat)
        loss, output = self.forward_language_model(v300, text)
        t = output["text_feat"]
        value = F.relu(self.proj_value(torch.cat([v300[None, :], t], dim=2)))
        return loss, output, v300[None, :], F.relu(t), value

    _vkv = False

    def _frozen_forward(self, visual_feat, text):
        """Forward without autograd (frozen teacher while the student trains, test-with-GT): the class-collapsed path on the
        tensor-core GEMM, ops.text_domination_forward."""
        if not hasattr(self, "_plan"):
            self._plan = ops.TextDominationWeights()
        named = {k: v for k, v in self.named_parameters() if not k.startswith(("mlp_adapter", "proj_k"))}
        w = self._plan.refresh(named, (self.embed,))
        out = ops.text_domination_forward(visual_feat, text, w, self._vkv)
        return {}, {"text_feat": w["table"][text][None, :], "sim2stext": out[None, :]}

    def forward(self, visual_feat, text, num_preds_per_image=None):
        if visual_feat.is_cuda and not torch.is_grad_enabled():
            return self._frozen_forward(visual_feat, text)
        loss, output, q, k, v = self._qkv(visual_feat, text)
        output["sim2stext"] = self.proj2(F.relu(self.attention(q=q, k=k, v=v)[0]))
        return loss, output


class LV_attention_textDomination_VKV(LV_attention_textDomination):
    _vkv = True

    def forward(self, visual_feat, text, num_preds_per_image=None):
        if visual_feat.is_cuda and not torch.is_grad_enabled():
            return self._frozen_forward(visual_feat, text)
        loss, output, q, k, v = self._qkv(visual_feat, text)
        output["sim2stext"] = self.proj2(F.relu(self.attention(q=v, k=k, v=v)[0]))
        return loss, output
