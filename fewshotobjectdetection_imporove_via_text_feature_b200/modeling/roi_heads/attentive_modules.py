"""Text-fusion attention for the B200 ROI head.

API / state-dict mirror of defrcn/modeling/roi_heads/attentive_modules.py: `ScaledDotProductAttention` (:36-55),
`FFN` (:58-75), `SingleHeadSiameseAttention` (:78-177), `SematicProposalAttention` (:191-294).  Parameter names
are identical (`attention.w_q.weight`, `attention.linear1.0.weight`, `attention.ffn.norm3.weight`,
`key_projection.*`, ...), so reference checkpoints load unchanged.

Execution: in eval mode on CUDA the whole chain runs on the hand-written tcgen05/TMA bf16 GEMM plus the fused
attention / LayerNorm kernels (ops.text_fusion_forward); the text-side projections — constant at inference but
recomputed on every forward by the reference (:274-277) — are cached.  While fine-tuning (module.training with
autograd on) the differentiable torch expression below is used, still on the GPU.
"""
import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

from ... import ops
from ...utils.class_embedding import (SEMANTIC_DIM, create_normalized_orthogonal_tensor, get_class_embed,
                                      get_class_name)


class ScaledDotProductAttention(nn.Module):
    def __init__(self, temperature, dropout=0.0):
        super().__init__()
        self.temperature = temperature
        self.dropout = nn.Dropout(dropout)

    def forward(self, q, k, v):
        logits = torch.bmm(q, k.transpose(1, 2)) / self.temperature
        attn = self.dropout(F.softmax(logits, dim=2))
        return torch.bmm(attn, v), attn, F.log_softmax(logits, dim=2)


class FFN(nn.Module):
    def __init__(self, d_model, dropout=0.0, d_ffn=1024):
        super().__init__()
        self.d_model = d_model
        self.linear1 = nn.Linear(d_model, d_ffn)
        self.activation = F.relu
        self.dropout3 = nn.Dropout(dropout)
        self.linear2 = nn.Linear(d_ffn, d_model)
        self.dropout4 = nn.Dropout(dropout)
        self.norm3 = nn.LayerNorm(d_model)

    def forward(self, tgt):
        return self.norm3(tgt + self.dropout4(self.linear2(self.dropout3(self.activation(self.linear1(tgt))))))


class SingleHeadSiameseAttention(nn.Module):
    """Single head; q/k share nothing but the scale; a learned dummy key with a zero value is appended."""

    def __init__(self, d_model, dropout=0):
        super().__init__()
        self.n_head, self.d_model = 1, d_model
        self.w_q = nn.Linear(d_model, d_model, bias=False)
        self.w_k = nn.Linear(d_model, d_model, bias=False)
        self.w_v = nn.Linear(d_model, d_model, bias=False)
        self.attention = ScaledDotProductAttention(temperature=np.power(d_model, 0.5), dropout=dropout)
        std = np.sqrt(2.0 / (d_model + d_model))
        for lin in (self.w_q, self.w_k, self.w_v):
            nn.init.normal_(lin.weight, mean=0, std=std)
        self.dummy = nn.Parameter(torch.Tensor(1, d_model))
        nn.init.normal_(self.dummy)
        self.linear1 = nn.Sequential(nn.Linear(d_model, d_model // 2), nn.ReLU(inplace=True))
        self.linear2 = nn.Sequential(nn.Linear(d_model, d_model // 2), nn.ReLU(inplace=True))
        self.linear3 = nn.Linear(d_model * 2, d_model)
        self.ffn = FFN(d_model, dropout)
        self.dropout = nn.Dropout(dropout)

    def project_kv(self, k, v):
        """(B, Lk, d) keys/values -> projected, with the dummy key / zero value appended (:125-135)."""
        b = k.shape[0]
        kp = torch.cat([self.w_k(k), self.dummy.reshape(1, 1, -1).expand(b, -1, -1)], dim=1)
        vp = self.w_v(v)
        vp = torch.cat([vp, vp.new_zeros(b, 1, vp.shape[2])], dim=1)
        return kp, vp

    def forward(self, q, k, v):
        """Differentiable torch path: q (B,Lq,d), k/v (B,Lk,d) -> (out (B,Lq,d), attn (B,Lq,Lk+1))."""
        residual = q
        kp, vp = self.project_kv(k, v)
        out, attn, _ = self.attention(self.w_q(q), kp, vp)
        o1 = self.linear1(out * residual)
        o2 = self.linear2(residual - out)
        out = self.linear3(torch.cat([o1, o2, residual], dim=2))
        return self.ffn(out), attn


def _init_parameters(module, init_scale):
    for m in module.modules():
        if isinstance(m, (nn.Linear, nn.Embedding)):
            m.weight.data.normal_(mean=0.0, std=init_scale)
            if isinstance(m, nn.Linear) and m.bias is not None:
                m.bias.data.zero_()


class SematicProposalAttention(nn.Module):
    """ROI features attend over the class-name text embeddings (+ a background row)."""

    def __init__(self, input_size, cfg=None, is_multi=False, dropout=0, bg_generator=None):
        super().__init__()
        self.is_multi, self.dropout = is_multi, dropout
        self.addition_model = cfg.MODEL.ADDITION.NAME
        self.num_classes = cfg.MODEL.ROI_HEADS.NUM_CLASSES
        self.semantic_dim = SEMANTIC_DIM[self.addition_model]
        self.fixed_bg = False
        self.class_names = get_class_name(cfg)
        from ...config import b200_opt
        embed = get_class_embed(self.class_names, self.addition_model, include_bg=self.fixed_bg,
                                root=b200_opt(cfg, "EMBED_DIR", "datasets"))
        # like the reference these are plain tensors, not buffers: they are not part of the checkpoint
        self.embed = embed.float()
        self.class_embed = self.embed
        self.bg_feature = create_normalized_orthogonal_tensor(self.embed.mean(dim=0, keepdim=True), bg_generator)
        self.attention = SingleHeadSiameseAttention(input_size)
        self.query_projection = nn.Linear(input_size, self.semantic_dim)
        self.output_projection = nn.Linear(input_size, self.semantic_dim)
        self.key_projection = nn.Linear(self.semantic_dim, input_size)
        self.value_projection = nn.Linear(self.semantic_dim, input_size)
        with torch.no_grad():
            _init_parameters(self.attention, 0.02)
        self._plan = ops.TextFusionWeights()
        self.fold_query = True   # scores via the cached folded operand Kp.Wq (ops.text_fusion_forward)

    def _apply(self, fn, *a, **k):
        super()._apply(fn, *a, **k)
        self.embed, self.bg_feature = fn(self.embed), fn(self.bg_feature)
        self.class_embed = self.embed
        return self

    def forward_language_model(self):
        return {"text_feat": torch.cat([self.embed, self.bg_feature], dim=0)}

    def extra_weights(self):
        """Hook for subclasses/owners to have more matrices cast to bf16 with the same cache (name -> tensor)."""
        return {}

    def fused_weights(self, extra=None):
        named = {k: v for k, v in self.named_parameters()}
        for k, v in (extra or {}).items():
            named["extra." + k] = v
        return self._plan.refresh(named, (self.embed, self.bg_feature))

    def forward(self, visual_feat, extra=None):
        """Returns (attn (1,R,K+2), {'sim2stext' (R,d), 'text_feat' (K+1,D)}) like the reference; on the fused path
        the dict also carries bf16 copies ('sim2stext_bf16', 'x_bf16') for the predictor GEMMs."""
        output = self.forward_language_model()
        text_feat = output["text_feat"]
        if self.training and torch.is_grad_enabled():
            kt = F.relu(self.key_projection(text_feat))
            vt = F.relu(self.value_projection(text_feat))
            sim, attn = self.attention(q=visual_feat[None, :], k=kt[None, :], v=vt[None, :])
            output["sim2stext"] = F.relu(sim)[0]
            output["text_feat"] = text_feat.detach().clone()
            return attn, output
        if not visual_feat.is_cuda:
            raise RuntimeError("b200roi SematicProposalAttention: inference runs on CUDA only (no CPU fallback)")
        w = self.fused_weights(extra)
        z, zb, attn, xb = ops.text_fusion_forward(visual_feat, w, self.fold_query)
        output.update(sim2stext=z, sim2stext_bf16=zb, x_bf16=xb, text_feat=text_feat.detach().clone(), fused_w=w)
        return attn[None], output
