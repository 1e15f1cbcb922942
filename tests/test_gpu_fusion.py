"""Text-fusion chain kernels (tcgen05 GEMM, attention core, LayerNorm) vs fp32 torch / the oracle.
Tolerance: GEMM against fp32 math on the same bf16-rounded operands 1e-3 (accumulation order only);
whole chain against the fp32 oracle 2e-2 relative (north_star bf16 bar)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from oracle import oracle as O


def rel_err(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-12))


@pytest.mark.parametrize("M,N,K", [(512, 2048, 2048), (300, 21, 2048), (129, 80, 2048), (1000, 2048, 4096),
                                   (128, 64, 72), (1, 32, 64), (77, 1024, 16), (640, 512, 2048), (4096, 2048, 1024)])
@pytest.mark.parametrize("relu,use_bias", [(False, True), (True, False)])
def test_gemm_bf16(M, N, K, relu, use_bias):
    from fewshotobjectdetection_imporove_via_text_feature_b200 import ops
    gen = torch.Generator().manual_seed(M + N + K)
    a = (torch.randn(M, K, generator=gen) * 0.5).to(torch.bfloat16)
    b = (torch.randn(N, K, generator=gen) * 0.05).to(torch.bfloat16)
    bias = torch.randn(N, generator=gen) if use_bias else None
    ref = a.double() @ b.double().t()
    if bias is not None:
        ref = ref + bias.double()
    if relu:
        ref = ref.clamp_min(0)
    if bias is not None:      # biases are views into the optimizer's flat buffer: any 4-byte alignment must work
        bias_d = torch.zeros(N + 3, device="cuda")[3:]
        bias_d.copy_(bias)
        assert bias_d.data_ptr() % 16 != 0
        out = ops.gemm_bf16(a.cuda(), b.cuda(), bias_d, relu=relu)
        torch.testing.assert_close(out.cpu().double(), ref, rtol=1e-3, atol=1e-3)
    out = ops.gemm_bf16(a.cuda(), b.cuda(), None if bias is None else bias.cuda(), relu=relu)
    assert out.dtype == torch.float32 and out.shape == (M, N)
    torch.testing.assert_close(out.cpu().double(), ref, rtol=1e-3, atol=1e-3)
    outb = ops.gemm_bf16(a.cuda(), b.cuda(), None if bias is None else bias.cuda(), relu=relu, out_dtype=torch.bfloat16)
    torch.testing.assert_close(outb.cpu().double(), ref, rtol=1e-2, atol=1e-2)


@pytest.mark.parametrize("ctas", [1, 3, 5])
@pytest.mark.parametrize("M,N,K", [(1000, 2048, 512), (700, 80, 2048), (515, 1000, 136)])
def test_gemm_persistent_many_tiles_per_cta(M, N, K, ctas):
    """Persistent scheduler under pressure: few CTAs walk many tiles each, so both TMEM accumulators and every smem
    stage wrap their mbarrier phases several times; the backward epilogue options (ReLU mask, fp32 accumulate, second
    bf16 copy) ride along.  Results must not depend on the number of CTAs."""
    from fewshotobjectdetection_imporove_via_text_feature_b200 import _lib, train_ops
    gen = torch.Generator().manual_seed(M + K)
    a = (torch.randn(M, K, generator=gen) * 0.5).to(torch.bfloat16).cuda()
    b = (torch.randn(N, K, generator=gen) * 0.05).to(torch.bfloat16).cuda()
    mask = torch.randn(M, (N + 7) // 8 * 8, generator=gen).to(torch.bfloat16).cuda()[:, :N]
    acc0 = torch.randn(M, N, generator=gen).cuda()
    ref = a.double() @ b.double().t()
    ref = torch.where(mask.double() > 0, ref, torch.zeros((), dtype=torch.float64, device="cuda")) + acc0.double()

    def run():
        out, d2 = acc0.clone(), torch.empty(M, (N + 7) // 8 * 8, dtype=torch.bfloat16, device="cuda")[:, :N]
        train_ops.gemm_ex(a, b, out=out, out2=d2, accumulate=True, mask=mask)
        return out, d2
    base, base2 = run()
    torch.testing.assert_close(base.double(), ref, rtol=1e-3, atol=1e-3)
    _lib.set_option("gemm_ctas", ctas)
    try:
        out, d2 = run()
    finally:
        _lib.set_option("gemm_ctas", 0)
    assert torch.equal(out, base) and torch.equal(d2, base2)
    assert torch.equal(d2, base.to(torch.bfloat16))
    _lib.set_option("gemm_generic_epilogue", 1)          # element-wise epilogue == vector epilogue, bit for bit
    try:
        out, d2 = run()
    finally:
        _lib.set_option("gemm_generic_epilogue", 0)
    assert torch.equal(out, base) and torch.equal(d2, base2)


def test_gemm_strided_and_second_output():
    from fewshotobjectdetection_imporove_via_text_feature_b200 import ops
    gen = torch.Generator().manual_seed(0)
    M, N, K = 200, 1024, 2048
    big = (torch.randn(M, 2 * K, generator=gen) * 0.5).to(torch.bfloat16).cuda()
    a = big[:, K:]                                        # row stride 2K, like the concat buffer
    b = (torch.randn(N, K, generator=gen) * 0.05).to(torch.bfloat16).cuda()
    dst = torch.zeros(M, 4 * N, dtype=torch.bfloat16, device="cuda")
    d2 = torch.zeros(M, N, dtype=torch.bfloat16, device="cuda")
    out = ops.gemm_bf16(a, b, out=dst[:, N:2 * N])
    ref = (a.double() @ b.double().t()).cpu()
    torch.testing.assert_close(dst[:, N:2 * N].cpu().double(), ref, rtol=1e-2, atol=1e-2)
    assert float(dst[:, :N].abs().max()) == 0 and float(dst[:, 2 * N:].abs().max()) == 0   # no stray writes
    f32 = ops.gemm_bf16(a, b, out2=d2)
    torch.testing.assert_close(f32.cpu().double(), ref, rtol=1e-3, atol=1e-3)
    torch.testing.assert_close(d2.float(), f32.to(torch.bfloat16).float(), rtol=0, atol=0)


@pytest.mark.parametrize("R,d,L", [(512, 2048, 22), (37, 2048, 82), (600, 64, 7), (1500, 2048, 22)])
def test_text_attention(R, d, L):
    from fewshotobjectdetection_imporove_via_text_feature_b200 import ops
    gen = torch.Generator().manual_seed(R)
    q = (torch.randn(R, d, generator=gen)).to(torch.bfloat16)
    x = torch.relu(torch.randn(R, d, generator=gen))
    kp, vp = torch.randn(L, d, generator=gen), torch.randn(L, d, generator=gen)
    vp[-1] = 0
    s = (q.float() @ kp.t()) / np.sqrt(d)
    attn = F.softmax(s, dim=1)
    o = attn @ vp
    p1 = torch.empty(R, d, dtype=torch.bfloat16, device="cuda")
    p2 = torch.empty(R, d, dtype=torch.bfloat16, device="cuda")
    got = ops.text_attention(q.cuda(), x.cuda(), kp.cuda(), vp.cuda(), p1, p2)
    torch.testing.assert_close(got.cpu(), attn, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(p1.float().cpu(), o * x, rtol=1e-2, atol=1e-2)
    torch.testing.assert_close(p2.float().cpu(), x - o, rtol=1e-2, atol=1e-2)
    # scores-in mode (folded query GEMM feeds the kernel): same softmax / AV / gate outputs
    p1b, p2b = torch.empty_like(p1), torch.empty_like(p2)
    got2 = ops.text_attention(None, x.cuda(), None, vp.cuda(), p1b, p2b, scores=s.cuda().contiguous())
    torch.testing.assert_close(got2.cpu(), attn, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(p1b.float().cpu(), o * x, rtol=1e-2, atol=1e-2)
    torch.testing.assert_close(p2b.float().cpu(), x - o, rtol=1e-2, atol=1e-2)


def test_residual_layernorm():
    from fewshotobjectdetection_imporove_via_text_feature_b200 import ops
    gen = torch.Generator().manual_seed(1)
    R, d = 333, 2048
    y, y2 = torch.randn(R, d, generator=gen), torch.randn(R, d, generator=gen) * 3
    g, b = torch.randn(d, generator=gen), torch.randn(d, generator=gen)
    ref = F.relu(F.layer_norm(y + y2, (d,), g, b, 1e-5))
    f, h = ops.residual_layernorm(y.cuda(), y2.cuda(), g.cuda(), b.cuda(), 1e-5, relu=True)
    torch.testing.assert_close(f.cpu(), ref, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(h.float().cpu(), ref, rtol=1e-2, atol=1e-2)


def _build_attention(K, d, seed=0):
    from fewshotobjectdetection_imporove_via_text_feature_b200 import config
    from fewshotobjectdetection_imporove_via_text_feature_b200.modeling.roi_heads.attentive_modules import SematicProposalAttention
    cfg = config.get_cfg()
    cfg.MODEL.ROI_HEADS.NUM_CLASSES = K
    cfg.MODEL.ADDITION.NAME = "clip"
    torch.manual_seed(seed)
    m = SematicProposalAttention(d, cfg=cfg, bg_generator=torch.Generator().manual_seed(5))
    with torch.no_grad():   # reference init (std 0.02) gives near-uniform attention; sharpen it so the test bites
        m.attention.w_q.weight.mul_(6.0)
        m.attention.w_k.weight.mul_(6.0)
        m.key_projection.weight.mul_(3.0)
    return m.eval()


@pytest.mark.parametrize("fold", [True, False])
@pytest.mark.parametrize("R,K", [(512, 20), (200, 80)])
def test_chain_vs_oracle(R, K, fold):
    d = 2048
    m = _build_attention(K, d)
    m.fold_query = fold
    x = torch.relu(torch.randn(R, d, generator=torch.Generator().manual_seed(2)))
    p = {"attention." + k: v.detach() for k, v in m.state_dict().items()}
    text = torch.cat([m.embed, m.bg_feature], 0)
    sim_ref, attn_ref = O.sematic_proposal_attention(x, text, p)
    m = m.cuda()
    with torch.no_grad():
        attn, out = m(x.cuda())
    assert attn.shape == (1, R, K + 2)
    torch.testing.assert_close(attn[0].sum(1).cpu(), torch.ones(R), rtol=1e-4, atol=1e-4)
    assert rel_err(attn[0].cpu(), attn_ref) < 2e-2
    assert rel_err(out["sim2stext"].cpu(), sim_ref) < 2e-2
    # elementwise, north_star's bf16 bar: |got - ref| <= 2e-2 |ref| + 2e-2 max |ref|
    torch.testing.assert_close(out["sim2stext"].cpu().float(), sim_ref, rtol=2e-2, atol=2e-2 * float(sim_ref.abs().max()))
    torch.testing.assert_close(attn[0].cpu().float(), attn_ref, rtol=2e-2, atol=2e-2 * float(attn_ref.abs().max()))


def test_inference_chain_has_no_attention_or_normalisation_kernel():
    """north_star / VERDICT r1 item 10: softmax and the gate operands leave as epilogues of the two products around them, so the
    inference chain's launch list holds GEMMs, one cast and one LayerNorm — and both forms agree to the bf16 bar."""
    from fewshotobjectdetection_imporove_via_text_feature_b200 import _lib, ops
    m = _build_attention(20, 2048).cuda()
    x = torch.relu(torch.randn(640, 2048, generator=torch.Generator().manual_seed(5))).cuda()
    outs = {}
    for flag in (True, False):
        ops.ATTENTION_AS_EPILOGUES[0] = flag
        try:
            with torch.no_grad():
                m(x)
                _lib.PROFILE = {}
                attn, out = m(x)
                names = set(_lib.PROFILE)
        finally:
            _lib.PROFILE = None
            ops.ATTENTION_AS_EPILOGUES[0] = True
        outs[flag] = (attn[0].float(), out["sim2stext"].float(), names)
    assert "b200_text_attention" in outs[False][2] and "b200_text_attention" not in outs[True][2]
    assert outs[True][2] <= {"b200_gemm2", "b200_gemm_bf16", "b200_gemm_bf16_ex", "b200_cast_bf16", "b200_residual_layernorm"}, outs[True][2]
    assert rel_err(outs[True][0], outs[False][0]) < 2e-3 and rel_err(outs[True][1], outs[False][1]) < 2e-2


def test_chain_golden_small(golden):
    """d_model = 32 golden produced by the reference's own SematicProposalAttention."""
    g = golden("attention_small")
    m = _build_attention(20, 32)
    sd = {k[len("attention."):]: torch.from_numpy(g[k]) for k in g.files if k.startswith("attention.")}
    m.load_state_dict(sd)
    m.embed, m.bg_feature = torch.from_numpy(g["embed"]), torch.from_numpy(g["bg_feature"])
    m = m.cuda()
    with torch.no_grad():
        attn, out = m(torch.from_numpy(g["x"]).cuda())
    assert rel_err(attn[0].cpu(), torch.from_numpy(g["attn"])) < 2e-2
    assert rel_err(out["sim2stext"].cpu(), torch.from_numpy(g["sim2stext"])) < 2e-2


def test_cosine_logits_option(golden):
    """Optional cosine + temperature form of the prototype logits: the row-normalisation kernel (fp32 and bf16 sources, the
    eps clamp on a zero row) and normalise -> tcgen05 GEMM against the reference's own `sim_matrix` output (bf16 bar)."""
    from fewshotobjectdetection_imporove_via_text_feature_b200 import ops
    g = golden("cosine")
    a, t, tau = torch.from_numpy(g["a"]).cuda(), torch.from_numpy(g["t"]).cuda(), float(g["tau"])
    an = ops.l2_normalize_rows(a)
    ref_an = a / a.norm(dim=1, keepdim=True).clamp_min(1e-12)
    torch.testing.assert_close(an.float(), ref_an, rtol=2 ** -8, atol=1e-6)
    assert float(an[3].abs().max()) == 0.0
    anb = ops.l2_normalize_rows(a.to(torch.bfloat16))
    torch.testing.assert_close(anb.float(), ref_an, rtol=2e-2, atol=2e-3)
    logits = ops.gemm_bf16(an, ops.l2_normalize_rows(t, scale=tau))
    ref = torch.from_numpy(g["bsim"]).cuda()
    assert rel_err(logits, ref) < 2e-2
    torch.testing.assert_close(logits, ref, rtol=2e-2, atol=2e-2 * tau)
    # the head option: CrossOutput logits become tau * cos(a, T)
    gen = torch.Generator().manual_seed(3)
    x = torch.randn(300, 512, generator=gen).cuda()
    T = torch.randn(81, 512, generator=gen).cuda()
    got = ops.gemm_bf16(ops.l2_normalize_rows(x), ops.l2_normalize_rows(T, scale=tau))
    want = O.sim_matrix(x.cpu(), T.cpu(), tau=tau)
    assert rel_err(got.cpu(), want) < 2e-2


def test_gemm_res5_sized_problem_with_mask():
    """A res5-sized product (65536 rows: 4096 ROIs x 16 pixels; 512 tiles of 128 x 256 on 148 persistent CTAs) with the
    ReLU-backward mask epilogue, against torch's bf16 matmul of the same operands (fp32 accumulate both; bf16 output)."""
    from fewshotobjectdetection_imporove_via_text_feature_b200 import train_ops
    gen = torch.Generator(device="cuda").manual_seed(0)
    M, N, K = 65536, 2048, 512
    a = (torch.randn(M, K, device="cuda", generator=gen) * 0.5).to(torch.bfloat16)
    b = (torch.randn(N, K, device="cuda", generator=gen) * 0.05).to(torch.bfloat16)
    mask = torch.randn(M, N, device="cuda", generator=gen).to(torch.bfloat16)
    out = train_ops.gemm_ex(a, b, out_dtype=torch.bfloat16, mask=mask)
    ref = torch.where(mask > 0, (a @ b.t()).float(), torch.zeros((), device="cuda"))
    assert rel_err(out.float(), ref) < 4e-3
    assert float(out[mask <= 0].abs().max()) == 0.0


@pytest.mark.parametrize("cls_name", ["LV_attention", "LV_attention_VKV"])
@pytest.mark.parametrize("R,d,K", [(1024, 2048, 20), (203, 256, 7), (640, 512, 80)])
def test_teacher_attention_fused_forward_vs_dense_torch(cls_name, R, d, K):
    """A7: the class-collapsed fused forward (no (R, R+1) matrix) vs the module's dense fp32 torch expression
    (attentive_modules.py:403-437 / :452-487 semantics), bf16 bar 2e-2.  Sharpened attention weights so that the
    softmax — and with it the + log n_c / class-mean identity — matters; one class has no ROI at all."""
    from fewshotobjectdetection_imporove_via_text_feature_b200.modeling.roi_heads import teacher_modules as tm
    torch.manual_seed(3)
    gen = torch.Generator().manual_seed(5)
    embed = torch.randn(K, 300, generator=gen)
    from fewshotobjectdetection_imporove_via_text_feature_b200 import config
    cfg = config.get_cfg()
    cfg.MODEL.ADDITION.NAME = "glove"
    m = getattr(tm, cls_name)(d, cfg=cfg, class_embed=embed).cuda().eval()
    with torch.no_grad():
        m.attention.w_q.weight.mul_(12.0)
        m.attention.w_k.weight.mul_(12.0)
    x = torch.relu(torch.randn(R, d, generator=gen)).cuda()
    labels = torch.randint(0, K, (R,), generator=gen)
    labels[torch.rand(R, generator=gen) < 0.6] = K                       # background majority
    labels[labels == 2] = 3                                              # class 2 absent
    labels = labels.cuda()
    with torch.enable_grad():
        _, ref = m(x, labels)                                            # dense torch path
    with torch.no_grad():
        _, out = m(x, labels)                                            # fused path
    got, want = out["sim2stext"], ref["sim2stext"].detach()
    assert got.shape == want.shape == (1, R, d)
    rel = float((got - want).norm() / want.norm())
    assert rel < 2e-2, rel
    assert out["text_feat"].shape == ref["text_feat"].shape
    torch.testing.assert_close(out["text_feat"], ref["text_feat"].detach(), rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("cls_name", ["LV_attention_textDomination", "LV_attention_textDomination_VKV"])
@pytest.mark.parametrize("R,d,K", [(1024, 2048, 20), (203, 256, 7), (640, 512, 80)])
def test_text_domination_fused_forward_vs_dense_torch(cls_name, R, d, K):
    """A7: the 300-d text-space teachers on the CTA-pair GEMM (operands in pitch-304 / 608 buffers, contraction over
    exactly 300 / 150 columns) vs the module's dense fp32 torch expression (attentive_modules.py:597-634 / :650-687),
    bf16 bar 2e-2, sharpened softmax, one class absent."""
    from fewshotobjectdetection_imporove_via_text_feature_b200.modeling.roi_heads import teacher_modules as tm
    from fewshotobjectdetection_imporove_via_text_feature_b200 import config
    torch.manual_seed(4)
    gen = torch.Generator().manual_seed(6)
    embed = torch.randn(K, 300, generator=gen)
    cfg = config.get_cfg()
    cfg.MODEL.ADDITION.NAME = "glove"
    m = getattr(tm, cls_name)(d, cfg=cfg, class_embed=embed).cuda().eval()
    with torch.no_grad():
        m.attention.w_q.weight.mul_(10.0)
        m.attention.w_k.weight.mul_(10.0)
    x = torch.relu(torch.randn(R, d, generator=gen)).cuda()
    labels = torch.randint(0, K, (R,), generator=gen)
    labels[torch.rand(R, generator=gen) < 0.6] = K
    labels[labels == 2] = 3
    labels = labels.cuda()
    with torch.enable_grad():
        _, ref = m(x, labels)
    with torch.no_grad():
        _, out = m(x, labels)
    got, want = out["sim2stext"], ref["sim2stext"].detach()
    assert got.shape == want.shape == (1, R, d)
    want0 = ref["sim2stext"].detach()
    assert float((want0 - want0.mean(1, keepdim=True)).abs().max()) > 1e-3   # rows differ: the attention matters
    rel = float((got - want).norm() / want.norm())
    assert rel < 2e-2, rel
    torch.testing.assert_close(out["text_feat"], ref["text_feat"].detach(), rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("tag,name", [("lv", "LV_attention"), ("td", "LV_attention_textDomination")])
def test_teacher_fused_forward_matches_reference_golden(golden, tag, name):
    """The frozen-teacher kernel path against the REFERENCE's own forward (tests/golden/teacher.npz, generated by importing
    the reference; weights replayed by seed), bf16 bar 2e-2."""
    import numpy as np
    from fewshotobjectdetection_imporove_via_text_feature_b200 import config
    from fewshotobjectdetection_imporove_via_text_feature_b200.modeling import roi_heads as RH
    from oracle.gen_golden import seeded_fill
    g = golden("teacher")
    cfg = config.get_cfg()
    cfg.MODEL.ADDITION.NAME = "glove"
    m = getattr(RH, name)(32, cfg=cfg, class_embed=torch.from_numpy(g[tag + "_embed"])).eval()
    seeded_fill(m, 77)
    m = m.cuda()
    with torch.no_grad():
        _, out = m(torch.from_numpy(g["x"]).cuda(), torch.from_numpy(g["lab"]).cuda())
    want = torch.from_numpy(np.asarray(g[tag + "_sim2stext"])).cuda()
    got = out["sim2stext"].reshape(want.shape)
    rel = float((got - want).norm() / want.norm())
    assert rel < 2e-2, rel


def test_class_mean_rows_and_gather_rows():
    from fewshotobjectdetection_imporove_via_text_feature_b200 import ops
    gen = torch.Generator().manual_seed(8)
    R, d, C = 777, 320, 9
    x = torch.randn(R, d, generator=gen).cuda()
    lab = torch.randint(0, C - 1, (R,), generator=gen).cuda()           # last class empty
    mean, cnt = ops.class_mean_rows(x, lab, C)
    ref = torch.zeros(C, d, dtype=torch.float64, device="cuda").index_add_(0, lab, x.double())
    n = torch.bincount(lab, minlength=C)
    assert torch.equal(cnt.long(), n)
    torch.testing.assert_close(mean.double(), ref / n.clamp(min=1)[:, None], rtol=1e-5, atol=1e-6)
    assert float(mean[C - 1].abs().max()) == 0.0
    m2, _ = ops.class_mean_rows(x, lab, C)
    assert torch.equal(mean, m2)                                         # fixed summation order
    table = torch.randn(C, d, generator=gen).cuda()
    dst = torch.zeros(R, 2 * d, dtype=torch.bfloat16, device="cuda")
    ops.gather_rows_bf16(table, lab, dst[:, d:], relu=True)
    assert torch.equal(dst[:, d:], torch.relu(table[lab]).to(torch.bfloat16)) and float(dst[:, :d].abs().max()) == 0.0
