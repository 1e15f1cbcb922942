import sys, torch
sys.path.insert(0, "/root/repo")
from fewshotobjectdetection_imporove_via_text_feature_b200 import ops
torch.manual_seed(0)
dev = "cuda"
rnd = lambda *s, sc=0.5: (torch.randn(*s, device=dev) * sc).to(torch.bfloat16)
def chk(name, got, ref):
    d = (got.double() - ref).abs()
    rows = (d.max(1).values > 2e-2 * ref.abs().max()).nonzero().flatten()
    print("%-28s rel %.2e max %.2e bad rows %d  first %s" % (name, float((got.double() - ref).norm() / ref.norm()), float(d.max()), rows.numel(), rows[:12].tolist()))
    if rows.numel():
        r = int(rows[0]); cols = (d[r] > 2e-2 * ref.abs().max()).nonzero().flatten()
        print("    row", r, "bad cols", cols.numel(), cols[:8].tolist(), cols[-4:].tolist())
for R in (1024, 4096):
    dd = rnd(R, 80); Wb = rnd(80, 2048, sc=0.05)
    dx = torch.empty(R, 2048, device=dev)
    ops.gemm2(dd[:, :80], Wb, b_mn=True, out_f32=dx, want_out=False)
    chk("R=%d dd@Wb K=80" % R, dx, dd.double() @ Wb.double())
    dyb = rnd(R, 2048); W3 = rnd(2048, 4096, sc=0.05)
    base = dx.double().clone()
    ops.gemm2(dyb, W3[:, 2048:], b_mn=True, out_f32=dx, accumulate=True, want_out=False)
    chk("R=%d += dyb@W3[:,d:]" % R, dx, base + dyb.double() @ W3[:, 2048:].double())
    dS = rnd(R, 24); kq = rnd(22, 2048, sc=0.05)
    base = dx.double().clone()
    ops.gemm2(dS[:, :22], kq, b_mn=True, out_f32=dx, accumulate=True, want_out=False)
    chk("R=%d += dS@kq K=22" % R, dx, base + dS[:, :22].double() @ kq.double())
    dl = rnd(R, 24); Wc = rnd(21, 2048, sc=0.05)
    o = ops.gemm2(dl[:, :21], Wc, b_mn=True)
    chk("R=%d dzd = dl@Wc K=21 bf16" % R, o, dl[:, :21].double() @ Wc.double())
