"""B2 seam on the GPU: SdOptimizationSdxe.apply(classes=...) installs the sdxe-backed forwards on CrossAttention / AttnBlock
module classes (here: the oracle's structural twins of ldm's) and reproduces the reference's SDP forwards
(modules/sd_hijack_optimizations.py:508-546, 637-655) within 16-bit tolerance; undo() restores the stock forwards."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def rel_err(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / (b.norm() + 1e-12)).item()


def test_sd_optimization_sdxe_apply_and_undo(cuda):
    from oracle.synth import init_module_
    from oracle.unet import CrossAttention
    from oracle.vae import AttnBlock
    from sdwebui_b200.sd_hijack_optimizations import SdOptimizationSdxe

    g = torch.Generator(device="cuda").manual_seed(2)
    cases = []
    for qd, cd, heads, dh, n, nk in ((320, None, 8, 40, 1024, None), (320, 768, 8, 40, 1024, 77), (640, 2048, 10, 64, 1024, 77), (1280, None, 8, 160, 64, None)):
        m = init_module_(CrossAttention(qd, cd, heads, dh), 7).eval().to(cuda).half()
        x = torch.randn(2, n, qd, device=cuda, generator=g).half()
        ctx = None if cd is None else torch.randn(2, nk, cd, device=cuda, generator=g).half()
        cases.append((m, x, ctx))
    ab = init_module_(AttnBlock(512), 8).eval().to(cuda).half()   # the VAE mid block: one head of d = 512
    xa = torch.randn(2, 512, 32, 32, device=cuda, generator=g).half()
    with torch.no_grad():
        ref = [m(x, context=c) for m, x, c in cases]
        ref32 = [m.float()(x.float(), context=None if c is None else c.float()) for m, x, c in cases]
        for m, _, _ in cases:
            m.half()
        ref_ab, ref_ab32 = ab(xa), ab.float()(xa.float())
        ab.half()
    stock_ca, stock_ab = CrossAttention.forward, AttnBlock.forward
    opt = SdOptimizationSdxe()
    assert opt.is_available()
    opt.apply(classes={"CrossAttention": [CrossAttention], "AttnBlock": [AttnBlock]})
    try:
        assert CrossAttention.forward is not stock_ca and AttnBlock.forward is not stock_ab
        with torch.no_grad():
            for (m, x, c), r16, r32 in zip(cases, ref, ref32):
                out = m(x, context=c)
                e, e_ref = rel_err(out, r32), rel_err(r16, r32)
                print(f"CrossAttention heads {m.heads} ctx {None if c is None else tuple(c.shape)}: sdxe {e:.3e}  torch-SDP fp16 {e_ref:.3e}")
                assert out.shape == r16.shape and e < max(3 * e_ref, 2e-3)
            out = ab(xa)
            e, e_ref = rel_err(out, ref_ab32), rel_err(ref_ab, ref_ab32)
            print(f"AttnBlock d=512 N=1024: sdxe {e:.3e}  torch-SDP fp16 {e_ref:.3e}")
            assert e < max(3 * e_ref, 2e-3)
    finally:
        opt.undo()
    assert CrossAttention.forward is stock_ca and AttnBlock.forward is stock_ab
