"""Torch-tensor wrappers over the stand-alone C-ABI primitives (attention, GEMM, conv, norms).

Device memory and streams come from PyTorch (plumbing); all arithmetic happens inside libsdxe.so.
"""
from __future__ import annotations

import torch

from . import lib as L


def _require_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise L.SdxeError("sdxe ops need CUDA tensors: there is no CPU fallback")


def attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, scale: float | None = None) -> torch.Tensor:
    """softmax(q k^T * scale) v.  q [B,H,Nq,D], k/v [B,H,Nk,D] -> [B, Nq, H*D]
    (the layout `scaled_dot_product_attention_forward` reshapes to, sd_hijack_optimizations.py:539)."""
    _require_cuda(q, k, v)
    B, H, Nq, D = q.shape
    Nk = k.shape[2]
    q, k, v = q.contiguous(), k.contiguous(), v.contiguous()
    out = torch.empty(B, Nq, H * D, dtype=q.dtype, device=q.device)
    sc = float(scale) if scale is not None else float(D) ** -0.5
    L.check(L.load().sdxe_attention(L.ptr(q), L.ptr(k), L.ptr(v), L.ptr(out), B, H, Nq, Nk, D, sc,
                                    L.torch_dtype_code(q.dtype), L.current_stream()), "sdxe_attention")
    return out


def gemm(a: torch.Tensor, w: torch.Tensor, bias: torch.Tensor | None = None, residual: torch.Tensor | None = None,
         geglu: bool = False, force_bn: int = 0) -> torch.Tensor:
    """a [M,K] @ w[N,K]^T (+bias fp32) (+residual) ; geglu: out [M, N/2]."""
    _require_cuda(a, w)
    M, K = a.shape
    N = w.shape[0]
    out = torch.empty(M, N // 2 if geglu else N, dtype=a.dtype, device=a.device)
    if bias is not None:
        bias = bias.float().contiguous()
    flags = (1 if geglu else 0) | (force_bn << 8)
    # every converted tensor is bound to a local that outlives the C call: a temporary passed straight into L.ptr()
    # is freed on return from ptr() and the caching allocator may hand its block to the next temporary
    a_c, w_c = a.contiguous(), w.contiguous()
    res_c = residual.contiguous() if residual is not None else None
    L.check(L.load().sdxe_gemm(L.ptr(a_c), L.ptr(w_c), L.ptr(out), M, N, K, L.ptr(bias),
                               L.ptr(res_c), flags, L.torch_dtype_code(a.dtype), L.current_stream()), "sdxe_gemm")
    del a_c, w_c, res_c
    return out


def conv3x3_nhwc(x: torch.Tensor, w: torch.Tensor, bias: torch.Tensor | None = None) -> torch.Tensor:
    """x [n,h,w,cin] (NHWC), w [cout, cin, 3, 3] (torch layout; repacked tap-major here) -> [n,h,w,cout]."""
    _require_cuda(x, w)
    n, h, wd, cin = x.shape
    cout = w.shape[0]
    wp = w.permute(0, 2, 3, 1).contiguous().reshape(cout, 9 * cin)
    out = torch.empty(n, h, wd, cout, dtype=x.dtype, device=x.device)
    if bias is not None:
        bias = bias.float().contiguous()
    x_c = x.contiguous()
    L.check(L.load().sdxe_conv3x3_nhwc(L.ptr(x_c), L.ptr(wp), L.ptr(out), n, h, wd, cin, cout, L.ptr(bias),
                                       L.torch_dtype_code(x.dtype), L.current_stream()), "sdxe_conv3x3_nhwc")
    del x_c
    return out


def group_norm_nhwc(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, groups: int = 32, eps: float = 1e-5,
                    silu: bool = False) -> torch.Tensor:
    _require_cuda(x)
    n, h, wd, c = x.shape
    x_c, g, b = x.contiguous(), gamma.float().contiguous(), beta.float().contiguous()
    out = torch.empty_like(x_c)
    L.check(L.load().sdxe_group_norm_nhwc(L.ptr(x_c), L.ptr(g), L.ptr(b), L.ptr(out), n, h * wd, c, groups, eps,
                                          1 if silu else 0, L.torch_dtype_code(x.dtype), L.current_stream()),
            "sdxe_group_norm_nhwc")
    del x_c, g, b
    return out


def layer_norm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    _require_cuda(x)
    c = x.shape[-1]
    rows = x.numel() // c
    x_c, g, b = x.contiguous(), gamma.float().contiguous(), beta.float().contiguous()
    out = torch.empty_like(x_c)
    L.check(L.load().sdxe_layer_norm(L.ptr(x_c), L.ptr(g), L.ptr(b), L.ptr(out), rows, c, eps,
                                     L.torch_dtype_code(x.dtype), L.current_stream()), "sdxe_layer_norm")
    del x_c, g, b
    return out
