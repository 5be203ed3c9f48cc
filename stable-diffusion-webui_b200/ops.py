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
    L.check(L.load().sdxe_gemm(L.ptr(a.contiguous()), L.ptr(w.contiguous()), L.ptr(out), M, N, K, L.ptr(bias),
                               L.ptr(residual), flags, L.torch_dtype_code(a.dtype), L.current_stream()), "sdxe_gemm")
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
    L.check(L.load().sdxe_conv3x3_nhwc(L.ptr(x.contiguous()), L.ptr(wp), L.ptr(out), n, h, wd, cin, cout, L.ptr(bias),
                                       L.torch_dtype_code(x.dtype), L.current_stream()), "sdxe_conv3x3_nhwc")
    return out


def group_norm_nhwc(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, groups: int = 32, eps: float = 1e-5,
                    silu: bool = False) -> torch.Tensor:
    _require_cuda(x)
    n, h, wd, c = x.shape
    out = torch.empty_like(x)
    L.check(L.load().sdxe_group_norm_nhwc(L.ptr(x.contiguous()), L.ptr(gamma.float().contiguous()),
                                          L.ptr(beta.float().contiguous()), L.ptr(out), n, h * wd, c, groups, eps,
                                          1 if silu else 0, L.torch_dtype_code(x.dtype), L.current_stream()),
            "sdxe_group_norm_nhwc")
    return out


def layer_norm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    _require_cuda(x)
    c = x.shape[-1]
    rows = x.numel() // c
    out = torch.empty_like(x)
    L.check(L.load().sdxe_layer_norm(L.ptr(x.contiguous()), L.ptr(gamma.float().contiguous()),
                                     L.ptr(beta.float().contiguous()), L.ptr(out), rows, c, eps,
                                     L.torch_dtype_code(x.dtype), L.current_stream()), "sdxe_layer_norm")
    return out
