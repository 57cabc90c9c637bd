"""TEST DOUBLE (never imported by the product): a torch-CPU stand-in, in fp64, for the ``radzero_b200.ops``
entry points the AlignTransformer forward / backward calls, with the same signatures -- so that the HOST
logic of ``radzero_b200.align`` (what is saved, the order of the products, LayerScale folded into the
transposed weights, the folded 1/8 of the query projection, the 2^k gradient scale, the order of the
returned gradients) can be checked against torch autograd without a GPU.  Each function states the
contract of the CUDA kernel it stands in for; derivatives come from autograd, not from the kernels' closed forms."""
import math

import torch
import torch.nn.functional as F

HIDDEN = 768


def ln_rows(x, gamma, beta, eps):
    return F.layer_norm(x.double(), (HIDDEN,), gamma.double(), beta.double(), eps)


def linear(a, w, bias, epilogue="bias", *, scale=None, residual=None, out=None):
    acc = a.double() @ w.double().T
    if bias is not None:
        acc = acc + bias.double()
    if epilogue == "bias":
        return acc
    if epilogue == "gelu":
        return F.gelu(acc)
    res = residual.double() + (acc if scale is None else scale.double() * acc)
    if out is not None:
        out.copy_(res)
        return out
    return res


def _heads(t, B, L, heads):
    return t.reshape(B, L, heads, 64).transpose(1, 2)


def _attn(qkv, heads):
    B, L, W = qkv.shape
    q, k, v = [_heads(t, B, L, heads) for t in qkv.split(W // 3, dim=-1)]
    p = torch.softmax(q @ k.transpose(2, 3), dim=-1)          # the 1/sqrt(64) is folded into q
    return (p @ v).transpose(1, 2).reshape(B, L, W // 3)


def attention(qkv, heads):
    return _attn(qkv.double(), heads)


def attention_bwd(qkv, out16, dout16, heads, q_scale):
    x = qkv.double().detach().requires_grad_(True)
    with torch.enable_grad():              # called from inside an autograd.Function's backward
        o = _attn(x, heads)
    o.backward(dout16.double())
    g = x.grad.clone()
    g[..., : g.shape[-1] // 3] *= q_scale                     # gradient of the UNSCALED query projection
    return g


def grad_scale(grad):
    amax = grad.abs().max().item()
    k = 0
    if 0 < amax < 3e38:
        k = max(-60, min(60, 6 - (math.frexp(amax)[1])))
    sc = torch.empty(4 + 3072, dtype=torch.float64)
    sc[0], sc[1], sc[2], sc[3] = 2.0 ** k, 2.0 ** -k, amax, k
    sc[4:] = 2.0 ** -k
    return sc


def ls_cast_bwd(dy, ls, o16, sc, dls):
    out = dy.double() * sc[0]
    if ls is not None:
        out = out * ls.double()
    if dls is not None and o16 is not None:
        dls += (dy.double() * o16.double()).sum(0)
    return out


def ls_weight_bwd(g, w, bias, colsum, ls):
    dls = (w.double() * g.double()).sum(1) + (0 if bias is None else bias.double() * colsum.double())
    g.mul_(ls[:, None])
    colsum.mul_(ls)
    return dls


def transpose_pad(x16, sc=None, colsum=None, want_out=True):
    if colsum is not None:
        colsum += x16.double().sum(0) * sc[1]
    if not want_out:
        return None
    rows, cols = x16.shape
    out = x16.new_zeros((cols, (rows + 63) // 64 * 64))
    out[:, :rows] = x16.T
    return out


def gelu_bwd(dg16, u16):
    u = u16.double().detach().requires_grad_(True)
    with torch.enable_grad():
        g = F.gelu(u)
    g.backward(dg16.double())
    return u.grad


def ln_rows_bwd(x, dh16, gamma, eps, dres, sc, dgamma, dbeta, out=None, out16=None):
    xd = x.double().detach().requires_grad_(True)
    g = gamma.double().detach().requires_grad_(True)
    b = torch.zeros_like(g).requires_grad_(True)
    with torch.enable_grad():
        h = F.layer_norm(xd, (HIDDEN,), g, b, eps)
    h.backward(dh16.double() * sc[1])
    res = xd.grad if dres is None else dres.double() + xd.grad
    if dgamma is not None:
        dgamma += g.grad
    if dbeta is not None:
        dbeta += b.grad
    if out16 is not None:
        out16.copy_(res * sc[0])
    if out is not None:
        out.copy_(res)
        return out
    return res
