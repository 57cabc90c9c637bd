"""TEST DOUBLE (never imported by the product): a torch-CPU stand-in for ``radzero_b200.ops``
with the same call signatures, so that the HOST logic of the sharded contrastive step
(ragged gathers, column offsets, all-reduces, gradient routing and scaling) can be tested with
``gloo`` on machines without a GPU.  The math follows ``oracle/vlcabs.py``; backward pieces use
autograd instead of the closed forms the CUDA kernels implement."""
import torch
import torch.nn.functional as F

HIDDEN = 768


def padded_tokens_bwd(tokens: int) -> int:
    return (tokens + 127) // 128 * 128


def _norm(x, g, b, l2):
    y = F.layer_norm(x, (HIDDEN,), g, b, 1e-5) if g is not None else x
    return F.normalize(y, p=2, dim=-1, eps=1e-12) if l2 else y


def prep_rows(x, gamma, beta, *, rows_per_group=None, rows_per_group_padded=None, l2=True, **kw):
    x2 = x.reshape(-1, HIDDEN).double()
    rows = x2.shape[0]
    rpg = rows_per_group or max(rows, 1)
    rpp = rows_per_group_padded or rpg
    k = _norm(x2, None if gamma is None else gamma.double(), None if beta is None else beta.double(), l2)
    out = k.new_zeros((rows // rpg, rpp, HIDDEN))
    out[:, :rpg] = k.view(rows // rpg, rpg, HIDDEN)
    return out.view(-1, HIDDEN), None, None


def _fwd(k, q, scale):
    s = torch.einsum("nd,bld->bnl", q, k) * scale
    p = torch.softmax(s, dim=-1)
    o = torch.einsum("bnl,bld->bnd", p, k)
    z = (q.unsqueeze(0) * F.normalize(o, dim=-1)).sum(-1).T
    return s, o, z


def sim_fwd(k16, q16, tokens, scale, *, log_tau_scale=None, want_scores=False, drop_cls=True,
            want_stats=False, want_pooled=False, **kw):
    if log_tau_scale is not None:
        scale = float(torch.exp(-log_tau_scale.double()))
    k = k16[:, :tokens]
    s, o, z = _fwd(k, q16, scale)
    return dict(z=z.contiguous(), scores=(s[:, :, 1:] if drop_cls else s) if want_scores else None,
                lse=torch.logsumexp(s, -1), onorm=o.norm(dim=-1), pooled=o)


def mpnce_partials(z, group_map, col0, inv_tau=1.0, *, log_tau=None, rowpos=None, eps=1e-8, col_sum=False,
                   b_global=None):
    if log_tau is not None:
        inv_tau = float(torch.exp(-log_tau.double()))
    n, bl = z.shape
    E = torch.exp(z * inv_tau)
    g = group_map - col0
    local = (g >= 0) & (g < bl)
    pos = torch.zeros(n, dtype=z.dtype)
    idx = torch.nonzero(local).flatten()
    pos[idx] = E[idx, g[idx]]
    onehot = torch.zeros_like(E)
    onehot[idx, g[idx]] = 1.0
    rs = E.sum(1)
    if rowpos is not None:          # the caller's (2, n) all-reduce buffer (fp32 in the product)
        rowpos[0].copy_(rs)
        rowpos[1].copy_(pos)
        rs, pos = rowpos[0].to(z.dtype), rowpos[1].to(z.dtype)
    return rs, pos, (E * (1 - onehot)).sum(0), (E * onehot).sum(0)


def mpnce_finish(z, group_map, col0, b_global, inv_tau, rowsum, pos, colneg, colpos, *, eps=1e-8,
                 row_sum=False, col_sum=False, want_dz=True, log_tau=None):
    assert not row_sum and not col_sum, "the CPU stand-in covers the radzero configuration"
    if log_tau is not None:
        inv_tau = float(torch.exp(-log_tau.double()))
    with torch.enable_grad():
        return _mpnce_finish(z, group_map, col0, inv_tau, rowsum.to(z.dtype), pos.to(z.dtype), eps)


def _mpnce_finish(z, group_map, col0, inv_tau, rowsum, pos, eps):
    n, bl = z.shape
    zz = z.detach().clone().requires_grad_(True)
    E = torch.exp(zz * inv_tau)
    g = group_map - col0
    local = (g >= 0) & (g < bl)
    idx = torch.nonzero(local).flatten()
    # global row sums: other ranks' share is a constant, the local share carries the gradient
    R = (rowsum - E.detach().sum(1)) + E.sum(1)
    P = pos.clone()
    Pl = E[idx, g[idx]]
    P = P.index_put((idx,), Pl)
    row_terms = -torch.log(P / (R + eps) + eps)
    onehot = torch.zeros_like(E)
    onehot[idx, g[idx]] = 1.0
    cneg = (E * (1 - onehot)).sum(0)
    col_terms = -torch.log(Pl / (Pl + cneg[g[idx]] + eps) + eps)
    total = (row_terms.sum() + col_terms.sum()) / (2 * n)
    total.backward()
    dz = zz.grad
    rt, ct = row_terms[idx].sum().detach(), col_terms.sum().detach()
    terms = torch.stack([rt, ct, (dz * z).sum(), (rt + ct) / (2 * n)])
    return terms, dz


def sim_bwd(k16, q16, tokens, inv_tau, z, dz, lse, onorm, pooled, *, log_tau=None, p=None, mref=None, lsum=None):
    with torch.enable_grad():
        return _sim_bwd(k16, q16, tokens, inv_tau, dz, log_tau)


def _sim_bwd(k16, q16, tokens, inv_tau, dz, log_tau):
    k = k16.detach().clone().requires_grad_(True)
    q = q16.detach().clone().requires_grad_(True)
    lt = (log_tau.detach().double().clone() if log_tau is not None
          else torch.log(torch.tensor([1.0 / inv_tau], dtype=torch.float64))).requires_grad_(True)
    _, _, zz = _fwd(k[:, :tokens], q, torch.exp(-lt))
    (zz * dz).sum().backward()
    return q.grad, k.grad, lt.grad.reshape(1)


def prep_rows_bwd(x, gamma, beta, dnorm, *, rows_per_group=None, rows_per_group_padded=None, l2=True,
                  dgamma=None, dbeta=None, accumulate=False, native_dx=False):
    with torch.enable_grad():
        return _prep_rows_bwd(x, gamma, beta, dnorm, rows_per_group, rows_per_group_padded, l2, dgamma,
                              dbeta, accumulate)


def _prep_rows_bwd(x, gamma, beta, dnorm, rows_per_group, rows_per_group_padded, l2, dgamma, dbeta,
                   accumulate):
    x2 = x.reshape(-1, HIDDEN).double().detach().clone().requires_grad_(True)
    rows = x2.shape[0]
    rpg = rows_per_group or max(rows, 1)
    rpp = rows_per_group_padded or rpg
    g = gamma.double().detach().clone().requires_grad_(True) if gamma is not None else None
    b = beta.double().detach().clone().requires_grad_(True) if beta is not None else None
    k = _norm(x2, g, b, l2)
    d = dnorm.reshape(rows // rpg, rpp, HIDDEN)[:, :rpg].reshape(rows, HIDDEN)
    (k * d).sum().backward()
    if g is not None:
        if accumulate and dgamma is not None:
            dgamma = dgamma + g.grad
            dbeta = dbeta + b.grad
        else:
            dgamma, dbeta = g.grad, b.grad
    return x2.grad, dgamma, dbeta
