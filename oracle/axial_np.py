"""CPU ORACLE (test infrastructure, NOT product code) -- numpy restatement of the
reference's axial-attention op, forward and analytic backward.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
leg may import this module.  The product path (medical-transformer_b200/) never does.

Parity status: PINNED.  tests/test_oracle.py checks every function here against
golden vectors produced by executing the reference modules themselves
(tests/golden/make_golden.py imports /root/reference/lib/models/axialnet.py).

What is restated (reference file:line, all in lib/models/axialnet.py):
  AxialAttention_dynamic.forward  :142-189   variant "gated"  (pos-emb + 4 gates)
  AxialAttention.forward          :52-92     variant "pos"    (pos-emb, gates == 1)
  AxialAttention_wopos.forward    :222-253   variant "nopos"  (qk only, BN2d(G), BN1d(C))
  Toeplitz indexing of `relative` :131-135,155  E[c,i,j] = relative[c, i-j+L-1]
The backward is not written anywhere in the reference (it is autograd over the ops
above); the closed forms below were derived by hand and are pinned against
torch.autograd of the reference module in fp64 by the golden fixtures.

Extra (product feature, SURVEY 7 hard part 4): `ngroups` > 1 splits the N images into
`ngroups` consecutive blocks that get SEPARATE batch statistics -- this is how the 16
LoGo patches (medt_net._forward_impl :661-700, one BN call per patch) are batched.
"""
from __future__ import annotations

import numpy as np

GROUPS = 8  # reference hard-codes groups=8 at every call site (axialnet.py:400,715-728)


# --------------------------------------------------------------------------- helpers
def _to_strips(x, axis):
    """NCHW -> (B, C, L) strips.  axis=0: attend along H (width=False, :146-148);
    axis=1: attend along W (width=True, :143-144)."""
    n, c, h, w = x.shape
    if axis == 0:
        return np.ascontiguousarray(x.transpose(0, 3, 1, 2)).reshape(n * w, c, h)
    return np.ascontiguousarray(x.transpose(0, 2, 1, 3)).reshape(n * h, c, w)


def _from_strips(s, n, axis):
    """(B, C, L) -> NCHW (inverse of _to_strips; reference :181-184)."""
    b, c, l = s.shape
    other = b // n
    s = s.reshape(n, other, c, l)
    if axis == 0:
        return np.ascontiguousarray(s.transpose(0, 2, 3, 1))
    return np.ascontiguousarray(s.transpose(0, 2, 1, 3))


def toeplitz(rel, L):
    """E[c,i,j] = rel[c, i-j+L-1]  (reference index_select with flatten_index, :155)."""
    i = np.arange(L)[:, None]
    j = np.arange(L)[None, :]
    return rel[:, i - j + L - 1]


def _bn_train(x, red_axes, gamma, beta, eps):
    """Batch-stat BN over `red_axes`; x carries a leading stat-group axis that is never
    reduced.  Returns y, mean, biased var (keepdims)."""
    mean = x.mean(axis=red_axes, keepdims=True)
    var = x.var(axis=red_axes, keepdims=True)
    xhat = (x - mean) / np.sqrt(var + eps)
    return xhat * gamma + beta, mean, var


def _ema(running, stat_per_group, momentum):
    """Sequential running-stat update, one per stat group in order (what 16 successive
    BN calls do in the LoGo loop)."""
    r = running.copy()
    for s in stat_per_group:
        r = (1.0 - momentum) * r + momentum * s
    return r


# --------------------------------------------------------------------------- forward
def axial_forward(x, p, axis, stride=1, variant="gated", training=True, ngroups=1,
                  eps=1e-5, momentum=0.1):
    """One axial-attention op.

    x : (N, C, H, W) float array.
    p : dict with
        w_qkv (2C, C); relative (2gp, 2L-1) [not for nopos]; gates = (f_qr, f_kr, f_sv, f_sve)
        [gated only]; bn_qkv_{w,b,rm,rv} (2C); bn_sim_{w,b,rm,rv} (3G or G);
        bn_out_{w,b,rm,rv} (2C or C).
    Returns out (N, C, H/stride, W/stride), cache (for backward), new_running (dict of
    updated running stats, only when training).
    """
    x = np.asarray(x)
    dt = x.dtype
    n, c, h, w = x.shape
    G = GROUPS
    gp = c // G
    d = gp // 2
    L = h if axis == 0 else w
    other = w if axis == 0 else h
    B = n * other
    pos = variant != "nopos"
    nt = 3 if pos else 1
    if variant == "gated":
        f_qr, f_kr, f_sv, f_sve = [dt.type(v) for v in p["gates"]]
    else:
        f_qr = f_kr = f_sv = f_sve = dt.type(1.0)

    X = _to_strips(x, axis)                                   # (B, C, L)
    Wq = np.asarray(p["w_qkv"], dtype=dt).reshape(2 * c, c)
    y = np.einsum("oc,bcl->bol", Wq, X)                       # conv1d k=1 (:151)

    Bg = B // ngroups
    yg = y.reshape(ngroups, Bg, 2 * c, L)
    g1 = np.asarray(p["bn_qkv_w"], dtype=dt)[None, None, :, None]
    b1 = np.asarray(p["bn_qkv_b"], dtype=dt)[None, None, :, None]
    new_running = {}
    if training:
        yn, mu1, var1 = _bn_train(yg, (1, 3), g1, b1, eps)
        cnt = Bg * L
        new_running["bn_qkv_rm"] = _ema(np.asarray(p["bn_qkv_rm"], dtype=dt), mu1.reshape(ngroups, -1), momentum)
        new_running["bn_qkv_rv"] = _ema(np.asarray(p["bn_qkv_rv"], dtype=dt),
                                        var1.reshape(ngroups, -1) * (cnt / (cnt - 1)), momentum)
    else:
        mu1 = np.asarray(p["bn_qkv_rm"], dtype=dt)[None, None, :, None]
        var1 = np.asarray(p["bn_qkv_rv"], dtype=dt)[None, None, :, None]
        yn = (yg - mu1) / np.sqrt(var1 + eps) * g1 + b1
    yn = yn.reshape(B, G, 2 * gp, L)
    q, k, v = yn[:, :, :d], yn[:, :, d:2 * d], yn[:, :, gp:]   # split (:152)

    S = [np.einsum("bgci,bgcj->bgij", q, k)]                   # qk (:159)
    if pos:
        rel = np.asarray(p["relative"], dtype=dt)
        E = toeplitz(rel, L)
        Eq, Ek, Ev = E[:d], E[d:2 * d], E[gp:]
        S.append(f_qr * np.einsum("bgci,cij->bgij", q, Eq))    # qr (:157,163)
        S.append(f_kr * np.einsum("bgcj,cji->bgij", k, Ek))    # kr, transposed (:158,164)
    S = np.stack(S, axis=1)                                    # (B, nt, G, L, L) == cat on channel (:166)

    Sg = S.reshape(ngroups, Bg, nt, G, L, L)
    g2 = np.asarray(p["bn_sim_w"], dtype=dt).reshape(1, 1, nt, G, 1, 1)
    b2 = np.asarray(p["bn_sim_b"], dtype=dt).reshape(1, 1, nt, G, 1, 1)
    if training:
        Sn, mu2, var2 = _bn_train(Sg, (1, 4, 5), g2, b2, eps)
        cnt = Bg * L * L
        new_running["bn_sim_rm"] = _ema(np.asarray(p["bn_sim_rm"], dtype=dt), mu2.reshape(ngroups, -1), momentum)
        new_running["bn_sim_rv"] = _ema(np.asarray(p["bn_sim_rv"], dtype=dt),
                                        var2.reshape(ngroups, -1) * (cnt / (cnt - 1)), momentum)
    else:
        mu2 = np.asarray(p["bn_sim_rm"], dtype=dt).reshape(1, 1, nt, G, 1, 1)
        var2 = np.asarray(p["bn_sim_rv"], dtype=dt).reshape(1, 1, nt, G, 1, 1)
        Sn = (Sg - mu2) / np.sqrt(var2 + eps) * g2 + b2
    Z = Sn.sum(axis=2).reshape(B, G, L, L)                     # sum of the normalised terms (:167)
    Z = Z - Z.max(axis=3, keepdims=True)
    P = np.exp(Z)
    P /= P.sum(axis=3, keepdims=True)                          # softmax over keys j (:170)

    O = [f_sv * np.einsum("bgij,bgcj->bgci", P, v)]            # sv (:171,175)
    if pos:
        O.append(f_sve * np.einsum("bgij,cij->bgci", P, Ev))   # sve (:172,176)
    no = len(O)
    O = np.stack(O, axis=-1)                                   # (B, G, gp, L, no): cat(dim=-1) then view (:178)
    # reference view: (B, G, gp, L*no) -> (B, C*no, L): channel index = (g*gp+c)*no + s  when the
    # trailing [L, no]-cat is re-read as [no, L]: cat([sv, sve], dim=-1) has last dim 2L, i.e.
    # sv occupies [0,L), sve [L,2L)  => channel (g*gp+c)*2 + s, position l.
    O = np.ascontiguousarray(np.moveaxis(O, -1, 3)).reshape(B, c * no, L)

    Og = O.reshape(ngroups, Bg, c * no, L)
    g3 = np.asarray(p["bn_out_w"], dtype=dt)[None, None, :, None]
    b3 = np.asarray(p["bn_out_b"], dtype=dt)[None, None, :, None]
    if training:
        On, mu3, var3 = _bn_train(Og, (1, 3), g3, b3, eps)
        cnt = Bg * L
        new_running["bn_out_rm"] = _ema(np.asarray(p["bn_out_rm"], dtype=dt), mu3.reshape(ngroups, -1), momentum)
        new_running["bn_out_rv"] = _ema(np.asarray(p["bn_out_rv"], dtype=dt),
                                        var3.reshape(ngroups, -1) * (cnt / (cnt - 1)), momentum)
    else:
        mu3 = np.asarray(p["bn_out_rm"], dtype=dt)[None, None, :, None]
        var3 = np.asarray(p["bn_out_rv"], dtype=dt)[None, None, :, None]
        On = (Og - mu3) / np.sqrt(var3 + eps) * g3 + b3
    out_s = On.reshape(B, c, no, L).sum(axis=2)                # add sv+sve (:179)
    out = _from_strips(out_s, n, axis)
    if stride > 1:                                             # AvgPool2d(stride) (:186-187)
        nn_, cc, hh, ww = out.shape
        out = out.reshape(nn_, cc, hh // stride, stride, ww // stride, stride).mean(axis=(3, 5))

    cache = dict(X=X, Wq=Wq, yg=yg, mu1=mu1, var1=var1, q=q, k=k, v=v, S=S, mu2=mu2, var2=var2,
                 P=P, O=O, mu3=mu3, var3=var3, shape=(n, c, h, w), axis=axis, stride=stride,
                 variant=variant, ngroups=ngroups, eps=eps, training=training,
                 gates=(f_qr, f_kr, f_sv, f_sve), p=p)
    return out, cache, new_running


# --------------------------------------------------------------------------- backward
def _bn_bwd(dy, x, mean, var, gamma, red_axes, eps, training):
    """BN backward. Returns dx, dgamma, dbeta (summed over stat groups)."""
    inv = 1.0 / np.sqrt(var + eps)
    xhat = (x - mean) * inv
    all_axes = (0,) + tuple(red_axes)
    dgamma = (dy * xhat).sum(axis=all_axes)
    dbeta = dy.sum(axis=all_axes)
    if training:
        m1 = dy.mean(axis=red_axes, keepdims=True)
        m2 = (dy * xhat).mean(axis=red_axes, keepdims=True)
        dx = gamma * inv * (dy - m1 - xhat * m2)
    else:
        dx = gamma * inv * dy
    return dx, dgamma, dbeta


def axial_backward(dout, cache):
    """Gradients of one axial op.  Returns dict: dx, dw_qkv, drelative, dgates (4,),
    dbn_qkv_w/b, dbn_sim_w/b, dbn_out_w/b."""
    c_ = cache
    n, c, h, w = c_["shape"]
    axis, stride, variant, ngroups, eps, training = (c_[k] for k in
                                                     ("axis", "stride", "variant", "ngroups", "eps", "training"))
    p = c_["p"]
    dt = c_["X"].dtype
    G = GROUPS
    gp = c // G
    d = gp // 2
    L = h if axis == 0 else w
    other = w if axis == 0 else h
    B = n * other
    Bg = B // ngroups
    pos = variant != "nopos"
    nt = 3 if pos else 1
    no = 2 if pos else 1
    f_qr, f_kr, f_sv, f_sve = c_["gates"]
    dout = np.asarray(dout, dtype=dt)

    if stride > 1:                                             # AvgPool backward
        dout = np.repeat(np.repeat(dout, stride, axis=2), stride, axis=3) / (stride * stride)
    g_s = _to_strips(dout, axis)                               # (B, C, L)
    dOn = np.repeat(g_s[:, :, None, :], no, axis=2).reshape(ngroups, Bg, c * no, L)

    g3 = np.asarray(p["bn_out_w"], dtype=dt)[None, None, :, None]
    Og = c_["O"].reshape(ngroups, Bg, c * no, L)
    dO, dg3, db3 = _bn_bwd(dOn, Og, c_["mu3"], c_["var3"], g3, (1, 3), eps, training)
    dO = dO.reshape(B, G, gp, no, L)
    dO_sv = dO[:, :, :, 0]                                     # grad wrt f_sv * sv
    P, q, k, v = c_["P"], c_["q"], c_["k"], c_["v"]
    dgates = np.zeros(4, dtype=dt)
    # sv = f_sv * (P v)
    dgates[2] = (dO_sv * np.einsum("bgij,bgcj->bgci", P, v)).sum()
    dsv = dO_sv * f_sv
    dP = np.einsum("bgci,bgcj->bgij", dsv, v)
    dv = np.einsum("bgij,bgci->bgcj", P, dsv)
    drel = None
    if pos:
        rel = np.asarray(p["relative"], dtype=dt)
        E = toeplitz(rel, L)
        Eq, Ek, Ev = E[:d], E[d:2 * d], E[gp:]
        dO_sve = dO[:, :, :, 1]
        sve_raw = np.einsum("bgij,cij->bgci", P, Ev)
        dgates[3] = (dO_sve * sve_raw).sum()
        dsve = dO_sve * f_sve
        dP = dP + np.einsum("bgci,cij->bgij", dsve, Ev)
        dEv = np.einsum("bgij,bgci->cij", P, dsve)
    dZ = P * (dP - (P * dP).sum(axis=3, keepdims=True))        # softmax backward

    # BN2d backward: every normalised term receives the same dZ (they are summed)
    g2 = np.asarray(p["bn_sim_w"], dtype=dt).reshape(1, 1, nt, G, 1, 1)
    Sg = c_["S"].reshape(ngroups, Bg, nt, G, L, L)
    dSn = np.repeat(dZ[:, None], nt, axis=1).reshape(ngroups, Bg, nt, G, L, L)
    dS, dg2, db2 = _bn_bwd(dSn, Sg, c_["mu2"], c_["var2"], g2, (1, 4, 5), eps, training)
    dS = dS.reshape(B, nt, G, L, L)
    dS_qk = dS[:, 0]
    dq = np.einsum("bgij,bgcj->bgci", dS_qk, k)
    dk = np.einsum("bgij,bgci->bgcj", dS_qk, q)
    if pos:
        qr_raw = np.einsum("bgci,cij->bgij", q, Eq)
        kr_raw = np.einsum("bgcj,cji->bgij", k, Ek)
        dgates[0] = (dS[:, 1] * qr_raw).sum()
        dgates[1] = (dS[:, 2] * kr_raw).sum()
        dqr = dS[:, 1] * f_qr
        dkr = dS[:, 2] * f_kr
        dq = dq + np.einsum("bgij,cij->bgci", dqr, Eq)
        dk = dk + np.einsum("bgij,cji->bgcj", dkr, Ek)
        dEq = np.einsum("bgij,bgci->cij", dqr, q)
        dEk = np.einsum("bgij,bgcj->cji", dkr, k)
        dE = np.concatenate([dEq, dEk, dEv], axis=0)           # (2gp, L, L)
        drel = np.zeros_like(rel)
        ii = np.arange(L)[:, None]
        jj = np.arange(L)[None, :]
        idx = (ii - jj + L - 1).ravel()
        for ch in range(2 * gp):                               # index_add along diagonals
            np.add.at(drel[ch], idx, dE[ch].ravel())
    dyn = np.concatenate([dq, dk, dv], axis=2).reshape(ngroups, Bg, 2 * c, L)

    g1 = np.asarray(p["bn_qkv_w"], dtype=dt)[None, None, :, None]
    dy, dg1, db1 = _bn_bwd(dyn, c_["yg"], c_["mu1"], c_["var1"], g1, (1, 3), eps, training)
    dy = dy.reshape(B, 2 * c, L)
    dW = np.einsum("bol,bcl->oc", dy, c_["X"])
    dX = np.einsum("oc,bol->bcl", c_["Wq"], dy)
    dx = _from_strips(dX, n, axis)
    return dict(dx=dx, dw_qkv=dW, drelative=drel, dgates=dgates,
                dbn_qkv_w=dg1.reshape(-1), dbn_qkv_b=db1.reshape(-1),
                dbn_sim_w=dg2.reshape(-1), dbn_sim_b=db2.reshape(-1),
                dbn_out_w=dg3.reshape(-1), dbn_out_b=db3.reshape(-1))
