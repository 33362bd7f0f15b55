"""Hand-derived forward/backward of one SAC-EO update in NumPy.  TEST INFRASTRUCTURE ONLY.

Same contract as ``oracle/sac_eo_oracle.py`` (see its header for the parity status; checker only).
This twin uses NO autograd: every gradient is the explicit chain the CUDA kernels implement
(SURVEY.md App. A "analytic backward"), phase by phase and GEMM by GEMM, so that the kernel
sequence can be validated on the CPU against ``sac_eo_update`` (autograd) before any GPU time
is spent.  Reference citations are on the autograd twin; this file cites the twin.
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence

import numpy as np

from .sac_eo_oracle import (ADAM_B1, ADAM_B2, ADAM_EPS, LOG2, LOG2PI, MAX_LOG_STD, MIN_LOG_STD,
                            NetCfg)


def _act(name, z):
    if name == "relu":
        return np.maximum(z, 0)
    if name == "tanh":
        return np.tanh(z)
    if name == "elu":
        return np.where(z > 0, z, np.expm1(np.minimum(z, 0)))
    raise ValueError(name)


def _dact_from_out(name, h):
    """Activation derivative expressed from the POST-activation value (what the kernels keep)."""
    if name == "relu":
        return (h > 0).astype(h.dtype)
    if name == "tanh":
        return 1 - h * h
    if name == "elu":
        return np.where(h > 0, 1.0, h + 1.0).astype(h.dtype)
    raise ValueError(name)


def _softplus(x):
    return np.logaddexp(0, x)


def mlp_fwd(theta, x, acts):
    h1 = _act(acts[0], x @ theta[0] + theta[1])
    h2 = _act(acts[1], h1 @ theta[2] + theta[3])
    out = h2 @ theta[4] + theta[5]
    return h1, h2, out


def mlp_bwd(theta, x, h1, h2, dout, acts, need_dx=False, need_dw=True, out_cols=None):
    """Backward of mlp_fwd.  ``out_cols`` restricts the last layer to the first columns
    (model: reward column dropped)."""
    W2 = theta[4] if out_cols is None else theta[4][:, :out_cols]
    g: List = [None] * 6
    if need_dw:
        g[4] = h2.T @ dout
        g[5] = dout.sum(0)
    dh2 = (dout @ W2.T) * _dact_from_out(acts[1], h2)
    if need_dw:
        g[2] = h1.T @ dh2
        g[3] = dh2.sum(0)
    dh1 = (dh2 @ theta[2].T) * _dact_from_out(acts[0], h1)
    if need_dw:
        g[0] = x.T @ dh1
        g[1] = dh1.sum(0)
    dx = dh1 @ theta[0].T if need_dx else None
    return g, dx


def norm(x, mean, std):
    return (x - mean) / np.maximum(std, 1e-8)


def head_fwd(cfg: NetCfg, out, logstd_var, u, act_limit, deterministic=False):
    A = cfg.A
    if cfg.per_state_std:
        mean, ls_raw = out[:, :A], out[:, A:]
    else:
        mean, ls_raw = out, np.broadcast_to(logstd_var, out.shape)
    ls = np.clip(ls_raw, MIN_LOG_STD, MAX_LOG_STD)
    std = np.exp(ls)
    z = mean if deterministic else mean + std * u
    nlp = 0.5 * (((z - mean) / std) ** 2 + 2 * ls + LOG2PI).sum(-1) \
        + (2.0 * (LOG2 - z - _softplus(-2.0 * z))).sum(-1)
    t = np.tanh(z)
    cache = dict(ls_raw=ls_raw, std=std, t=t, u=u, det=deterministic)
    return act_limit * t, nlp, cache


def head_bwd(cfg: NetCfg, cache, g_pi, g_nlp, act_limit):
    """g_pi [R,A] = dL/d(pi), g_nlp [R] = dL/d(neglogp) -> dL/d(out) (+ dL/d(logstd_var))."""
    t, std, u = cache["t"], cache["std"], cache["u"]
    dz = g_pi * act_limit * (1 - t * t) + g_nlp[:, None] * (-2.0 * t)
    dmean = dz
    mask = ((cache["ls_raw"] >= MIN_LOG_STD) & (cache["ls_raw"] <= MAX_LOG_STD)).astype(t.dtype)
    if cache["det"]:
        dls = g_nlp[:, None] * np.ones_like(t) * mask
    else:
        dls = (dz * std * u + g_nlp[:, None]) * mask
    if cfg.per_state_std:
        return np.concatenate([dmean, dls], -1), None
    return dmean, dls.sum(0, keepdims=True)


def adam(theta, g, m, v, t, lr):
    t = t + 1
    dt = theta.dtype
    lr_t = dt.type(lr * math.sqrt(1.0 - ADAM_B2 ** t) / (1.0 - ADAM_B1 ** t))
    m = dt.type(ADAM_B1) * m + dt.type(1.0 - ADAM_B1) * g
    v = dt.type(ADAM_B2) * v + dt.type(1.0 - ADAM_B2) * g * g
    theta = theta - lr_t * m / (np.sqrt(v) + dt.type(ADAM_EPS))
    return theta, m, v, t


def analytic_update(cfg: NetCfg, st: Dict, batch: Dict, hyper: Dict, dtype=np.float32) -> Dict:
    """Explicit-gradient twin of ``sac_eo_update``; state/batch are NumPy (``make_problem`` /
    ``draw_batch``).  Returns the same keys (NumPy)."""
    f = lambda x: np.asarray(x, dtype=dtype)
    L = lambda ws: [f(w) for w in ws]
    S, A = cfg.S, cfg.A
    s, a, sp, r = f(batch["s"]), f(batch["a"]), f(batch["sp"]), f(batch["r"])
    omd = f(1.0 - np.asarray(batch["d"], np.float64))
    B = s.shape[0]
    alpha = f(st["alpha"])
    al = f(st["act_limit"])
    pi_th = L(st["actor"])
    lsv = pi_th[6] if not cfg.per_state_std else None
    sm, ss, am, as_ = f(st["s_mean"]), f(st["s_std"]), f(st["a_mean"]), f(st["a_std"])
    ret = np.maximum(f(st["ret_std"]), dtype(1e-8))
    out: Dict = {}

    # 1. target
    _, _, o = mlp_fwd(pi_th, norm(sp, sm, ss), cfg.actor_acts)
    a1, nlp1, _ = head_fwd(cfg, o, lsv, f(batch["u1"]), al)
    xc = np.concatenate([norm(sp, sm, ss), norm(a1, am, as_)], -1)
    qt = [mlp_fwd(L(st[k]), xc, cfg.critic_acts)[2][:, 0] * ret for k in ("t1", "t2")]
    y = r + dtype(hyper["gamma"]) * (omd * (np.minimum(qt[0], qt[1]) + alpha * nlp1))
    out["y"] = y

    # 2. critics
    new: Dict = {}
    xc = np.concatenate([norm(s, sm, ss), norm(a, am, as_)], -1)
    for k in ("q1", "q2"):
        th = L(st[k])
        h1, h2, q = mlp_fwd(th, xc, cfg.critic_acts)
        diff = q[:, 0] - y
        out["L_" + k] = (0.5 * diff * diff).mean()
        g, _ = mlp_bwd(th, xc, h1, h2, (diff / B)[:, None], cfg.critic_acts)
        out["g_" + k] = g
        ad = st["adam_" + k]
        res = [adam(p, gi, f(mi), f(vi), ad["t"], hyper["lr_q"]) for p, gi, mi, vi in
               zip(th, g, ad["m"], ad["v"])]
        new[k] = [x[0] for x in res]
        new["adam_" + k] = dict(m=[x[1] for x in res], v=[x[2] for x in res], t=ad["t"] + 1)

    # 3. actor: rows [0,B) = Lpi on s ; rows [B,B+E) = expert term
    nm = cfg.num_models
    x0 = norm(s, sm, ss)
    if nm > 0:
        sE, spE = f(batch["sE"]), f(batch["spE"])
        parts = [np.asarray(batch["I1"])] + ([np.asarray(batch["I2"])] if nm == 2 else [])
        us = [f(batch["u3"])] + ([f(batch["u4"])] if nm == 2 else [])
        x0 = np.concatenate([x0] + [norm(sE[I], sm, ss) for I in parts], 0)
    h1, h2, o = mlp_fwd(pi_th, x0, cfg.actor_acts)
    a2, nlp2, c_main = head_fwd(cfg, o[:B], lsv, f(batch["u2"]), al)
    xq = np.concatenate([norm(s, sm, ss), norm(a2, am, as_)], -1)
    qs, caches = [], []
    for k in ("q1", "q2"):
        ch1, ch2, q = mlp_fwd(new[k], xq, cfg.critic_acts)
        qs.append(q[:, 0])
        caches.append((ch1, ch2))
    minq = np.minimum(qs[0], qs[1])
    l_pi = (-alpha * nlp2 - minq).mean()
    out["L_pi"] = l_pi
    eps = dtype(hyper["eps"]) if nm > 0 else dtype(0)
    w_pi = (dtype(1) - eps) if nm > 0 else dtype(1)
    # d(-minq)/dq_k: ties split equally (tf.reduce_min)
    sel0 = np.where(qs[0] < qs[1], 1.0, np.where(qs[0] == qs[1], 0.5, 0.0)).astype(dtype)
    da2 = np.zeros_like(a2)
    for k, key in enumerate(("q1", "q2")):
        sel = sel0 if k == 0 else (1 - sel0)
        dq = (-w_pi / B) * sel
        _, dx = mlp_bwd(new[key], xq, caches[k][0], caches[k][1], dq[:, None], cfg.critic_acts,
                        need_dx=True, need_dw=False)
        da2 += dx[:, S:] / np.maximum(as_, dtype(1e-8))
    g_nlp = np.full(B, -alpha * w_pi / B, dtype)
    dout_main, dlsv = head_bwd(cfg, c_main, da2, g_nlp, al)
    douts = [dout_main]
    mse = dtype(0)
    if nm > 0:
        msm, mss = f(st["m_s_mean"]), f(st["m_s_std"])
        mam, mas = f(st["m_a_mean"]), f(st["m_a_std"])
        mdm, mds = f(st["m_d_mean"]), np.maximum(f(st["m_d_std"]), dtype(1e-8))
        row = B
        nrow_mean = len(parts[0])          # mean over E/2 rows (or E rows for one model)
        for k, I in enumerate(parts):
            n = len(I)
            c, _, c_cache = head_fwd(cfg, o[row:row + n], lsv, us[k], al)
            th_m = L(st["m%d" % (k + 1)])
            xm = np.concatenate([norm(sE[I], msm, mss), norm(c, mam, mas)], -1)
            mh1, mh2, mo = mlp_fwd(th_m, xm, cfg.model_acts)
            delta = mo[:, :S]
            cmask = np.ones_like(delta)
            if cfg.delta_clip_pred:
                cmask = ((delta >= -cfg.delta_clip_pred) & (delta <= cfg.delta_clip_pred)).astype(dtype)
                delta = np.clip(delta, -cfg.delta_clip_pred, cfg.delta_clip_pred)
            pred = sE[I] + delta * mds + mdm
            err = spE[I] - pred
            mse = mse + (0.5 * (err * err).sum(-1)).sum() / nrow_mean
            dpred = -err / nrow_mean * eps
            ddelta = dpred * mds * cmask
            _, dxm = mlp_bwd(th_m, xm, mh1, mh2, ddelta, cfg.model_acts, need_dx=True,
                             need_dw=False, out_cols=S)
            dc = dxm[:, S:] / np.maximum(mas, dtype(1e-8))
            d_o, dl = head_bwd(cfg, c_cache, dc, np.zeros(n, dtype), al)
            douts.append(d_o)
            if dlsv is not None:
                dlsv = dlsv + dl
            row += n
    out["mse"] = mse
    out["p_loss"] = w_pi * l_pi + eps * mse if nm > 0 else l_pi
    dout = np.concatenate(douts, 0)
    g, _ = mlp_bwd(pi_th[:6], x0, h1, h2, dout, cfg.actor_acts)
    if not cfg.per_state_std:
        g = g + [dlsv]
    out["g_actor"] = g
    ad = st["adam_actor"]
    res = [adam(p, gi, f(mi), f(vi), ad["t"], hyper["lr_pi"]) for p, gi, mi, vi in
           zip(pi_th, g, ad["m"], ad["v"])]
    new["actor"] = [x[0] for x in res]
    new["adam_actor"] = dict(m=[x[1] for x in res], v=[x[2] for x in res], t=ad["t"] + 1)

    # 4. temperature
    _, _, o = mlp_fwd(new["actor"], norm(s, sm, ss), cfg.actor_acts)
    lsv2 = new["actor"][6] if not cfg.per_state_std else None
    _, nlp3, _ = head_fwd(cfg, o, lsv2, f(batch["u5"]), al)
    te = dtype(hyper["target_entropy"])
    mean_term = (-nlp3 + te).mean()
    out["alpha_loss"] = -alpha * mean_term
    out["g_alpha"] = -mean_term
    ad = st["adam_alpha"]
    al_new, am_, av_, at_ = adam(alpha, out["g_alpha"], f(ad["m"]), f(ad["v"]), ad["t"], hyper["lr_alpha"])
    new["alpha"] = np.maximum(al_new, dtype(1e-5))
    new["adam_alpha"] = dict(m=am_, v=av_, t=at_)

    # 5. Polyak
    tau = dtype(hyper["tau"])
    for tk, qk in (("t1", "q1"), ("t2", "q2")):
        if hyper.get("do_polyak", True):
            new[tk] = [f(tg) * dtype(1.0 - hyper["tau"]) + lv * tau for tg, lv in zip(st[tk], new[qk])]
        else:
            new[tk] = L(st[tk])
    out["new"] = new
    return out


# ----------------------------------------------------------------------------------
# Fisher-vector product in JVP -> metric -> VJP form (what the kernels do)
# ----------------------------------------------------------------------------------
def fvp_gn(cfg: NetCfg, theta: Sequence[np.ndarray], x_flat: np.ndarray, s_all, st: Dict,
           damp: float, dtype=np.float32) -> np.ndarray:
    f = lambda v: np.asarray(v, dtype=dtype)
    th = [f(w) for w in theta]
    A = cfg.A
    x_flat = f(x_flat)
    tang, o = [], 0
    for w in th:
        tang.append(x_flat[o:o + w.size].reshape(w.shape))
        o += w.size
    X = norm(f(s_all), f(st["s_mean"]), f(st["s_std"]))
    N = X.shape[0]
    acts = cfg.actor_acts
    h1, h2, out = mlp_fwd(th, X, acts)
    d1 = _dact_from_out(acts[0], h1) * (X @ tang[0] + tang[1])
    d2 = _dact_from_out(acts[1], h2) * (d1 @ th[2] + h1 @ tang[2] + tang[3])
    dout = d2 @ th[4] + h2 @ tang[4] + tang[5]
    floor = math.log(1e-3)
    if cfg.per_state_std:
        o2 = out[:, A:]
        sp_ = _softplus(o2)
        ls = np.log(sp_) + dtype(math.log(cfg.std_mult) - math.log(math.log(2.0)))
        dls_do2 = (1.0 / (1.0 + np.exp(-o2))) / sp_
        mask = (ls >= floor).astype(dtype)
        ls = np.maximum(ls, dtype(floor))
        jm, jl = dout[:, :A], dout[:, A:] * dls_do2 * mask
        gm = np.exp(-2 * ls) * jm / N
        gl = 2.0 * jl / N
        g_out = np.concatenate([gm, gl * dls_do2 * mask], -1)
        g_lsv = None
    else:
        ls = np.broadcast_to(th[6] + dtype(math.log(cfg.std_mult)), (N, A))
        mask = (ls >= floor).astype(dtype)
        ls = np.maximum(ls, dtype(floor))
        jm, jl = dout, np.broadcast_to(tang[6], (N, A)) * mask
        g_out = np.exp(-2 * ls) * jm / N
        g_lsv = (2.0 * jl / N * mask).sum(0, keepdims=True)
    g, _ = mlp_bwd(th[:6], X, h1, h2, f(g_out), acts)
    if g_lsv is not None:
        g = g + [f(g_lsv)]
    return np.concatenate([gi.reshape(-1) for gi in g]).astype(dtype) + dtype(damp) * x_flat
