"""Generates the committed fixtures under tests/golden/.  Run in the BUILD container only:

    python tests/golden/make_golden.py

It reads the two weight-bearing pickles the reference ships (the only data artefacts it has; it has no
tests or golden vectors) and stores (a) the real trained Pendulum actor + running-normaliser statistics of
/root/reference/sac_eo/logs/TEMPLOG_0 as inputs, and (b) outputs of the fp64 CPU oracle on those inputs.
The reference itself (TensorFlow eager) cannot be executed here, so (b) pins the ORACLE, not the reference:
the pins against the reference's own code are make_golden_reference.py's.  /root/reference does not exist on the GPU box, hence
the fixtures are committed.
"""
import os
import pickle
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.sac_eo_oracle import (NetCfg, cg, draw_batch, flat, make_F, make_problem, sac_eo_update,  # noqa: E402
                                  to_torch_state)

OUT = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/sac_eo/logs/TEMPLOG_0"


def main():
    d = pickle.load(open(REF, "rb"))
    aw = [np.asarray(w, np.float32) for w in d["final"]["actor_weights"]]      # [3,64],[64],[64,64],[64],[64,1],[1],[1,1]
    rms = d["final"]["rms_stats"]
    std = lambda k: np.sqrt(np.asarray(rms[k]["var"], np.float32))            # RunningNormalizer.instantiate, normalizer.py:111-122
    stats = dict(s_mean=np.asarray(rms["s_rms"]["mean"], np.float32), s_std=std("s_rms"),
                 a_mean=np.asarray(rms["a_rms"]["mean"], np.float32).reshape(1), a_std=std("a_rms").reshape(1),
                 ret_std=np.float32(std("ret_rms")),
                 d_mean=np.asarray(rms["delta_rms"]["mean"], np.float32), d_std=std("delta_rms"))
    np.savez(os.path.join(OUT, "pendulum_templog0.npz"), **{f"actor_{i}": w for i, w in enumerate(aw)}, **stats)

    # ---- one plain-SAC and one SAC-EO update around the real actor / normalisers --------------
    for tag, nm in (("sac", 0), ("saceo", 2)):
        cfg = NetCfg(S=3, A=1, actor_hidden=(64, 64), critic_hidden=(64, 64), model_hidden=(64, 64),
                     actor_acts=("tanh", "tanh"), critic_acts=("tanh", "tanh"), per_state_std=False, num_models=nm)
        st, replay, expert, hyper = make_problem(cfg, B=32, E=8, N=300, seed=123, perturb=0.02)
        st["actor"] = [w.copy() for w in aw]
        st["adam_actor"] = dict(m=[np.zeros_like(w) for w in aw], v=[np.zeros_like(w) for w in aw], t=0)
        st.update(s_mean=stats["s_mean"], s_std=stats["s_std"], a_mean=stats["a_mean"], a_std=stats["a_std"],
                  ret_std=stats["ret_std"], m_s_mean=stats["s_mean"], m_s_std=stats["s_std"],
                  m_a_mean=stats["a_mean"], m_a_std=stats["a_std"], m_d_mean=stats["d_mean"], m_d_std=stats["d_std"])
        hyper["eps"] = 0.25
        batch = draw_batch(cfg, replay, expert, 32, seed=321)
        o = sac_eo_update(cfg, to_torch_state(st, torch.float64), batch, hyper)
        out = dict(y=o["y"].numpy(), L_q1=o["L_q1"].numpy(), L_q2=o["L_q2"].numpy(), L_pi=o["L_pi"].numpy(),
                   mse=o["mse"].numpy(), p_loss=o["p_loss"].numpy(), alpha_loss=o["alpha_loss"].numpy(),
                   g_alpha=o["g_alpha"].numpy(), alpha_new=o["new"]["alpha"].numpy(),
                   g_q1=flat(o["g_q1"]).numpy(), g_q2=flat(o["g_q2"]).numpy(), g_actor=flat(o["g_actor"]).numpy(),
                   new_actor=flat(o["new"]["actor"]).numpy(), new_q1=flat(o["new"]["q1"]).numpy(),
                   new_t1=flat(o["new"]["t1"]).numpy())
        np.savez(os.path.join(OUT, f"golden_pendulum_{tag}.npz"), **out)

    # ---- Fisher-vector product and CG on the real actor (TRPO path of the pickled run) --------
    cfg = NetCfg(S=3, A=1, actor_hidden=(64, 64), critic_hidden=(8, 8), actor_acts=("tanh", "tanh"),
                 per_state_std=False, num_models=0)
    rng = np.random.default_rng(7)
    states = (rng.standard_normal((100, 3)) * stats["s_std"] + stats["s_mean"]).astype(np.float32)
    st = dict(actor=aw, s_mean=stats["s_mean"], s_std=stats["s_std"])
    th = to_torch_state(st, torch.float64)
    F = make_F(cfg, th["actor"], states, th, damp=0.01)
    nA = sum(w.size for w in aw)
    x = rng.standard_normal(nA)
    b = rng.standard_normal(nA) * 0.05
    sol = cg(F, torch.from_numpy(b), cg_iters=20)
    np.savez(os.path.join(OUT, "golden_pendulum_fvp.npz"), states=states, x=x, b=b, Fx=F(torch.from_numpy(x)).numpy(),
             cg_x=sol.numpy(), vFv=float(sol.dot(F(sol))))
    print("wrote fixtures to", OUT)


if __name__ == "__main__":
    main()
