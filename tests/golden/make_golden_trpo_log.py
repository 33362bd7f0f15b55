"""Extracts the TRPO training log the reference itself recorded in /root/reference/sac_eo/logs/TEMPLOG_0 (a real run
of its TRPO update on Pendulum-v1: alg_type 'mbrl', mf_algo 'trpo', two updates) into a small committed fixture.
Run in the BUILD container only:    python tests/golden/make_golden_trpo_log.py

Unlike the other fixtures (oracle outputs on the reference's weights), these numbers are OUTPUTS OF THE REFERENCE:
the per-update `ent`, `tv_pre`, `kl_pre`, `tv`, `kl`, `adj`, `improve`, `alpha` that `TRPO._backtrack` returned
(trpo.py:303-313) and the hyper-parameters they were produced with.  They pin the pieces of the restatement they
determine: the GaussianActor entropy formula and its state-independent logstd parameterisation, and the line-search
accept / shrink rule."""
import json
import os
import pickle

import numpy as np

OUT = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/sac_eo/logs/TEMPLOG_0"


def main():
    d = pickle.load(open(REF, "rb"))
    tr, fin = d["train"], d["final"]
    mf, ak = d["param"]["mf_update_kwargs"], d["param"]["actor_kwargs"]
    keys = ("ent", "tv_pre", "kl_pre", "tv", "kl", "adj", "improve", "alpha")
    out = {k: [float(np.float64(v)) for v in np.asarray(tr[k])] for k in keys}
    out["dtype"] = {k: str(np.asarray(tr[k]).dtype) for k in keys}
    out["hyper"] = {k: mf[k] for k in ("delta_trpo", "cg_it", "trust_sub", "trust_damp", "kl_maxfactor", "ent_reg",
                                        "adv_center", "adv_scale")}
    out["actor"] = {"std_mult": ak["actor_std_mult"], "per_state_std": ak["actor_per_state_std"],
                    "a_dim": int(np.asarray(fin["actor_weights"][-1]).size),
                    "final_logstd": [float(x) for x in np.asarray(fin["actor_weights"][-1]).ravel()]}
    json.dump(out, open(os.path.join(OUT, "templog0_trpo_log.json"), "w"), indent=1)
    print(out)


if __name__ == "__main__":
    main()
