"""Prints the worst relative errors of one injected-draw update (Ant-shaped, full size) against the CPU oracle for
both GEMM engines - the numbers quoted in DESIGN.md section 4."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.sac_eo_oracle import NetCfg
from sac_expert_b200 import lib as L
from tests.helpers import build, compare_update

for shape, (S, A) in (("hopper", (11, 3)), ("ant", (27, 8))):
    for mode, name in ((L.GEMM_FP32_SIMT, "fp32-simt"), (L.GEMM_TCGEN05_BF16X3, "tcgen05-bf16x3-fused")):
        cfg = NetCfg(S=S, A=A)
        worst = {}
        for seed in (5, 31, 77):
            pop, probs = build(cfg, n_agents=2, B=256, E=20, N=2000, seed=seed, gemm_mode=mode)
            w = compare_update(pop, cfg, probs)
            for k, v in w.items():
                worst[k] = max(worst.get(k, 0), v)
            pop.close()
        keys = ["y", "L_q1", "L_pi", "mse", "p_loss", "alpha_loss", "g_q1", "g_q2", "g_actor", "g_alpha", "dtheta_q1", "dtheta_actor",
                "dtheta_t1", "adam_m_actor", "adam_v_actor", "oracle32_vs_64_g_actor"]
        print(shape, name, " ".join(f"{k}={worst[k]:.1e}" for k in keys))
