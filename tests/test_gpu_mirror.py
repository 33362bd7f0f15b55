"""GPU: the reference-compatible class interface (init_actor / init_critics / init_world_models / init_alg,
alg._update) against the oracle, with the NumPy RNG streams consumed exactly like the reference does."""
import numpy as np
import pytest
import torch

from oracle.sac_eo_oracle import NetCfg, sac_eo_update, to_torch_state
from sac_expert_b200 import lib as L
from sac_expert_b200.sac_eo.actors.init_actor import init_actor
from sac_expert_b200.sac_eo.algs.init_alg import init_alg
from sac_expert_b200.sac_eo.common.update_utils import cg, make_F
from sac_expert_b200.sac_eo.critics.init_critic import init_critics
from sac_expert_b200.sac_eo.envs.synthetic import SyntheticEnv
from sac_expert_b200.sac_eo.models.init_world_models import init_world_models
from tests.helpers import rel

pytestmark = pytest.mark.gpu

SETUP = dict(separate_reward_nn=False, reward_loss_coef=1.0, scale_model_loss=False, delta_clip_loss=None,
             reward_clip_loss=None, delta_clip_pred=None, reward_clip_pred=None)


def build_alg(alg_type, S=11, A=3, B=32, E=8, per_state_std=True):
    np.random.seed(0)
    env = SyntheticEnv(S, A)
    actor = init_actor(env, [32, 32], ["relu"], 0.01, 1.0, "orthogonal", False, None, per_state_std, True, False)
    critics, q_targets, q_critics = init_critics(env, [32, 32], ["relu"], 1.0, None, 2, False, "orthogonal", False)
    models = init_world_models(env, [48, 48], ["relu"], 0.01, 1.0, None, [48, 48], ["relu"], 0.01, None, 2, False, SETUP)
    kw = dict(alg_type=alg_type, sac_batch_size=B, expert_buffer_size=E, gamma=0.99, alg_seed=5, epsilon=0.2,
              device_replay_capacity=1000, gemm_mode=L.GEMM_FP32_SIMT)
    alg = init_alg(0, env, env, env, actor, critics, q_targets, q_critics, models, kw, {}, None, None)
    rng = np.random.default_rng(1)
    n = 300
    alg.env_data.add(rng.standard_normal((n, S)).astype(np.float32), rng.uniform(-1, 1, (n, A)).astype(np.float32),
                     rng.standard_normal(n).astype(np.float32), rng.standard_normal((n, S)).astype(np.float32),
                     rng.random(n) < 0.05)
    return alg, rng


def snapshot(alg, cfg):
    st = dict(actor=alg.actor.get_weights(), q1=alg.q_critics[0].get_weights(), q2=alg.q_critics[1].get_weights(),
              t1=alg.q_targets[0].get_weights(), t2=alg.q_targets[1].get_weights(), alpha=np.float32(alg.alpha))
    if cfg.num_models:
        st["m1"], st["m2"] = alg.models[0].get_weights(), alg.models[1].get_weights()
    for k in ("q1", "q2", "actor"):
        st["adam_" + k] = dict(m=[np.zeros_like(w) for w in st[k]], v=[np.zeros_like(w) for w in st[k]], t=0)
    st["adam_alpha"] = dict(m=np.float32(0), v=np.float32(0), t=0)
    S, A = cfg.S, cfg.A
    z, o = (lambda n: np.zeros(n, np.float32)), (lambda n: np.ones(n, np.float32))
    st.update(s_mean=z(S), s_std=o(S), a_mean=z(A), a_std=o(A), ret_std=np.float32(1), m_s_mean=z(S), m_s_std=o(S),
              m_a_mean=z(A), m_a_std=o(A), m_d_mean=z(S), m_d_std=o(S), act_limit=o(A))
    return st


@pytest.mark.parametrize("alg_type", ["sac", "sac_imit"])
def test_alg_update_matches_oracle_with_reference_rng_order(alg_type):
    S, A, B, E = 11, 3, 32, 8
    alg, rng = build_alg(alg_type, S, A, B, E)
    cfg = NetCfg(S=S, A=A, actor_hidden=(32, 32), critic_hidden=(32, 32), model_hidden=(48, 48),
                 num_models=2 if alg_type == "sac_imit" else 0)
    st = snapshot(alg, cfg)
    hyper = dict(gamma=0.99, tau=5e-3, lr_q=3e-4, lr_pi=1e-4, lr_alpha=1e-4, eps=0.2, target_entropy=float(-A),
                 do_polyak=True)
    expert_reg = None
    if alg_type == "sac_imit":
        alg.expert_data.add(rng.standard_normal((E, S)).astype(np.float32), rng.uniform(-1, 1, (E, A)).astype(np.float32),
                            rng.standard_normal(E).astype(np.float32), rng.standard_normal((E, S)).astype(np.float32),
                            np.zeros(E, bool))
        expert_reg = alg._expert_preprocess()
        assert expert_reg[3] == 0.2 and expert_reg[4] is False
    # ---- device update, global RNG seeded
    np.random.seed(123)
    if expert_reg is None:
        alg._update(0)
    else:
        alg._update(0, expert_reg)
    # ---- replay the reference's draw order for the oracle
    np.random.seed(123)
    idx = np.random.randint(alg.env_data.current_size, size=B)
    assert np.array_equal(idx, alg.last_idx)
    u1 = np.random.normal(size=(B, A))
    batch = dict(idx=idx, s=alg.env_data.s_all[idx], a=alg.env_data.a_all[idx], sp=alg.env_data.sp_all[idx],
                 r=alg.env_data.r_all[idx], d=alg.env_data.d_all[idx], u1=u1)
    if expert_reg is not None:
        order = np.arange(E)
        np.random.default_rng(5).shuffle(order)           # alg_seed=5 -> self.rng
        I1, I2 = np.array_split(order, 2)
        batch.update(I1=I1, I2=I2, sE=expert_reg[0], spE=expert_reg[2])
    batch["u2"] = np.random.normal(size=(B, A))
    if expert_reg is not None:
        batch["u3"], batch["u4"] = np.random.normal(size=(E // 2, A)), np.random.normal(size=(E // 2, A))
    batch["u5"] = np.random.normal(size=(B, A))
    o = sac_eo_update(cfg, to_torch_state(st), batch, hyper)
    assert abs(alg.last_losses["p_loss"] - float(o["p_loss"])) < 1e-4 * max(1, abs(float(o["p_loss"])))
    assert abs(alg.last_losses["alpha_loss"] - float(o["alpha_loss"])) < 1e-4 * max(1, abs(float(o["alpha_loss"])))
    for name, net in (("actor", alg.actor), ("q1", alg.q_critics[0]), ("q2", alg.q_critics[1]),
                      ("t1", alg.q_targets[0]), ("t2", alg.q_targets[1])):
        for got, new, old in zip(net.get_weights(), o["new"][name], st[name]):
            d_ref = new.numpy() - old
            if np.linalg.norm(d_ref) > 0:
                assert rel(got - old, d_ref) < 1e-3, name
    assert alg.alpha == pytest.approx(float(o["new"]["alpha"]), rel=1e-5)
    if expert_reg is not None:
        assert alg.logger.train_dict["epsilon"] == [0.2]


def test_actor_critic_model_forwards_match_oracle():
    from oracle.sac_eo_oracle import head, model_sample, q_forward, q_value
    S, A = 11, 3
    alg, rng = build_alg("sac_imit", S, A)
    cfg = NetCfg(S=S, A=A, actor_hidden=(32, 32), critic_hidden=(32, 32), model_hidden=(48, 48))
    st = to_torch_state(snapshot(alg, cfg))
    s = rng.standard_normal((5, S)).astype(np.float32)
    a = rng.uniform(-1, 1, (5, A)).astype(np.float32)
    np.random.seed(9)
    pi, nlp = alg.actor.evaluate(s)
    np.random.seed(9)
    u = np.random.normal(size=(5, A))
    pi_ref, nlp_ref = head(cfg, st["actor"], torch.from_numpy(s), torch.from_numpy(u), st)
    assert rel(pi.numpy(), pi_ref.numpy()) < 1e-5 and rel(nlp.numpy(), nlp_ref.numpy()) < 1e-5
    det = alg.actor.sample(s[0], deterministic=True)
    assert det.shape == (A,)
    assert rel(alg.q_critics[1]._forward(s, a).numpy(), q_forward(cfg, st["q2"], torch.from_numpy(s), torch.from_numpy(a), st).numpy()) < 1e-5
    assert rel(alg.q_targets[0].value(s, a).numpy(), q_value(cfg, st["t1"], torch.from_numpy(s), torch.from_numpy(a), st).numpy()) < 1e-5
    assert rel(alg.models[1].sample(s, a).numpy(), model_sample(cfg, st["m2"], torch.from_numpy(s), torch.from_numpy(a), st).numpy()) < 1e-5


def test_make_F_and_cg_interface_on_golden_pendulum_actor():
    """Reference-shaped call sequence of trpo.py:179-187 on the REAL trained Pendulum actor shipped in the
    reference's TEMPLOG_0 (tests/golden/pendulum_templog0.npz), against the committed fp64 oracle outputs."""
    import os
    g = os.path.join(os.path.dirname(__file__), "golden")
    pend, gold = np.load(os.path.join(g, "pendulum_templog0.npz")), np.load(os.path.join(g, "golden_pendulum_fvp.npz"))
    env = SyntheticEnv(3, 1)
    actor = init_actor(env, [64, 64], ["tanh"], 0.01, 1.0, "orthogonal", False, [pend[f"actor_{i}"] for i in range(7)],
                       False, True, False)

    class _N:   # normaliser carrying the pickled running statistics
        def get_rms(self):
            from sac_expert_b200.sac_eo.common.normalizer import RunningNormalizer
            r = RunningNormalizer(3); r.mean, r.std = pend["s_mean"], pend["s_std"]
            return r, None, None, None, None
    actor.set_rms(_N())
    F = make_F(actor, gold["states"], trust_sub=1, trust_damp=0.01)
    assert rel(F(gold["x"]).numpy(), gold["Fx"]) < 1e-3
    v = cg(F, gold["b"].astype(np.float32), cg_iters=20)
    vFv = float(np.dot(v, F(v).numpy()))
    assert rel(v, gold["cg_x"]) < 2e-2            # 20 fp32 CG iterations vs the fp64 oracle
    assert vFv == pytest.approx(float(gold["vFv"]), rel=2e-2)


@pytest.mark.parametrize("shuffle,holdout", [(True, 0.0), (False, 0.0), (True, 0.2)])
def test_update_models_matches_oracle_with_reference_rng_order(shuffle, holdout):
    """SAC_exp._update_models (SAC_expert.py:480-609): the index stream (holdout shuffle, per-epoch per-model
    shuffles, ragged tail dropped, model_max_updates cap) comes from the global NumPy RNG like the reference's,
    the gradient steps run on the device; the oracle replays the same stream step by step."""
    from oracle.sac_eo_oracle import apply_model_grads
    S, A, B, E = 11, 3, 32, 8
    alg, rng = build_alg("sac_imit", S, A, B, E)
    alg.alg_kwargs.update(model_num_epochs=2, model_batch_size=50, model_batch_shuffle=shuffle, model_lr=2e-3,
                          model_max_updates=9, model_max_grad_norm=0.5, model_holdout_ratio=holdout)
    d = alg.env_data
    alg.model_data.add(d.s_all, d.a_all, d.r_all, d.sp_all, d.d_all)
    alg.expert_data.add(rng.standard_normal((E, S)).astype(np.float32), rng.uniform(-1, 1, (E, A)).astype(np.float32),
                        rng.standard_normal(E).astype(np.float32), rng.standard_normal((E, S)).astype(np.float32),
                        np.zeros(E, bool))
    cfg = NetCfg(S=S, A=A, actor_hidden=(32, 32), critic_hidden=(32, 32), model_hidden=(48, 48))
    st = to_torch_state(snapshot(alg, cfg))
    np.random.seed(77)
    n_upd = alg._update_models()
    n = d.current_size
    n_train = int(n * (1 - holdout)) if holdout else n
    assert n_upd == min(9, 2 * (n_train // 50))
    # ---- replay with the oracle
    np.random.seed(77)
    rows = np.arange(n)
    if holdout:
        perm = np.arange(n)
        np.random.shuffle(perm)
        rows = perm[:n_train]
    models = [st["m1"], st["m2"]]
    adam = dict(m=[[torch.zeros_like(w) for w in m] for m in models], v=[[torch.zeros_like(w) for w in m] for m in models], t=0)
    done = 0
    for ep in range(2):
        if shuffle:
            order = []
            for _ in range(2):
                o_ = np.arange(n_train)
                np.random.shuffle(o_)
                order.append(o_)
        else:
            o_ = np.arange(n_train)
            np.random.shuffle(o_)
            order = [o_, o_]
        for k in range(n_train // 50):
            if done >= 9:
                break
            b = [{key: torch.as_tensor(getattr(d, key + "_all")[rows[order[m][50 * k:50 * (k + 1)]]]) for key in ("s", "a", "sp", "r")}
                 for m in range(2)]
            out = apply_model_grads(cfg, models, adam, b, st, dict(model_lr=2e-3, model_max_grad_norm=0.5))
            models, adam = out["models"], dict(m=out["m"], v=out["v"], t=out["t"])
            done += 1
    assert done == n_upd and int(alg.pop.t["model_t"][0]) == done
    for m in range(2):
        assert abs(float(alg.last_model_losses[m]) - float(out["losses"][m])) < 1e-3 * abs(float(out["losses"][m]))
        for got, ref, old in zip(alg.models[m].get_weights(), models[m], st["m%d" % (m + 1)]):
            assert rel(got - old.numpy(), ref.numpy() - old.numpy()) < 2e-3, m     # 9 compounded Adam steps
    assert len(alg.model_MSE_on_expert_data) == 1 and len(alg.model_MSE_on_expert_counterfactual_action) == 1


def test_update_models_gaussian_mirror():
    """GaussianModel through the class interface: logstd joins the device optimiser, get_weights() returns it,
    and one _apply_model_grads call matches the oracle's Gaussian NLL step."""
    from oracle.sac_eo_oracle import apply_model_grads
    S, A, B, E = 11, 3, 32, 8
    np.random.seed(0)
    env = SyntheticEnv(S, A)
    actor = init_actor(env, [32, 32], ["relu"], 0.01, 1.0, "orthogonal", False, None, True, True, False)
    critics, q_targets, q_critics = init_critics(env, [32, 32], ["relu"], 1.0, None, 2, False, "orthogonal", False)
    setup = dict(SETUP, scale_model_loss=True, reward_loss_coef=0.8)
    models = init_world_models(env, [48, 48], ["tanh"], 0.01, 0.5, None, [48, 48], ["relu"], 0.01, None, 2, True, setup)
    kw = dict(alg_type="sac_imit", sac_batch_size=B, expert_buffer_size=E, gamma=0.99, alg_seed=5, epsilon=0.2,
              device_replay_capacity=1000, gemm_mode=L.GEMM_FP32_SIMT, model_batch_size=40, model_lr=1e-3)
    alg = init_alg(0, env, env, env, actor, critics, q_targets, q_critics, models, kw, {}, None, None)
    rng = np.random.default_rng(1)
    n = 200
    rows = (rng.standard_normal((n, S)).astype(np.float32), rng.uniform(-1, 1, (n, A)).astype(np.float32),
            rng.standard_normal(n).astype(np.float32), rng.standard_normal((n, S)).astype(np.float32), rng.random(n) < 0.05)
    alg.env_data.add(*rows)
    alg.model_data.add(*rows)
    w0 = [m.get_weights() for m in alg.models]
    assert len(w0[0]) == 7 and np.allclose(w0[0][-1], np.log(0.5))          # logstd initialised to log(std_mult)
    cfg = NetCfg(S=S, A=A, actor_hidden=(32, 32), critic_hidden=(32, 32), model_hidden=(48, 48), model_acts=("tanh", "tanh"))
    st = to_torch_state(snapshot(alg, cfg))
    idx = np.stack([rng.permutation(n)[:40] for _ in range(2)])
    losses = alg._apply_model_grads(idx)
    mods = [[torch.from_numpy(np.asarray(w, np.float32)) for w in ws] for ws in w0]
    adam = dict(m=[[torch.zeros_like(w) for w in m] for m in mods], v=[[torch.zeros_like(w) for w in m] for m in mods], t=0)
    d = alg.model_data
    b = [{k: torch.as_tensor(getattr(d, k + "_all")[idx[m]]) for k in ("s", "a", "sp", "r")} for m in range(2)]
    out = apply_model_grads(cfg, mods, adam, b, st, dict(model_lr=1e-3, gaussian=True, scale_model_loss=True, reward_loss_coef=0.8))
    for m in range(2):
        assert abs(float(losses[0, 0, m]) - float(out["losses"][m])) < 1e-4 * abs(float(out["losses"][m]))
        got = alg.models[m].get_weights()
        for gw, ref, old in zip(got, out["models"][m], w0[m]):
            assert rel(np.asarray(gw) - np.asarray(old), ref.numpy() - np.asarray(old)) < 1e-3


def test_update_models_separate_reward_nn_mirror():
    """--separate_reward_nn through the class interface: the reward networks are uploaded when the joint optimiser is
    bound, one _apply_model_grads call matches the oracle on BOTH networks of both models, and
    get_reward_weights() / set_reward_weights() reach the device tables."""
    from oracle.sac_eo_oracle import apply_model_grads
    S, A, B, E = 11, 3, 32, 8
    np.random.seed(0)
    env = SyntheticEnv(S, A)
    actor = init_actor(env, [32, 32], ["relu"], 0.01, 1.0, "orthogonal", False, None, True, True, False)
    critics, q_targets, q_critics = init_critics(env, [32, 32], ["relu"], 1.0, None, 2, False, "orthogonal", False)
    setup = dict(SETUP, separate_reward_nn=True, reward_loss_coef=0.6)
    models = init_world_models(env, [48, 48], ["tanh"], 0.01, 1.0, None, [32, 40], ["relu"], 0.01, None, 2, False, setup)
    kw = dict(alg_type="sac_imit", sac_batch_size=B, expert_buffer_size=E, gamma=0.99, alg_seed=5, epsilon=0.2,
              device_replay_capacity=1000, gemm_mode=L.GEMM_FP32_SIMT, model_batch_size=40, model_lr=1e-3)
    alg = init_alg(0, env, env, env, actor, critics, q_targets, q_critics, models, kw, {}, None, None)
    rng = np.random.default_rng(1)
    n = 200
    rows = (rng.standard_normal((n, S)).astype(np.float32), rng.uniform(-1, 1, (n, A)).astype(np.float32),
            rng.standard_normal(n).astype(np.float32), rng.standard_normal((n, S)).astype(np.float32), rng.random(n) < 0.05)
    alg.env_data.add(*rows)
    alg.model_data.add(*rows)
    w0 = [m.get_weights() for m in alg.models]
    r0 = [[w.copy() for w in m.get_reward_weights()] for m in alg.models]
    assert [w.shape for w in r0[0]] == [(S + A, 32), (32,), (32, 40), (40,), (40, 1), (1,)] and w0[0][-2].shape == (48, S)
    cfg = NetCfg(S=S, A=A, actor_hidden=(32, 32), critic_hidden=(32, 32), model_hidden=(48, 48), model_acts=("tanh", "tanh"),
                 separate_reward_nn=True, reward_hidden=(32, 40), reward_acts=("relu", "relu"))
    st = to_torch_state(snapshot(alg, cfg))
    idx = np.stack([rng.permutation(n)[:40] for _ in range(2)])
    losses = alg._apply_model_grads(idx)
    mods = [[torch.from_numpy(np.asarray(w, np.float32)) for w in list(w0[k]) + list(r0[k])] for k in range(2)]
    adam = dict(m=[[torch.zeros_like(w) for w in m] for m in mods], v=[[torch.zeros_like(w) for w in m] for m in mods], t=0)
    d = alg.model_data
    b = [{k: torch.as_tensor(getattr(d, k + "_all")[idx[m]]) for k in ("s", "a", "sp", "r")} for m in range(2)]
    out = apply_model_grads(cfg, mods, adam, b, st, dict(model_lr=1e-3, reward_loss_coef=0.6))
    for m in range(2):
        assert abs(float(losses[0, 0, m]) - float(out["losses"][m])) < 1e-4 * abs(float(out["losses"][m]))
        got = list(alg.models[m].get_weights()) + list(alg.models[m].get_reward_weights())
        for gw, ref, old in zip(got, out["models"][m], mods[m]):
            assert rel(np.asarray(gw) - old.numpy(), ref.numpy() - old.numpy()) < 1e-3
    alg.models[0].set_reward_weights(r0[0])
    assert all(np.array_equal(a_, b_) for a_, b_ in zip(alg.models[0].get_reward_weights(), r0[0]))


def test_bc_alg_update_matches_oracle_with_reference_rng_order():
    """init_alg(alg_type='bc'): BC._update -> _update_actor, RNG order of BC.py:329-341."""
    from oracle.sac_eo_oracle import bc_update
    S, A, B, E = 11, 3, 32, 8
    alg, rng = build_alg("bc", S, A, B, E)
    cfg = NetCfg(S=S, A=A, actor_hidden=(32, 32), critic_hidden=(32, 32), model_hidden=(48, 48))
    st = snapshot(alg, cfg)
    alg.expert_data.add(rng.standard_normal((E, S)).astype(np.float32), rng.uniform(-1, 1, (E, A)).astype(np.float32),
                        rng.standard_normal(E).astype(np.float32), rng.standard_normal((E, S)).astype(np.float32),
                        np.zeros(E, bool))
    expert_reg = alg._expert_preprocess()
    q_before = alg.q_critics[0].get_weights()
    np.random.seed(321)
    alg._update(0, expert_reg)
    np.random.seed(321)
    order = np.arange(E)
    np.random.default_rng(5).shuffle(order)               # alg_seed=5 -> self.rng
    I1, I2 = np.array_split(order, 2)
    batch = dict(sE=expert_reg[0], spE=expert_reg[2], I1=I1, I2=I2, u3=np.random.normal(size=(E // 2, A)),
                 u4=np.random.normal(size=(E // 2, A)))
    o = bc_update(cfg, to_torch_state(st), batch, dict(lr_pi=1e-4))
    assert abs(alg.last_losses["BC_MSE_loss"] - float(o["mse"])) < 1e-4 * abs(float(o["mse"]))
    for got, new, old in zip(alg.actor.get_weights(), o["new"]["actor"], st["actor"]):
        assert rel(got - old, new.numpy() - old) < 1e-3
    for a_, b_ in zip(q_before, alg.q_critics[0].get_weights()):
        assert np.array_equal(a_, b_)
    assert alg.logger.train_dict["BC_MSE_loss"] == [alg.last_losses["BC_MSE_loss"]] or len(alg.logger.train_dict["BC_MSE_loss"]) == 1


@pytest.mark.parametrize("use_expert_actions", [True, False])
def test_adaptive_expert_weight_matches_oracle(use_expert_actions):
    """MSE bookkeeping on the expert transitions (SAC_expert.py:579-608) and the adaptive weight of
    _expert_preprocess (:381-418, scale_epsilon_by_true_MSE with min_mult and exp_mult) through the device forwards,
    against the oracle's model_mse_on_expert / adaptive_epsilon with the same NumPy draw."""
    from oracle.sac_eo_oracle import adaptive_epsilon, model_mse_on_expert
    S, A, B, E = 11, 3, 32, 8
    alg, rng = build_alg("sac_imit", S, A, B, E)
    alg.scale_epsilon_by_true_MSE, alg.use_expert_actions = True, use_expert_actions
    alg.min_mult, alg.exp_mult, alg.mult_coeff = True, True, 0.7
    alg.epsilon, alg.current_reward, alg.expert_reward = 0.5, 40.0, 100.0
    sE, aE = rng.standard_normal((E, S)).astype(np.float32), rng.uniform(-1, 1, (E, A)).astype(np.float32)
    spE = rng.standard_normal((E, S)).astype(np.float32)
    alg.expert_data.add(sE, aE, rng.standard_normal(E).astype(np.float32), spE, np.zeros(E, bool))
    cfg = NetCfg(S=S, A=A, actor_hidden=(32, 32), critic_hidden=(32, 32), model_hidden=(48, 48))
    st = to_torch_state(snapshot(alg, cfg))
    np.random.seed(9)
    on_exp, cf = alg._expert_mse_bookkeeping()
    np.random.seed(9)
    u = None if use_expert_actions else np.random.normal(size=(E, A))
    ref_on = float(model_mse_on_expert(cfg, st, sE, aE, spE, use_expert_actions=True))
    ref_cf = ref_on if use_expert_actions else float(model_mse_on_expert(cfg, st, sE, aE, spE, u=u))
    assert on_exp == pytest.approx(ref_on, rel=1e-4) and cf == pytest.approx(ref_cf, rel=1e-4)
    eps = alg._expert_preprocess()[3]
    ref_eps = adaptive_epsilon(0.5, scale_by_true_mse=True, mse_cf=ref_cf, j_cur=40.0, j_exp=100.0, min_mult=True,
                               exp_mult=True, mult_coeff=0.7)
    assert eps == pytest.approx(ref_eps, rel=1e-4)


def test_model_disagreement_weight_matches_oracle():
    """_calc_disc (SAC_expert.py:427-460) + the scale_max/median/total_disc branches of _expert_preprocess."""
    from oracle.sac_eo_oracle import adaptive_epsilon, model_sample
    S, A, B, E = 11, 3, 32, 8
    alg, rng = build_alg("sac_imit", S, A, B, E)
    alg.use_expert_actions = True
    sE, aE = rng.standard_normal((E, S)).astype(np.float32), rng.uniform(-1, 1, (E, A)).astype(np.float32)
    alg.expert_data.add(sE, aE, rng.standard_normal(E).astype(np.float32), rng.standard_normal((E, S)).astype(np.float32),
                        np.zeros(E, bool))
    cfg = NetCfg(S=S, A=A, actor_hidden=(32, 32), critic_hidden=(32, 32), model_hidden=(48, 48))
    st = to_torch_state(snapshot(alg, cfg))
    p0 = model_sample(cfg, st["m1"], torch.as_tensor(sE), torch.as_tensor(aE), st).numpy()
    p1 = model_sample(cfg, st["m2"], torch.as_tensor(sE), torch.as_tensor(aE), st).numpy()
    disc = np.linalg.norm(p0 - p1, axis=1)
    ratio, mx, med, tot = alg._calc_disc(sE, aE, None)
    assert mx == pytest.approx(float(disc.max()), rel=1e-4) and med == pytest.approx(float(np.median(disc)), rel=1e-4)
    assert tot == pytest.approx(float(disc.sum()), rel=1e-4) and np.allclose(ratio, disc / disc.sum(), rtol=1e-4)
    alg.epsilon = 2.0
    for flag, mode in (("scale_max_disc", "max"), ("scale_median_disc", "median"), ("scale_total_disc", "total")):
        alg.scale_max_disc = alg.scale_median_disc = alg.scale_total_disc = False
        setattr(alg, flag, True)
        assert alg._expert_preprocess()[3] == pytest.approx(adaptive_epsilon(2.0, disc_mode=mode, disc=disc), rel=1e-4)
