"""GPU: the reference's training entry point on the CUDA path (round-1 VERDICT row b2).  ``train.main`` with the
reference's flag names drives ``init_env -> init_actor -> init_critics -> init_world_models -> init_alg -> alg.train``
(``/root/reference/sac_eo/train.py:33-107``, ``algs/SAC_expert.py:685-824``) on the synthetic MuJoCo-shaped environment:
initial data collection, per-episode model fitting + adaptive expert weight, one ``_update`` per environment step,
evaluation and the checkpoint pickle."""
import os
import pickle
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

COMMON = ["--env_type", "synthetic", "--env_name", "hopper", "--task_name", "60", "--actor_squash", "--actor_per_state_std",
          "--actor_layers", "256", "256", "--critic_layers", "256", "256", "--actor_activations", "relu",
          "--critic_activations", "relu", "--model_layers", "128", "128", "--env_horizon", "60", "--env_batch_size_init", "150",
          "--total_timesteps", "350", "--model_num_epochs", "1", "--model_batch_size", "50", "--eval_freq", "200",
          "--eval_num_traj", "1", "--seed", "3"]


@pytest.mark.parametrize("alg_type,extra", [("sac_imit", ["--scale_epsilon_by_true_MSE"]), ("sac", []), ("bc", [])])
def test_train_entry_point_runs_200_steps(tmp_path, alg_type, extra):
    from sac_expert_b200.sac_eo import train as T
    out = T.main(COMMON + ["--alg_type", alg_type, "--save_path", str(tmp_path)] + extra)
    with open(out, "rb") as f:
        logs = pickle.load(f)
    assert len(logs) == 1 and set(logs[0]) == {"param", "train", "final"}
    log = logs[0]
    assert {"actor_weights", "critic_weights", "rms_stats"} <= set(log["final"])
    assert log["param"]["alg_kwargs"]["alg_type"] == alg_type
    tr = log["train"]
    assert "J_tot" in tr and "J_tot_eval" in tr and np.all(np.isfinite(tr["J_tot"]))
    assert 290 <= int(np.sum(tr["steps"])) <= 350          # the unfinished last episode is not logged (SAC_expert.py:751-764)
    if alg_type == "sac_imit":
        assert len(tr["p_loss"]) == 200 and np.all(np.isfinite(tr["p_loss"])) and np.all(np.isfinite(tr["alpha_loss"]))
        assert np.all((tr["epsilon"] > 0) & (tr["epsilon"] <= 1))            # 1 / (epsilon * MSE_cf + 1)
        assert "model_MSE_on_expert_data" in tr or "model_updates" in tr
    if alg_type == "bc":
        assert len(tr["BC_MSE_loss"]) == 200 and np.all(np.isfinite(tr["BC_MSE_loss"]))
    w = log["final"]["actor_weights"]
    assert [np.shape(x) for x in w] == [(11, 256), (256,), (256, 256), (256,), (256, 6), (6,)]
    assert all(np.all(np.isfinite(x)) for x in w)


def test_reference_package_name_resolves_to_the_cuda_path():
    """``shims/``: `import sac_eo...` (the reference's package name) yields the modules of this repository, and the three
    TensorFlow / gym touch points of the reference's train.py exist - what a maintainer needs to run it unmodified."""
    sys.path.insert(0, os.path.join(ROOT, "shims"))
    try:
        import tensorflow as tf
        import gym
        import sac_eo.algs.SAC_expert as A
        import sac_expert_b200.sac_eo.algs.SAC_expert as B
        from sac_eo.envs import init_env
        from sac_eo.algs import init_alg
        assert A is B and callable(init_env) and callable(init_alg)
        assert tf.config.experimental.list_physical_devices("GPU") == [] and tf.random.set_seed(0) is None
        assert gym.spaces.utils.flatdim(gym.spaces.Box(-1.0, 1.0, (3,))) == 3
    finally:
        sys.path.remove(os.path.join(ROOT, "shims"))
        for k in [k for k in sys.modules if k == "tensorflow" or k == "gym" or k.startswith("gym.")]:
            sys.modules.pop(k, None)
