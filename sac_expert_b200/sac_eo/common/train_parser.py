"""Command-line surface of ``train.py`` - the flags, defaults and kwarg groups of the reference's parser
(``/root/reference/sac_eo/common/train_parser.py``) that reach the SAC / SAC-EO / BC path, written as one table.
Flags of the on-policy model-based path (simulated rollouts, V-critic fitting, adaptive model horizon) are accepted and
ignored by the algorithms that do not use them, like in the reference."""
import argparse

# (flag, type | "true" | "false:<dest>", default, nargs)
_T = "true"
_FLAGS = {
    "setup_kwargs": [("runs", int, 1), ("runs_start", int, 0), ("cores", int, None), ("seed", int, 0),
                     ("setup_seed", int, None), ("sim_seed", int, None), ("eval_seed", int, None),
                     ("expert_seed", int, None), ("save_path", str, "./logs"), ("save_file", str, None),
                     ("import_path", str, "./logs"), ("import_file", str, None), ("import_idx", int, None),
                     ("import_all", _T, None), ("expert_file", str, None), ("expert_path", str, "./experts")],
    "env_kwargs": [("env_type", str, "gym"), ("env_name", str, "Pendulum-v1"), ("task_name", str, None)],
    "actor_kwargs": [("actor_layers", int, [64, 64], "+"), ("actor_activations", str, ["tanh"], "+"),
                     ("actor_gain", float, 0.01), ("actor_std_mult", float, 1.0), ("actor_init_type", str, "orthogonal"),
                     ("actor_layer_norm", _T, None), ("actor_per_state_std", _T, None), ("actor_squash", _T, None)],
    "critic_kwargs": [("critic_layers", int, [64, 64], "+"), ("critic_activations", str, ["tanh"], "+"),
                      ("critic_gain", float, 1.0), ("critic_ensemble", _T, None), ("critic_init_type", str, "orthogonal"),
                      ("critic_layer_norm", _T, None)],
    "model_kwargs": [("gaussian_model", _T, None), ("num_models", int, 2), ("model_layers", int, [512, 512], "+"),
                     ("model_activations", str, ["relu"], "+"), ("model_gain", float, 0.01), ("model_std_mult", float, 1.0),
                     ("reward_layers", int, [512, 512], "+"), ("reward_activations", str, ["relu"], "+"),
                     ("reward_gain", float, 0.01)],
    "model_setup_kwargs": [("separate_reward_nn", _T, None), ("reward_loss_coef", float, 1.0), ("scale_model_loss", _T, None),
                           ("delta_clip_loss", float, None), ("reward_clip_loss", float, None), ("delta_clip_pred", float, None),
                           ("reward_clip_pred", float, None)],
    "alg_kwargs": [
        # buffers
        ("gamma", float, 0.995), ("lam", float, 0.97), ("env_buffer_size", float, None), ("sim_buffer_size", float, None),
        ("model_buffer_size", float, 1e5), ("expert_buffer_size", float, 20),
        # training loop
        ("checkpoint_file", str, "TEMPLOG"), ("save_freq", float, None), ("eval_freq", float, None), ("eval_num_traj", int, 5),
        ("alg_type", str, "sac_imit"), ("mf_algo", str, "trpo"), ("total_timesteps", float, 5e5), ("env_horizon", int, 1000),
        ("env_batch_type", str, "steps"), ("env_batch_size_init", int, 5000), ("env_batch_size", int, 3000),
        ("s_noise_std", float, 0.0), ("s_noise_type", str, "all"), ("sim_horizon", int, 5), ("sim_batch_type", str, "steps"),
        ("sim_batch_size", int, 10000), ("exp_batch_type", str, "steps"),
        # model fitting
        ("model_lr", float, 1e-3), ("model_num_epochs", int, 10), ("model_batch_size", int, 200),
        ("no_model_batch_shuffle", "false:model_batch_shuffle", True), ("model_max_updates", float, 1e5),
        ("model_max_grad_norm", float, None), ("model_holdout_ratio", float, 0.0), ("model_holdout_epochs", int, 5),
        ("reset_model_optimizer", _T, None),
        # actor-critic (on-policy path)
        ("critic_lr", float, 3e-4), ("critic_update_it", int, 10), ("critic_nminibatch", int, 32), ("num_mf_updates", int, 25),
        # expert regularisation
        ("epsilon", float, 1e-3), ("scale_epsilon_by_true_MSE", _T, None), ("scale_max_disc", _T, None),
        ("scale_median_disc", _T, None), ("scale_total_disc", _T, None), ("use_expert_actions", _T, None), ("min_mult", _T, None),
        ("exp_mult", _T, None), ("mult_coeff", float, 1.0), ("init_from_expert", _T, None), ("max_exp_state_ratio", float, 0.25),
        # SAC / MBPO
        ("init_temperature", float, 1e-1), ("q_crit_lr", float, 3e-4), ("mbpo_actor_lr", float, 1e-4), ("mbpo_alpha_lr", float, 1e-4),
        ("mbpo_E", int, 1000), ("mbpo_G", int, 3), ("mbpo_M", int, 400), ("sac_batch_size", int, 256), ("expert_batch_size", int, None),
        ("soft_tau", float, 5e-3), ("target_update_int", int, 1), ("real_step_mod", int, 3), ("random_act", _T, None),
        ("update_normalizers", _T, None), ("only_model_normalizer", _T, None), ("adaptive_model_horizon", _T, None),
        ("modelhorx", float, 1), ("modelhory", float, 15), ("modelhora", float, 20), ("modelhorb", float, 100)],
    "mf_update_kwargs": [("no_adv_center", "false:adv_center", True), ("no_adv_scale", "false:adv_scale", True), ("ent_reg", _T, None),
                         ("alpha_lr", float, 3e-4), ("delta_trpo", float, 0.02), ("cg_it", int, 20), ("trust_sub", int, 1),
                         ("trust_damp", float, 0.01), ("kl_maxfactor", float, 1.5), ("actor_update_it", int, 10),
                         ("actor_nminibatch", int, 32), ("actor_lr", float, 3e-4), ("eps_ppo", float, 0.2), ("max_grad_norm", float, 0.5),
                         ("no_adaptlr", "false:adaptlr", True), ("adapt_factor", float, 0.03), ("adapt_minthresh", float, 0.0),
                         ("adapt_maxthresh", float, 1.0)],
}
# flags the reference parses but lists in no kwarg group except through train.py itself
_EXTRA = [("alg_seed", int, None)]

parser = argparse.ArgumentParser()
all_kwargs = {}
for _group, _rows in list(_FLAGS.items()) + [(None, _EXTRA)]:
    _names = []
    for _row in _rows:
        _flag, _typ, _dft = _row[:3]
        _dest = _flag
        if _typ == _T:
            parser.add_argument("--" + _flag, action="store_true")
        elif isinstance(_typ, str) and _typ.startswith("false:"):
            _dest = _typ.split(":")[1]
            parser.add_argument("--" + _flag, dest=_dest, default=True, action="store_false")
        elif len(_row) > 3:
            parser.add_argument("--" + _flag, nargs=_row[3], type=_typ, default=_dft)
        else:
            parser.add_argument("--" + _flag, type=_typ, default=_dft)
        _names.append(_dest)
    if _group:
        all_kwargs[_group] = _names
# the critic group also receives num_models, the train group also save_path (train_parser.py:75-78, :174-176)
all_kwargs["critic_kwargs"].insert(4, "num_models")
all_kwargs["alg_kwargs"].insert(6, "save_path")


def create_train_parser():
    return parser
