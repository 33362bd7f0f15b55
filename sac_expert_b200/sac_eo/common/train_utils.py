"""``gather_inputs`` / ``import_inputs`` / ``organize_rms_inputs`` with the reference's behaviour
(``/root/reference/sac_eo/common/train_utils.py``): group the parsed flags into the kwargs dicts ``train()`` consumes,
optionally seeding them from a saved log (weights, normaliser statistics)."""
import os
import pickle

from .train_parser import all_kwargs


def gather_inputs(args):
    a = vars(args)
    return {group: {name: a[name] for name in names} for group, names in all_kwargs.items()}


def import_inputs(inputs_dict):
    setup = inputs_dict["setup_kwargs"]
    path, fname, idx = setup["import_path"], setup["import_file"], setup["import_idx"]
    train_idx = setup["idx"] - setup["runs_start"]
    actor_w = critic_w = model_w = reward_w = rms = None
    if path and fname:
        with open(os.path.join(path, fname), "rb") as f:
            log = pickle.load(f)
        if idx is None:
            idx = train_idx if len(log) > train_idx else 0
        assert idx < len(log), "import_idx too large"
        param, final = log[idx]["param"], log[idx]["final"]
        if setup["import_all"]:
            inputs_dict = param
            inputs_dict["setup_kwargs"] = setup
        else:
            for k in ("env_kwargs", "actor_kwargs", "critic_kwargs"):
                inputs_dict[k] = param[k]
        actor_w, critic_w, rms = final["actor_weights"], final["critic_weights"], final["rms_stats"]
        if "model_weights" in final and "reward_weights" in final:      # only use model inputs if the import has model weights
            model_w, reward_w = final["model_weights"], final["reward_weights"]
            inputs_dict["model_kwargs"], inputs_dict["model_setup_kwargs"] = param["model_kwargs"], param["model_setup_kwargs"]
    inputs_dict["actor_kwargs"]["actor_weights"] = actor_w
    inputs_dict["critic_kwargs"]["critic_weights"] = critic_w
    inputs_dict["model_kwargs"]["model_weights"] = model_w
    inputs_dict["model_kwargs"]["reward_weights"] = reward_w
    inputs_dict["alg_kwargs"]["init_rms_stats"] = rms
    return inputs_dict


def organize_rms_inputs(logs_rms):
    if "rms_stats" in logs_rms:
        return logs_rms["rms_stats"]
    return {k: {"t": logs_rms[p + "_t"], "mean": logs_rms[p + "_mean"], "var": logs_rms[p + "_var"]}
            for k, p in (("s_rms", "s"), ("a_rms", "a"), ("r_rms", "r"), ("delta_rms", "delta"), ("ret_rms", "ret"))}
