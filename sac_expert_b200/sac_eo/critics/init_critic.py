"""``init_critics`` with the reference's signature (``/root/reference/sac_eo/critics/init_critic.py:5-38``):
returns ``(critics, q_targets, q_critics)``; targets start as copies of the live critics.  The on-policy
``VCritic`` list is out of scope and returned empty."""
from .critics import QCritic


def init_critics(env, critic_layers, critic_activations, critic_gain, critic_weights, num_models, critic_ensemble,
                 critic_init_type, critic_layer_norm):
    q_critics = [QCritic(env, critic_layers, critic_activations, critic_gain) for _ in range(2)]
    q_targets = [QCritic(env, critic_layers, critic_activations, critic_gain) for _ in range(2)]
    for tgt, live in zip(q_targets, q_critics):
        tgt.set_weights(live.get_weights())
    return [], q_targets, q_critics
