"""``init_env`` with the reference's signature (``/root/reference/sac_eo/envs/init_env.py:3-23``).  ``gym`` / ``dmc``
environments need the simulators (not installable offline; out of scope); ``synthetic`` - added for this
container - gives MuJoCo-SHAPED spaces with cheap linear dynamics (``envs/synthetic.py``), ``env_name`` one of
hopper / halfcheetah / ant / humanoid / pendulum or ``"S,A"``."""
from .synthetic import SyntheticEnv

_SHAPES = {"hopper": (11, 3), "halfcheetah": (17, 6), "ant": (27, 8), "humanoid": (376, 17), "pendulum": (3, 1)}


def init_env(env_type, env_name, task_name=None):
    if env_type == "synthetic":
        key = str(env_name).lower().split("-")[0]
        if key in _SHAPES:
            s_dim, a_dim = _SHAPES[key]
        else:
            try:
                s_dim, a_dim = (int(v) for v in str(env_name).split(","))
            except ValueError:
                raise ValueError("synthetic env_name must be one of %s or 'S,A'" % sorted(_SHAPES))
        horizon = int(task_name) if task_name else 1000
        return SyntheticEnv(s_dim, a_dim, horizon=horizon)
    if env_type in ("gym", "dmc"):
        raise ValueError("env_type '%s' needs its simulator package; this build ships env_type 'synthetic' only" % env_type)
    raise ValueError("Only gym and dmc env_type supported")
