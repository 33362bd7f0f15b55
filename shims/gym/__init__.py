"""Minimal ``gym`` stand-in: the space classes and helpers the reference imports (``spaces.Box``, ``spaces.Discrete``,
``spaces.utils.flatdim``, ``wrappers.RescaleAction``; ``/root/reference/sac_eo/actors/init_actor.py:2``,
``envs/wrappers/gym_wrapper.py:1-8``).  No simulators: ``gym.make`` raises - use ``--env_type synthetic``."""
from . import spaces, wrappers


def make(name, **kw):
    raise RuntimeError("shims/gym has no simulators; run with --env_type synthetic (sac_eo/envs/init_env.py)")
