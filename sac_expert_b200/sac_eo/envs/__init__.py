from .init_env import init_env
