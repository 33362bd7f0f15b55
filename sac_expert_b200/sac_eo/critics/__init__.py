from .init_critic import init_critics
