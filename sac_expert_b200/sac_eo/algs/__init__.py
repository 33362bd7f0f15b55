from .init_alg import init_alg
