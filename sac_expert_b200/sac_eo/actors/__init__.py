from .init_actor import init_actor
