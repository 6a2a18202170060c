from .step import discrete_step, evolve

__all__ = ["discrete_step", "evolve"]
