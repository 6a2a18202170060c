from .stats import compute_solution_variables, compute_stats

__all__ = ["compute_solution_variables", "compute_stats"]
