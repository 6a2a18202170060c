from .utils import sample_ellipse_cosine

__all__ = ["sample_ellipse_cosine"]
