"""Drop-in replacement of the reference `utils` package (utils/__init__.py:1)."""
from . import losses, metrics
from .metrics import Metric, Metric_for_Loss

__all__ = ["losses", "metrics", "Metric", "Metric_for_Loss"]
