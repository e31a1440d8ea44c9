"""gprotation_b200: the quasi-periodic GP log-likelihood of RuthAngus/GProtation on B200 (sm_100a)."""
from .engine import GPEngine, GPGroup, GProtError  # noqa: F401

__all__ = ["GPEngine", "GPGroup", "GProtError"]
