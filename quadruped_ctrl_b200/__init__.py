"""B200-native batched convex-MPC engine (drop-in for the reference's solveDenseMPC hot path)."""
from . import engine, gait, interface, records, workloads  # noqa: F401

__all__ = ["engine", "gait", "interface", "records", "workloads"]
