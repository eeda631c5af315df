"""B200-native batched convex-MPC engine (drop-in for the reference's solveDenseMPC hot path)."""
from . import engine, gait, interface, records, ticks, workloads  # noqa: F401

__all__ = ["engine", "gait", "interface", "records", "ticks", "workloads", "sharding"]


def __getattr__(name):  # sharding imports torch.distributed: load it on first use only
    if name == "sharding":
        import importlib
        return importlib.import_module(".sharding", __name__)
    raise AttributeError(name)
