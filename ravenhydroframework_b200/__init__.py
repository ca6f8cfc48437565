"""raven-b200: B200-native implementation of Raven's per-timestep water/energy balance."""
from . import api, synth  # noqa: F401
from .api import GpuSimulation, RavenError, RavenModel  # noqa: F401

__all__ = ["api", "synth", "RavenModel", "GpuSimulation", "RavenError"]
