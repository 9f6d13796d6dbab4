"""Model callables with the reference's names and signatures (hallmd/models/__init__.py:15-19).

`hallthruster_jl` is intentionally absent: the HallThruster.jl solve stays external (BASELINE.json north_star)."""
from .cathode import cathode_coupling
from .plume import current_density, plume_cathode

__all__ = ['cathode_coupling', 'current_density', 'plume_cathode']
