"""hallthrusterpem_b200 -- B200-native (sm_100a) plume + cathode Monte-Carlo hot path of HallThrusterPEM.

Drop-in for `hallmd.models.plume.current_density` and `hallmd.models.cathode.cathode_coupling`; the arithmetic
runs in hand-written CUDA kernels behind the C ABI of include/hpem.h (loaded with ctypes).  No CPU fallback.
"""
__version__ = '0.1.0'

from . import synthetic  # noqa: F401  (pure NumPy; importable without the CUDA library)

__all__ = ['models', 'synthetic', '__version__']


def __getattr__(name):
    if name == 'models':
        import importlib
        return importlib.import_module('.models', __name__)
    raise AttributeError(name)
