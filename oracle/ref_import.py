"""Load the UNMODIFIED reference hot-path functions (TEST INFRASTRUCTURE ONLY).

Works only where /root/reference exists (the build container).  Puts the `pem_core` shim and
/root/reference/src on sys.path, then imports hallmd.models.{plume,cathode}.
"""
import os
import sys
from pathlib import Path

REFERENCE_SRC = Path(os.environ.get('HPEM_REFERENCE_SRC', '/root/reference/src'))
_SHIM = Path(__file__).resolve().parent / '_shim'


def available() -> bool:
    return (REFERENCE_SRC / 'hallmd' / 'models' / 'plume.py').is_file()


def load():
    """Return (current_density, cathode_coupling, TORR_2_PA) of the real reference."""
    if not available():
        raise RuntimeError(f'reference sources not found under {REFERENCE_SRC}')
    os.environ.setdefault('HOME', '/tmp')  # thruster.py:56 reads it at import time
    for p in (str(_SHIM), str(REFERENCE_SRC)):
        if p not in sys.path:
            sys.path.insert(0, p)
    from hallmd.models.cathode import cathode_coupling
    from hallmd.models.plume import current_density
    from pem_core.constants import TORR_2_PA
    return current_density, cathode_coupling, TORR_2_PA
