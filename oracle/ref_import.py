"""Load the UNMODIFIED reference hot-path functions (TEST INFRASTRUCTURE ONLY).

In the build container /root/reference/src is imported directly (the whole `hallmd` package, through the `pem_core`
shim).  Where /root/reference does not exist (the GPU box) the byte-identical copies of plume.py / cathode.py that
`oracle/build_ref.py` staged under oracle/_ref are used instead.
"""
import os
import sys
from pathlib import Path

REFERENCE_SRC = Path(os.environ.get('HPEM_REFERENCE_SRC', '/root/reference/src'))
_SHIM = Path(__file__).resolve().parent / '_shim'


_STAGED = Path(__file__).resolve().parent / '_ref'


def available() -> bool:
    return (REFERENCE_SRC / 'hallmd' / 'models' / 'plume.py').is_file() or (_STAGED / 'hallmd' / 'models' / 'plume.py').is_file()


def load():
    """Return (current_density, cathode_coupling, TORR_2_PA) of the real reference."""
    if not available():
        raise RuntimeError(f'reference sources found neither under {REFERENCE_SRC} nor under {_STAGED}')
    os.environ.setdefault('HOME', '/tmp')  # thruster.py:56 reads it at import time
    src = REFERENCE_SRC if (REFERENCE_SRC / 'hallmd' / 'models' / 'plume.py').is_file() else _STAGED
    for p in (str(_SHIM), str(src)):
        if p not in sys.path:
            sys.path.insert(0, p)
    from hallmd.models.cathode import cathode_coupling
    from hallmd.models.plume import current_density
    from pem_core.constants import TORR_2_PA
    return current_density, cathode_coupling, TORR_2_PA
