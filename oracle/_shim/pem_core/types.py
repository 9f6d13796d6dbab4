"""Type aliases the reference imports from pem_core.types (plume.py:13, cathode.py:11)."""
from typing import Any

Dataset = dict
ArrayLike = Any
PathLike = Any
