"""Recipe for oracle/_ref: the UNMODIFIED reference hot-path modules, staged so they travel to the GPU box
(TEST / BASELINE INFRASTRUCTURE ONLY -- never imported by the product package).

/root/reference exists only in the build container.  `build()` copies, byte for byte,

    /root/reference/src/hallmd/models/plume.py     (current_density, plume.py:21-159)
    /root/reference/src/hallmd/models/cathode.py   (cathode_coupling, cathode.py:16-38)

into oracle/_ref/hallmd/models/ next to EMPTY package __init__ files (the reference's own models/__init__.py also imports
thruster.py, which needs Julia tooling that is not part of this path) and records their SHA-256 in oracle/_ref/MANIFEST.json.
oracle/_ref/ is git-ignored (reference sources never enter this repository's history) but not gpurun-ignored, so
`bench.py` can time the reference's own code on the GPU box's host cores (`cpu_baseline.kind = "reference"`; the
reference hard-codes 91 angles, plume.py:53).  `oracle.ref_import.load()` falls back to this copy when
/root/reference is absent.
"""
from __future__ import annotations

import hashlib
import json
import shutil
from pathlib import Path

HERE = Path(__file__).resolve().parent
REF_DIR = HERE / '_ref'
SRC = Path('/root/reference/src/hallmd/models')
FILES = ('plume.py', 'cathode.py')


def available() -> bool:
    return all((REF_DIR / 'hallmd' / 'models' / f).is_file() for f in FILES)


def build(verbose: bool = False) -> bool:
    """Stage the two reference modules; returns False (and changes nothing) where /root/reference does not exist."""
    if not all((SRC / f).is_file() for f in FILES):
        return False
    dst = REF_DIR / 'hallmd' / 'models'
    dst.mkdir(parents=True, exist_ok=True)
    (REF_DIR / 'hallmd' / '__init__.py').write_text('')
    (dst / '__init__.py').write_text('')
    manifest = {}
    for f in FILES:
        shutil.copyfile(SRC / f, dst / f)
        manifest[f] = {'source': str(SRC / f), 'sha256': hashlib.sha256((dst / f).read_bytes()).hexdigest()}
    (REF_DIR / 'MANIFEST.json').write_text(json.dumps(manifest, indent=1))
    if verbose:
        print(f'staged {", ".join(FILES)} under {dst}')
    return True


if __name__ == '__main__':
    print('built' if build(verbose=True) else 'reference sources not available here')
