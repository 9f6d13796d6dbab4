#!/usr/bin/env python
"""Developer probe: tools/angle_sweep.py <angles...> against every build/variants/libhpem_*.so."""
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
for lib in sorted((ROOT / 'build' / 'variants').glob('libhpem_*.so')):
    print('==', lib.stem, flush=True)
    subprocess.run([sys.executable, str(ROOT / 'tools' / 'angle_sweep.py')] + sys.argv[1:], env=dict(os.environ, HPEM_LIBRARY=str(lib)))
