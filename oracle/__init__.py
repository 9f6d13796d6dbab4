"""CPU oracle for the plume+cathode hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline / `--impl reference`
legs may import from here.  Nothing under `hallthrusterpem_b200/` imports this package; the
product path raises if its CUDA library is missing.

Contents
--------
* `ref_restated`  : NumPy/SciPy restatement of /root/reference/src/hallmd/models/plume.py:38-159
                    and cathode.py:24-38, with `n_angles` (plume.py:53 hard-codes 91) and
                    `torr_2_pa` (pem_core constant, un-vendored) as parameters.
* `ref_import`    : loads the UNMODIFIED reference functions from /root/reference/src through the
                    `_shim/pem_core` stand-in (only possible in the build container).
* `make_golden`   : mints tests/golden/*.npz from the real reference (A=91) and the restatement
                    (A != 91) after asserting the two are bit-identical at A=91.
* `inputs`        : seeded synthetic input generators for BASELINE.json's configs.

Pinning status: the restatement is asserted bit-equal to the imported reference on seeded
inputs (tests/test_oracle_vs_reference.py, runs only where /root/reference exists) and against
the committed golden vectors everywhere else.  The one un-pinned quantity is the numeric value of
pem_core.constants.TORR_2_PA (not in the reference tree): "parity unpinned" for that constant.
"""
