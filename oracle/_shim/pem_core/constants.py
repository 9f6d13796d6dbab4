"""Constants the reference imports from pem_core. Values are NOT present anywhere under the
reference tree; TORR_2_PA = 133.322 is the value historically used by hallmd ("parity unpinned"
for this one constant -- every entry point of this repo therefore takes `torr_2_pa=` explicitly)."""
TORR_2_PA = 133.322
# Only needed so that hallmd/models/__init__.py can import thruster.py (thruster.py:31)
AVOGADRO_CONSTANT = 6.02214076e23
FUNDAMENTAL_CHARGE = 1.602176634e-19
MOLECULAR_WEIGHTS = {'Xenon': 131.293, 'Argon': 39.948, 'Krypton': 83.798,
                     'Bismuth': 208.9804, 'Mercury': 200.59}
