"""Minimal stand-in for the un-vendored `pem_core` dependency (TEST INFRASTRUCTURE ONLY).

The reference pins JANUS-Institute/pem_core@77b5b083 (uv.lock:1655-1657) but does not vendor it.
Only the three names the hot path touches are provided so that `/root/reference/src/hallmd`
imports unmodified (plume.py:11-13, cathode.py:10-11, thruster.py:31).
"""
import logging


def get_logger(name):
    return logging.getLogger(name)
