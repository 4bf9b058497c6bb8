"""Shim: re-export the reference's own parametrizer for layers/gdn.py:9."""
from compress.ops.parametrizers import NonNegativeParametrizer  # noqa: F401
