"""Import shim (test infrastructure, NOT reference code, NOT product code).

The reference imports ``compressai.ans`` / ``compressai._CXX`` / ``compressai.ops`` from
the un-pinned pip package CompressAI (reference: src/compress/entropy_models/entropy_models.py:13,24-36).
oracle/build_ref.py compiles the reference's own vendored copies of those two native modules
into this package directory; this file only provides the one python symbol the reference asks
the package itself for.
"""


def available_entropy_coders():
    return ["ans"]
