"""Build the REAL reference native coder into oracle/_ref/ (test infrastructure only).

Compiles, from the sources where they lie under /root/reference (never copied into this repo):
  * src/compress/cpp_exts/rans/rans_interface.cpp (+ src/third_party/ryg_rans/rans64.h) -> compressai/ans*.so
  * src/compress/cpp_exts/ops/ops.cpp                                                   -> compressai/_CXX*.so
and lays the python import shims (oracle/shims/) beside them so that
``PYTHONPATH=oracle/_ref:/root/reference/src`` makes ``compress.models.ChannelProgresssiveWACNN``
importable unmodified.

The four ``-include`` flags are needed because the vendored rans_interface.hpp:19-21 uses
pybind11/std::vector/std::string before including them (upstream CompressAI's header has the
includes; this copy lost them).

On the GPU box /root/reference does not exist: the prebuilt oracle/_ref travels with the repo
snapshot (git-ignored, not gpurun-ignored) and this script is a no-op there.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("PCODEC_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REF, "src/compress/cpp_exts/rans/rans_interface.cpp"))


def ref_built() -> bool:
    suf = sysconfig.get_config_var("EXT_SUFFIX")
    return all(os.path.isfile(os.path.join(OUT, "compressai", n + suf)) for n in ("ans", "_CXX"))


def build(force: bool = False) -> bool:
    """Returns True when oracle/_ref is usable after the call."""
    if ref_built() and not force:
        return True
    if not reference_available():
        return ref_built()
    import pybind11

    suf = sysconfig.get_config_var("EXT_SUFFIX")
    pkg = os.path.join(OUT, "compressai")
    os.makedirs(pkg, exist_ok=True)
    # shims (our own files) -> _ref
    shims = os.path.join(HERE, "shims")
    for root, _dirs, files in os.walk(shims):
        rel = os.path.relpath(root, shims)
        os.makedirs(os.path.join(OUT, rel), exist_ok=True)
        for f in files:
            if f.endswith(".py"):
                shutil.copyfile(os.path.join(root, f), os.path.join(OUT, rel, f))
    inc = ["-I" + sysconfig.get_paths()["include"], "-I" + pybind11.get_include()]
    common = ["g++", "-O3", "-DNDEBUG", "-std=c++17", "-shared", "-fPIC"] + inc
    src = os.path.join(REF, "src")
    cmds = [
        common
        + ["-include", "vector", "-include", "string", "-include", "pybind11/pybind11.h", "-include", "pybind11/stl.h",
           "-I" + os.path.join(src, "third_party/ryg_rans"), "-I" + os.path.join(src, "compress/cpp_exts/rans"),
           os.path.join(src, "compress/cpp_exts/rans/rans_interface.cpp"), "-o", os.path.join(pkg, "ans" + suf)],
        common + [os.path.join(src, "compress/cpp_exts/ops/ops.cpp"), "-o", os.path.join(pkg, "_CXX" + suf)],
    ]
    for c in cmds:
        subprocess.check_call(c)
    return True


def import_reference():
    """Put oracle/_ref and the reference python tree on sys.path (container only)."""
    if not reference_available():
        raise RuntimeError("reference tree not present (expected on the GPU box)")
    build()
    for p in (OUT, os.path.join(REF, "src")):
        if p not in sys.path:
            sys.path.insert(0, p)


def import_ref_coder():
    """Import only the compiled reference coder (works on the GPU box from the prebuilt _ref)."""
    if not ref_built():
        if not build():
            raise RuntimeError("oracle/_ref not built and reference tree absent")
    if OUT not in sys.path:
        sys.path.insert(0, OUT)
    import compressai.ans as ans  # type: ignore
    import compressai._CXX as cxx  # type: ignore

    return ans, cxx


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv)
    print("oracle/_ref built:", ok)
