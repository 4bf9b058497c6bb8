"""TEST INFRASTRUCTURE (checker side).  State dict of the synthetic-weights model WITHOUT the product package.

bench.py's reference arm and cpu_baseline leg time the oracle port (oracle/codec_port.py) on the host cores; they
must not import, load or execute anything of progressivecodec_b200 (no libpcodec_b200.so in their process).  The
oracle needs a reference-layout state dict:
  * parameter names / shapes and the small constant buffers (GDN pedestals and bounds, relative_position_index,
    EntropyBottleneck.target, scale_bound) come from a committed skeleton, tests/golden/<case>_state_skeleton.npz,
    written by `python -m oracle.synthetic_state --write` (which instantiates the model once, here in the build
    container);
  * parameter VALUES come from the name-keyed generator in progressivecodec_b200/synthetic.py — a pure-torch helper
    that is executed from its file path (importlib), so the package __init__ (and with it the CUDA binding) never runs;
  * the CDF tables are rebuilt by the oracle's own restatement of update() (entropy_port.GaussianTables.build /
    BottleneckTables.rebuild — reference entropy_models.py:354-393, 599-624).
tests/test_oracle_golden.py pins the result bit-identical to the product model's state_dict() after
apply_synthetic_weights() + update().
"""
from __future__ import annotations

import importlib.util
import json
import os
import sys
from typing import Dict

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def _generator_module():
    path = os.path.join(ROOT, "progressivecodec_b200", "synthetic.py")
    spec = importlib.util.spec_from_file_location("_pcodec_synthetic_by_path", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def synthetic_image(shape, seed: int) -> torch.Tensor:
    return _generator_module().synthetic_image(shape, seed)


def synthetic_state_dict(case: str = "authors", seed: int = 0) -> Dict[str, torch.Tensor]:
    from .entropy_port import BottleneckTables, GaussianTables

    sk = np.load(os.path.join(GOLDEN, f"{case}_state_skeleton.npz"))
    layout = json.loads(str(sk["layout"]))
    gen = _generator_module()
    sd: Dict[str, torch.Tensor] = {}
    for name, shape, kind in layout:  # state_dict() order
        if kind == "param":
            sd[name] = gen.synthetic_tensor(name, torch.empty(shape), seed).float()
        elif kind == "buffer":
            sd[name] = torch.from_numpy(sk["buf:" + name].copy())
        else:  # tables: filled below
            sd[name] = torch.empty(0)
    g = GaussianTables.build()
    sd["gaussian_conditional.scale_table"] = g.scale_table.float()
    sd["gaussian_conditional._quantized_cdf"] = g.cdf.int()
    sd["gaussian_conditional._cdf_length"] = g.cdf_length.int()
    sd["gaussian_conditional._offset"] = g.offset.int()
    for k in ("_quantized_cdf", "_cdf_length", "_offset"):
        sd["entropy_bottleneck." + k] = torch.zeros(1, dtype=torch.int32)
    eb = BottleneckTables.from_state_dict(sd, "entropy_bottleneck")
    eb.rebuild()
    sd["entropy_bottleneck._quantized_cdf"] = eb.cdf.int()
    sd["entropy_bottleneck._cdf_length"] = eb.cdf_length.int()
    sd["entropy_bottleneck._offset"] = eb.offset.int()
    return sd


def _write(case: str, kwargs: dict) -> None:
    sys.path.insert(0, ROOT)
    from progressivecodec_b200 import ChannelProgresssiveWACNN

    net = ChannelProgresssiveWACNN(**kwargs).eval()
    params = {n for n, _ in net.named_parameters()}
    tables = {"_quantized_cdf", "_cdf_length", "_offset", "scale_table"}
    layout, arrays = [], {}
    for name, v in net.state_dict().items():
        leaf = name.rsplit(".", 1)[-1]
        if name in params:
            layout.append((name, list(v.shape), "param"))
        elif leaf in tables:
            layout.append((name, [], "table"))
        else:
            layout.append((name, list(v.shape), "buffer"))
            arrays["buf:" + name] = v.numpy()
    np.savez_compressed(os.path.join(GOLDEN, f"{case}_state_skeleton.npz"), layout=json.dumps(layout), **arrays)
    print(f"wrote {case}_state_skeleton.npz: {len(layout)} entries, {len(arrays)} buffers")


if __name__ == "__main__":
    if "--write" in sys.argv:
        _write("authors", dict(multiple_decoder=True, multiple_encoder=False, multiple_hyperprior=True, delta_encode=True,
                               support_progressive_slices=5, mask_policy="point-based-std"))
