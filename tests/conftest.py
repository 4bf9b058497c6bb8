import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    """The CUDA library must exist before anything imports it; (re)build is a no-op when up to date."""
    from progressivecodec_b200 import _lib

    if not os.path.isfile(_lib.LIB_PATH):
        _lib.build_library()
    from oracle import entropy_port

    entropy_port.build_c_port()
    yield


CASE_KWARGS = {
    "authors": dict(multiple_decoder=True, multiple_encoder=False, multiple_hyperprior=True, delta_encode=True,
                    support_progressive_slices=5, mask_policy="point-based-std"),
    "multienc": dict(multiple_decoder=True, multiple_encoder=True, multiple_hyperprior=True, delta_encode=True,
                     support_progressive_slices=5, mask_policy="point-based-std"),
    "allscalable": dict(multiple_decoder=True, multiple_encoder=False, multiple_hyperprior=True, delta_encode=True,
                        support_progressive_slices=5, mask_policy="point-based-std", all_scalable=True),
    "plain": dict(multiple_decoder=True, multiple_encoder=True, multiple_hyperprior=False, delta_encode=False,
                  support_progressive_slices=0, mask_policy="two-levels"),
}


def load_golden(name):
    import numpy as np

    return np.load(os.path.join(GOLDEN, f"{name}.npz"))


_MODEL_CACHE = {}


def build_pair(name, device=None):
    """(our model with synthetic weights [on `device`], oracle built from the same state dict)."""
    import torch

    from oracle.codec_port import CodecConfig, OracleCodec
    from progressivecodec_b200 import ChannelProgresssiveWACNN, apply_synthetic_weights

    if name not in _MODEL_CACHE:
        kw = CASE_KWARGS[name]
        net = ChannelProgresssiveWACNN(**kw).eval()
        apply_synthetic_weights(net, seed=0)
        net.update(force=True)
        orc = OracleCodec(net.state_dict(), CodecConfig(**kw))
        _MODEL_CACHE[name] = (net, orc)
    net, orc = _MODEL_CACHE[name]
    if device is not None:
        net = net.to(device)
        _MODEL_CACHE[name] = (net, orc)
    return net, orc
