"""progressivecodec_b200 — B200-native (sm_100a) inference hot path of EIDOSLAB/ProgressiveCodec.

Public surface mirrors the reference:
  * ``ChannelProgresssiveWACNN``            (compress/models/CHProg_cnn.py:30)
  * ``EntropyBottleneck``, ``GaussianConditional``, ``EntropyModel``  (compress/entropy_models)
  * ``ans.RansEncoder / RansDecoder / BufferedRansEncoder``           (compressai.ans)
  * ``pmf_to_quantized_cdf``                                          (compressai._CXX)
All arithmetic runs in libpcodec_b200.so (C-ABI: include/pcodec_b200.h); there is no CPU fallback.
"""
from . import _lib, ans, checkpoint, container, evaluation, pipeline  # noqa: F401
from ._lib import PcodecError, build_library  # noqa: F401
from .entropy_models import EntropyBottleneck, EntropyModel, GaussianConditional, pmf_to_quantized_cdf  # noqa: F401
from .models import ChannelProgresssiveWACNN, get_scale_table  # noqa: F401
from .rem import PostRateProcessedNetwork  # noqa: F401
from .synthetic import apply_synthetic_weights  # noqa: F401

models = {"channel": ChannelProgresssiveWACNN}

__version__ = "0.1.0"
