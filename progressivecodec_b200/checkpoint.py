"""Checkpoint ingestion: the reference's state-dict key handlers, so the authors' ``.pth.tar`` files load directly.

    replace_keys / complete_args      utils/state_dict_handler.py:10-27, 76-83
    initialize_model_from_pretrained  train.py:27-84   (start a progressive model from a single-rate WACNN checkpoint)
    load_checkpoint                   the evaluation entry's sequence: torch.load -> complete_args -> build -> replace_keys
                                      -> load_state_dict -> update()
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Any, Dict, Optional

import torch

from .models import ChannelProgresssiveWACNN


def replace_keys(checkpoint: Dict[str, Any], multiple_encoder: bool) -> "OrderedDict[str, Any]":
    """utils/state_dict_handler.py:10-27: old multi-encoder checkpoints name the two analysis transforms `g_a.` and
    `g_a_enh.`; the module tree calls them `g_a.0.` and `g_a.1.`."""
    out: "OrderedDict[str, Any]" = OrderedDict()
    already_indexed = "g_a.0.1.beta" in checkpoint
    for key, value in checkpoint.items():
        if multiple_encoder:
            if "g_a_enh." in key:
                out[key.replace("g_a_enh.", "g_a.1.")] = value
            elif "g_a." in key and not already_indexed:
                out[key.replace("g_a.", "g_a.0.")] = value
            else:
                out[key] = value
        else:
            out[key] = value
    return out


def complete_args(new_args):
    """utils/state_dict_handler.py:76-83: flags that older checkpoints' argparse namespaces lack default to False."""
    for name in ("multiple_encoder", "multiple_hyperprior", "delta_encode", "residual_before_lrp", "double_dim"):
        if name not in new_args:
            setattr(new_args, name, False)
    return new_args


def initialize_model_from_pretrained(checkpoint: Dict[str, Any], args, checkpoint_enh: Optional[Dict[str, Any]] = None):
    """train.py:27-84: map a single-rate WACNN state dict onto the progressive model's module tree (g_s -> g_s.0,
    g_a -> g_a.0, hyper-synthesis -> h_*_s.0 when multiple_hyperprior; h_a is dropped; an optional second checkpoint
    provides g_s.1)."""
    sub: "OrderedDict[str, Any]" = OrderedDict()
    for c in list(checkpoint.keys()):
        if "g_s" in c:
            sub["g_s.0." + c[4:] if args.multiple_decoder else c] = checkpoint[c]
        elif "g_a" in c:
            sub["g_a.0." + c[4:] if args.multiple_encoder else c] = checkpoint[c]
        else:  # the reference's `elif "cc_" in c or "lrp_" in c or "gaussian_conditional" or ...` is always true
            sub[c] = checkpoint[c]
    for c in list(sub.keys()):
        if "h_scale_s" in c or "h_a" in c or "h_mean_s" in c:
            sub.pop(c)
    if args.multiple_hyperprior:
        for c in list(checkpoint.keys()):
            if "h_mean_s" in c:
                sub["h_mean_s.0." + c[9:]] = checkpoint[c]
            elif "h_scale_s" in c:
                sub["h_scale_s.0." + c[10:]] = checkpoint[c]
    if checkpoint_enh is not None:
        for c in list(checkpoint_enh.keys()):
            if "g_s" in c:
                sub["g_s.1." + c[4:]] = checkpoint_enh[c]
    return sub


_CTOR_KEYS = ("N", "M", "multiple_decoder", "multiple_encoder", "multiple_hyperprior", "dim_chunk", "division_dimension",
              "mask_policy", "joiner_policy", "support_progressive_slices", "double_dim", "delta_encode",
              "residual_before_lrp", "support_std", "total_mu_rep", "all_scalable")


def load_checkpoint(path_or_dict, device="cuda", lmbda_list=None) -> ChannelProgresssiveWACNN:
    """Build the B200 model from a reference checkpoint ({"state_dict": ..., "args": argparse.Namespace}), exactly as the
    reference's evaluation entry does: complete_args, get_model(args) (models/__init__.py:49-68), replace_keys,
    load_state_dict, update()."""
    ckpt = torch.load(path_or_dict, map_location="cpu", weights_only=False) if isinstance(path_or_dict, str) else path_or_dict
    args = complete_args(ckpt["args"])
    kw = {k: getattr(args, k) for k in _CTOR_KEYS if k in args}
    if lmbda_list is None:
        lmbda_list = getattr(args, "lmbda_list", [0.005, 0.05])
    net = ChannelProgresssiveWACNN(lmbda_list=lmbda_list, **kw)
    net.load_state_dict(replace_keys(ckpt["state_dict"], bool(getattr(args, "multiple_encoder", False))), strict=True)
    net.update()
    return net.eval().to(device)
