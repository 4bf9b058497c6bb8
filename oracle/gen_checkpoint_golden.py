"""Golden key mappings for progressivecodec_b200.checkpoint, produced by the REAL reference functions
(utils/state_dict_handler.py replace_keys / complete_args, train.py initialize_model_from_pretrained).

The reference modules import packages that are absent here (pytorch_msssim, wandb, ...), so the three function
definitions are extracted from the source files with `ast` and executed on their own.  Run in the build container:
    python -m oracle.gen_checkpoint_golden        -> tests/golden/checkpoint_keys.json
"""
import argparse
import ast
import json
import os
from collections import OrderedDict

REF = "/root/reference/src"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "checkpoint_keys.json")


def _extract(path, names):
    src = open(path).read()
    tree = ast.parse(src)
    ns = {"OrderedDict": OrderedDict, "print": lambda *a, **k: None}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            exec(compile(ast.Module([node], []), path, "exec"), ns)
    return [ns[n] for n in names]


def main():
    replace_keys, complete_args = _extract(os.path.join(REF, "compress/utils/state_dict_handler.py"),
                                           ["replace_keys", "complete_args"])
    (init_pre,) = _extract(os.path.join(REF, "train.py"), ["initialize_model_from_pretrained"])
    old_multi = ["g_a.0.weight", "g_a.1.beta", "g_a_enh.0.weight", "g_a_enh.1.beta", "g_s.0.0.conv_a.0.conv.0.weight",
                 "h_a.0.weight", "cc_mean_transforms.0.0.weight", "entropy_bottleneck._matrices.0"]
    new_multi = ["g_a.0.0.weight", "g_a.0.1.beta", "g_a.1.0.weight", "g_s.0.0.conv_a.0.conv.0.weight", "h_a.0.weight"]
    wacnn = ["g_a.0.weight", "g_a.1.beta", "g_s.0.conv_a.0.conv.0.weight", "g_s.8.bias", "h_a.0.weight",
             "h_mean_s.0.weight", "h_scale_s.8.bias", "cc_mean_transforms.3.0.weight", "lrp_transforms.9.8.bias",
             "gaussian_conditional._offset", "entropy_bottleneck.quantiles"]
    enh = ["g_s.0.conv_a.0.conv.0.weight", "g_s.8.bias", "g_a.0.weight"]
    cases = {"replace_keys": [], "initialize_model_from_pretrained": [], "complete_args": []}
    for keys, me in ((old_multi, True), (new_multi, True), (old_multi, False)):
        ck = OrderedDict((k, i) for i, k in enumerate(keys))
        out = replace_keys(ck, me)
        cases["replace_keys"].append({"keys": keys, "multiple_encoder": me, "out": [[k, v] for k, v in out.items()]})
    for md, me, mh, use_enh in ((True, False, True, True), (True, True, False, False), (False, False, False, False)):
        a = argparse.Namespace(multiple_decoder=md, multiple_encoder=me, multiple_hyperprior=mh)
        ck = OrderedDict((k, i) for i, k in enumerate(wacnn))
        ce = OrderedDict((k, 100 + i) for i, k in enumerate(enh)) if use_enh else None
        out = init_pre(ck, a, ce)
        cases["initialize_model_from_pretrained"].append({"keys": wacnn, "enh": enh if use_enh else None,
                                                         "flags": [md, me, mh], "out": [[k, v] for k, v in out.items()]})
    for present in ([], ["multiple_encoder", "delta_encode"]):
        a = argparse.Namespace(**{k: True for k in present})
        out = complete_args(a)
        cases["complete_args"].append({"present": present, "out": dict(sorted(vars(out).items()))})
    with open(OUT, "w") as f:
        json.dump(cases, f, indent=1)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
