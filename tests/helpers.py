"""Shared test helpers: rebuild a golden case (weights + slide) from synth, load fixtures, compare digests."""
import ast
import glob
import os
from importlib import import_module

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
synth = import_module("multimodal-path-omic_b200.synth")


def golden_cases(unrounded=False):
    """MCAT / NaCAGaT fixtures.  The `unrounded` ones hold reference results for weights that are not
    bf16-representable: the oracle must match them exactly, the CUDA path only within the reported rounding delta."""
    names = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz"))
                   if not p.endswith("loss_known_answers.npz") and not os.path.basename(p).startswith(("ge_", "alt_")))
    return [n for n in names if ("unrounded" in n) == unrounded]


def ge_cases():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "ge_*.npz")))


def load_ge_case(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    n, seed, label, _, sharpen = z["meta"]
    names = [str(s) for s in z["param_names"]]
    shapes = {k: ast.literal_eval(str(s)) for k, s in zip(names, z["param_shapes"])}
    state = synth.make_state(shapes, int(seed), model="ge", sharpen=float(sharpen))
    bag, _, _, _ = synth.make_slide(int(seed), int(n))
    return dict(name=name, model="ge", n=int(n), seed=int(seed), label=int(label), state=state, bag=bag, gold=z,
                param_names=names)


def alt_cases():
    """fixtures of the other loss / fusion branches (sct, cesar, gated_concat): gradient digests of `loss_kind`."""
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "alt_*.npz")))


def load_case(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    n, seed, label, censor, sharpen = z["meta"]
    parts = name[4:].split("_") if name.startswith("alt_") else name.split("_")
    model, fusion = parts[0], parts[1]
    if fusion == "gated":
        fusion = "gated_concat"
    names = [str(s) for s in z["param_names"]]
    shapes = {k: ast.literal_eval(str(s)) for k, s in zip(names, z["param_shapes"])}
    state = synth.make_state(shapes, int(seed), model=model, sharpen=float(sharpen),
                             round_bag_weights="unrounded" not in name)
    bag, omics, lab, cen = synth.make_slide(int(seed), int(n))
    assert lab == int(label) and cen == float(censor)
    return dict(name=name, model=model, fusion=fusion, n=int(n), seed=int(seed), label=lab, censor=cen,
                state=state, bag=bag, omics=omics, gold=z, param_names=names)


def digest_errors(case, grads):
    """Per-parameter gradient check against the stored digests.

    Returns (worst norm-relative error over parameters whose reference gradient is not noise, details).
    A gradient counts as noise when its norm is < 1e-5 of the largest parameter-gradient norm of the case
    (several reference gradients are ~1e-9, e.g. the key bias that cancels in the softmax: SURVEY F3/F7)."""
    gold = case["gold"]
    norms = {k: float(gold["gd/" + k][0]) for k in case["param_names"]}
    gmax = max(norms.values())
    worst, details = 0.0, []
    for k in case["param_names"]:
        ref = gold["gd/" + k]
        got = synth.grad_digest(k, grads[k])
        floor = 1e-5 * gmax
        # projection and samples are compared relative to the tensor's norm (with the noise floor)
        scale = max(ref[0], floor)
        e_norm = abs(got[0] - ref[0]) / scale
        e_rest = float(np.max(np.abs(got[1:] - ref[1:]))) / scale
        err = max(e_norm, e_rest)
        details.append((k, ref[0], err))
        worst = max(worst, err)
    return worst, details
