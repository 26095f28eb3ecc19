"""Deterministic synthetic inputs and random-init-scale weights (numpy PCG64: identical on every machine).

Used by the golden-vector generator (run against the reference in the build container), by the parity tests
(run anywhere) and by bench.py, so that all of them see bit-identical slides and weights without shipping
16 MB of parameters per fixture.  Bag-facing weights (H.0.weight and, for NaCAGaT, the key block of
co_attention.in_proj_weight) are rounded to bf16-representable values on BOTH sides (SURVEY.md H2 option a):
the kernels stream them as bf16, the reference and the oracle get the same values in fp32.
"""
import numpy as np

OMIC_SIZES = (100, 200, 300, 400, 500, 600)     # reference: models/mcat/mcat.py:152


def bf16_round(a):
    """round-to-nearest-even fp32 -> bf16 -> fp32, in numpy."""
    a = np.ascontiguousarray(a, dtype=np.float32)
    u = a.view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(np.float32).reshape(a.shape)


def _is_norm_weight(name):
    return name.endswith(("norm1.weight", "norm2.weight", ".G.1.weight", ".E.1.weight"))


def make_state(shapes, seed, model="mcat", sharpen=1.0, round_bag_weights=True):
    """shapes: ordered {state_dict key: shape}.  Returns {key: float32 ndarray}.  round_bag_weights=False leaves the
    bag-facing weights un-rounded (the "unrounded" fixtures that measure what the kernels' own bf16 rounding costs)."""
    out = {}
    for idx, (name, shape) in enumerate(shapes.items()):
        shape = tuple(int(s) for s in shape)
        rng = np.random.default_rng([int(seed), idx])
        if len(shape) >= 2:
            bound = 1.0 / np.sqrt(shape[1])
            w = rng.uniform(-bound, bound, size=shape)
        elif _is_norm_weight(name):
            w = 1.0 + rng.uniform(-0.1, 0.1, size=shape)
        else:
            w = rng.uniform(-0.05, 0.05, size=shape)
        w = w.astype(np.float32)
        if name == "co_attention.in_proj_weight" or name == "self_attention.in_proj_weight":
            w = (w * np.float32(sharpen)).astype(np.float32)
        if name == "H.0.weight" and round_bag_weights:
            w = bf16_round(w)
        if name == "co_attention.in_proj_weight" and model == "nacagat" and round_bag_weights:
            e = shape[1]
            w[e:2 * e] = bf16_round(w[e:2 * e])
        out[name] = w
    return out


def make_slide(seed, n_patches, omic_sizes=OMIC_SIZES):
    """One synthetic slide: bf16-representable bag [N,1024] (fp32 array), omics list, label, censorship."""
    rng = np.random.default_rng([int(seed), 7919])
    bag = bf16_round(rng.standard_normal((int(n_patches), 1024), dtype=np.float32))
    omics = [rng.standard_normal(int(d), dtype=np.float32) for d in omic_sizes]
    label = int(seed) % 4
    censor = float(int(seed) % 2)
    return bag, omics, label, censor


def grad_digest(name, g, seed=0):
    """Compact, order-sensitive summary of one gradient tensor: L2 norm, a random +-1 projection, 16 samples."""
    g = np.asarray(g, dtype=np.float64).reshape(-1)
    rng = np.random.default_rng([int(seed), len(name), g.size])
    sign = rng.integers(0, 2, size=g.size) * 2.0 - 1.0
    idx = np.linspace(0, g.size - 1, num=min(16, g.size)).astype(np.int64)
    return np.concatenate([[np.linalg.norm(g)], [float(sign @ g)], g[idx]])
