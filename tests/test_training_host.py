"""Host logic of the training drivers and the bag ingestion (no GPU): c-index, window grouping, slide files."""
import itertools
import os
from importlib import import_module

import numpy as np
import pytest
import torch


def _pkg(name):
    return import_module("multimodal-path-omic_b200." + name)


def _brute_c_index(event, time, risk):
    conc = tied = comp = 0
    n = len(time)
    for i, j in itertools.permutations(range(n), 2):
        if not event[i]:
            continue
        if time[i] < time[j] or (time[i] == time[j] and not event[j]):
            comp += 1
            if abs(risk[i] - risk[j]) <= 1e-8:
                tied += 1
            elif risk[i] > risk[j]:
                conc += 1
    return (conc + 0.5 * tied) / comp


def test_concordance_index_matches_pairwise_definition():
    tr = _pkg("training")
    rng = np.random.default_rng(0)
    for n in (5, 17, 60):
        event = rng.integers(0, 2, n).astype(bool)
        event[0] = True
        time = rng.integers(1, 12, n).astype(float)          # many tied times
        risk = np.round(rng.standard_normal(n), 1)           # some tied risks
        assert abs(tr.concordance_index(event, time, risk) - _brute_c_index(event, time, risk)) < 1e-12
    # perfectly ordered risks: higher risk dies first
    assert tr.concordance_index([1, 1, 1], [1.0, 2.0, 3.0], [3.0, 2.0, 1.0]) == 1.0
    with pytest.raises(ValueError):
        tr.concordance_index([0, 0], [1.0, 2.0], [0.1, 0.2])


def test_iterate_windows_and_sample_adapter():
    ing, tr = _pkg("ingest"), _pkg("training")
    assert [len(w) for w in ing.iterate_windows(range(7), 3)] == [3, 3, 1]
    item = (12.5, 2, 1.0, [torch.zeros(3)], torch.zeros(4, 1024))
    s = tr._as_sample(item)
    assert s["months"] == 12.5 and s["label"] == 2 and s["censor"] == 1.0 and s["bag"].shape == (4, 1024)


def test_slide_file_source_pt_and_npy(tmp_path):
    ing = _pkg("ingest")
    a = torch.randn(5, 1024)
    torch.save(a, tmp_path / "TCGA-01.pt")
    torch.save(a.unsqueeze(0), tmp_path / "TCGA-02.pt")
    src = ing.SlideFileSource(str(tmp_path))
    assert src.kind == "pt" and src.has("TCGA-01.svs") and not src.has("TCGA-09.svs")      # dataset.py:125 naming
    assert torch.equal(src.load("TCGA-01.svs"), a) and src.load("TCGA-02").shape == (5, 1024)
    d2 = tmp_path / "npy"
    os.makedirs(d2)
    np.save(d2 / "S1.npy", a.numpy())
    assert torch.equal(ing.SlideFileSource(str(d2)).load("S1"), a)
    torch.save(torch.randn(5, 7), tmp_path / "bad.pt")
    with pytest.raises(RuntimeError, match="expected"):
        src.load("bad")
    dst = torch.empty((5, 1024), dtype=torch.bfloat16)
    ing.to_bf16_into(dst, a)
    assert torch.equal(dst, a.to(torch.bfloat16))
    try:
        import h5py  # noqa: F401
    except ImportError:
        with pytest.raises(ImportError, match="h5py"):
            ing.SlideFileSource(str(tmp_path / "bags.h5"), kind="h5")
