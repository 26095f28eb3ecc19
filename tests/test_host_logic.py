"""CPU-side tests: C ABI surface, host logic (tile tables, sharding), module construction parity, synth determinism.
No kernel is launched here (the container has no GPU)."""
import ctypes
import os
import re
import sys
import types
from importlib import import_module

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"


def _pkg(name):
    return import_module("multimodal-path-omic_b200." + name)


def test_library_loads_and_exports_every_declared_symbol():
    L = _pkg("_lib")
    lib = L.lib()
    header = open(os.path.join(ROOT, "include", "mpo_b200.h")).read()
    declared = set(re.findall(r"\b(mpo_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations found in the header"
    for name in declared:
        assert hasattr(lib, name), "libmpo_b200.so does not export %s" % name
    assert declared == set(L.exported_symbols()), declared ^ set(L.exported_symbols())
    assert lib.mpo_version() >= 100
    assert lib.mpo_sizeof(0) == ctypes.sizeof(L.MpoBag)
    assert lib.mpo_sizeof(1) == ctypes.sizeof(L.MpoModel)
    assert lib.mpo_sizeof(2) == ctypes.sizeof(L.MpoTailIo)


def test_workspace_layout_queries_do_not_need_a_gpu():
    L = _pkg("_lib")
    m = L.MpoModel()
    m.variant, m.fusion, m.n_classes = 1, 1, 4
    n32 = L.lib().mpo_tail_ws_floats(ctypes.byref(m), 32)
    n1 = L.lib().mpo_tail_ws_floats(ctypes.byref(m), 1)
    assert n32 > n1 > 0
    ln = ctypes.c_int64()
    off = L.lib().mpo_tail_ws_lookup(ctypes.byref(m), 32, b"path1_y2", ctypes.byref(ln))
    assert off > 0 and ln.value == 32 * 6 * 256
    assert L.lib().mpo_tail_ws_lookup(ctypes.byref(m), 32, b"no_such_buffer", ctypes.byref(ln)) == -1


def test_product_path_refuses_cpu_tensors():
    bp = _pkg("bagpass")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        bp.PackedBag(torch.zeros((4, 1024), dtype=torch.bfloat16), [4])
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        bp.cast_bf16(torch.zeros(8))
    loss = _pkg("loss").NegativeLogLikelihoodSurvivalLoss()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        loss(torch.rand(1, 4), torch.rand(1, 4), torch.tensor([[1]]), torch.tensor([0.0]))


def test_tile_table_ragged():
    bp = _pkg("bagpass")
    info, prefix = bp._tile_table_np((1, 129, 300, 128))
    assert prefix.tolist() == [0, 1, 3, 6, 7]
    assert info[:, 0].tolist() == [0, 1, 1, 2, 2, 2, 3]
    assert info[:, 1].tolist() == [0, 1, 129, 130, 258, 386, 430]
    assert info[:, 2].tolist() == [1, 128, 1, 128, 128, 44, 128]
    # every packed row is valid in exactly one tile
    covered = np.zeros(558, dtype=int)
    for _, r0, nv, _ in info:
        covered[r0:r0 + nv] += 1
    assert (covered == 1).all()


def test_sharding_helpers():
    dp = _pkg("dp")
    for world in (1, 2, 4, 8):
        seen = sorted(i for r in range(world) for i in dp.slide_shard(513, r, world))
        assert seen == list(range(513))
        for n in (1, 127, 128, 129, 200000, 16384):
            spans = [dp.patch_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
                assert a1 == b0 and a0 <= a1
            assert all(s % 128 == 0 for s, _ in spans if s < n)
    assert dp.patch_range(200000, 3, 8) == (3 * 25088, 4 * 25088)


def test_synth_is_deterministic_and_bf16_representable():
    synth = _pkg("synth")
    a = synth.make_slide(5, 64)
    b = synth.make_slide(5, 64)
    assert np.array_equal(a[0], b[0]) and all(np.array_equal(x, y) for x, y in zip(a[1], b[1]))
    assert np.array_equal(synth.bf16_round(a[0]), a[0])
    t = torch.from_numpy(a[0])
    assert torch.equal(t.bfloat16().float(), t)
    st = synth.make_state({"H.0.weight": (256, 1024), "H.0.bias": (256,)}, 3)
    assert np.array_equal(synth.bf16_round(st["H.0.weight"]), st["H.0.weight"])
    x = np.random.default_rng(0).standard_normal(1000).astype(np.float32)
    assert np.array_equal(synth.bf16_round(x), torch.from_numpy(x).bfloat16().float().numpy())


def test_constructor_contract():
    mcat = _pkg("mcat").MultimodalCoAttentionTransformer
    with pytest.raises(RuntimeError, match="not implemented"):
        mcat(omic_sizes=[10] * 6, fusion="nope")
    net = mcat(omic_sizes=[100, 200, 300, 400, 500, 600], model_size="small")
    assert net.get_trainable_parameters() > 0
    with pytest.raises(NotImplementedError, match="medium"):
        net._engine
    with pytest.raises(RuntimeError, match="only on 2 inputs"):
        _pkg("fusion").BilinearFusion()(torch.zeros(256))


@pytest.mark.skipif(not os.path.isdir(REF), reason="the reference checkout only exists in the build container")
@pytest.mark.parametrize("fusion", ["concat", "bilinear", "gated_concat"])
def test_state_dict_and_initialisation_match_the_reference(fusion):
    sys.modules.setdefault("h5py", types.ModuleType("h5py"))
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import warnings
    warnings.filterwarnings("ignore")
    from models.mcat.mcat import MultimodalCoAttentionTransformer as RefM
    from models.nacagat.nacagat import NarrowContextualAttentionGateTransformer as RefN
    pairs = [(_pkg("mcat").MultimodalCoAttentionTransformer, RefM),
             (_pkg("nacagat").NarrowContextualAttentionGateTransformer, RefN)]
    sizes = [100, 200, 300, 400, 500, 600]
    for mine, ref in pairs:
        torch.manual_seed(7)
        a = mine(omic_sizes=sizes, fusion=fusion)
        torch.manual_seed(7)
        b = ref(omic_sizes=sizes, fusion=fusion)
        sa, sb = a.state_dict(), b.state_dict()
        assert list(sa.keys()) == list(sb.keys())
        assert all(torch.equal(sa[k], sb[k]) for k in sa)
        assert a.get_trainable_parameters() == b.get_trainable_parameters()
        b.load_state_dict(sa)      # checkpoints travel both ways


def test_parameter_map_follows_reassigned_parameters_and_modules():
    """ModelBinding.params(refresh=True) reads the parameters through a cached index instead of named_parameters();
    it must still see a parameter or a submodule that was replaced after the index was built."""
    synth = _pkg("synth")
    net = _pkg("mcat").MultimodalCoAttentionTransformer(omic_sizes=list(synth.OMIC_SIZES))
    bnd = net._engine.binding

    def same(P):
        ref = dict(net.named_parameters())
        assert list(P) == list(ref)
        assert all(P[n] is ref[n] for n in ref)

    same(bnd.params(refresh=True))
    same(bnd.params(refresh=True))                       # second call goes through the index
    assert bnd._index                                    # and the index is in use
    net.H[0].weight = torch.nn.Parameter(torch.zeros_like(net.H[0].weight))
    same(bnd.params(refresh=True))
    net.classifier = torch.nn.Linear(net.classifier.in_features, net.classifier.out_features)
    same(bnd.params(refresh=True))
    net.load_state_dict({k: v.clone() for k, v in net.state_dict().items()})
    same(bnd.params(refresh=True))


def test_data_parallel_replica_binds_to_the_module_that_owns_the_parameters():
    """models/mcat/main.py:267-268 wraps the model in nn.DataParallel whenever the box has several GPUs;
    DataParallel.replicate() builds replicas whose _parameters dicts are empty.  A replica must resolve to the engine of
    the module that owns the leaf parameters."""
    synth = _pkg("synth")
    for cls in (_pkg("mcat").MultimodalCoAttentionTransformer, _pkg("nacagat").NarrowContextualAttentionGateTransformer):
        net = cls(omic_sizes=list(synth.OMIC_SIZES))
        # what torch.nn.parallel.replicate() does on one device: every module is shallow-copied with an empty
        # _parameters dict, children re-linked, the weights re-attached as plain (non-leaf) tensors
        mods = list(net.modules())
        idx = {m: i for i, m in enumerate(mods)}
        reps = [m._replicate_for_data_parallel() for m in mods]
        for m, r in zip(mods, reps):
            for key, child in m._modules.items():
                r._modules[key] = None if child is None else reps[idx[child]]
            for key, p in m._parameters.items():
                if p is not None:
                    setattr(r, key, p.detach().clone().requires_grad_())     # nn.Module.__setattr__: a plain tensor
        replica = reps[0]
        assert not list(replica.named_parameters())
        eng = replica._engine
        assert eng is net._engine and eng.binding.module is net
        assert eng.binding.names == [n for n, _ in net.named_parameters()]
        P = eng.binding.params(refresh=True)
        assert all(P[n] is p for n, p in net.named_parameters())
    ge = _pkg("ge_nacagat").GeneExprNarrowContextualAttentionGateTransformer()
    assert ge._replicate_for_data_parallel()._master_ref[0] is ge


def test_deepcopy_and_pickle_drop_the_engine_and_rebind():
    import copy
    import pickle
    synth = _pkg("synth")
    net = _pkg("mcat").MultimodalCoAttentionTransformer(omic_sizes=list(synth.OMIC_SIZES))
    eng = net._engine
    eng.binding.params(refresh=True)
    for clone in (copy.deepcopy(net), pickle.loads(pickle.dumps(net))):
        assert clone._engine_obj is None and clone._master_ref[0] is clone
        assert clone._engine is not eng and clone._engine.binding.module is clone


def test_batch_trainer_rebinds_gradient_views_after_zero_grad_set_to_none():
    """optimizer.zero_grad() (set_to_none=True by default, the call in models/mcat/main.py:72-74) drops the .grad views
    of the flat gradient buffer; rebind_grad_views() -- run by BatchTrainer.step() / zero_grad() and GraphedStep.replay()
    -- restores them."""
    synth = _pkg("synth")
    sp = _pkg("slidepath")
    net = _pkg("mcat").MultimodalCoAttentionTransformer(omic_sizes=list(synth.OMIC_SIZES))
    P = dict(net.named_parameters())
    flat = torch.full((sum(p.numel() for p in P.values()),), 3.0)
    views, off = {}, 0
    for n, p in P.items():
        views[n] = flat[off:off + p.numel()].view_as(p)
        p.grad = views[n]
        off += p.numel()
    opt = torch.optim.SGD(net.parameters(), lr=0.1)
    opt.zero_grad()                                      # set_to_none=True
    assert all(p.grad is None for p in net.parameters())
    net.classifier.bias.grad = torch.full_like(net.classifier.bias, 7.0)     # a foreign gradient tensor
    sp.rebind_grad_views(P, views)
    for n, p in P.items():
        assert p.grad is views[n]
        want = 7.0 if n == "classifier.bias" else 0.0
        assert float(p.grad.max()) == want and float(p.grad.min()) == want
    net.zero_grad(set_to_none=False)
    sp.rebind_grad_views(P, views)
    assert float(flat.abs().max()) == 0.0 and all(p.grad is views[n] for n, p in P.items())
