"""Slide engine: drives the C-ABI stages for a batch of slides (pre -> bag forward -> post -> loss -> post
backward -> bag backward -> pre backward) on behalf of the drop-in modules.

Stage order mirrors the reference forward (models/mcat/mcat.py:84-142) and its autograd graph.  PyTorch's role here
is limited to owning device memory / streams and, for the per-slide module API, to carrying the gradients back
into nn.Parameter.grad through one torch.autograd.Function.
"""
import ctypes
import itertools
import os

import torch

from . import _lib
from . import bagpass as bp
from .bagpass import D, Q, _ptr, _stream, require_cuda

VARIANT_MCAT, VARIANT_NACAGAT = 0, 1
FUSION_CONCAT, FUSION_BILINEAR, FUSION_GATED_CONCAT = 0, 1, 2
_FUSION_CODE = {"concat": FUSION_CONCAT, "bilinear": FUSION_BILINEAR, "gated_concat": FUSION_GATED_CONCAT}

_seed_counter = itertools.count(1)


def _rank_salt():
    """ranks seeded identically (the usual torch.manual_seed(0) on every rank) must still draw different dropout
    masks for their different slides (rank-distinct offsets)."""
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank() + 1
    except Exception:
        pass
    return 0


def _next_seed():
    return (torch.initial_seed() * 2654435761 + next(_seed_counter) * 40503 + _rank_salt() * 0x9E3779B1) & 0xFFFFFFFF


def _lin_names(prefix):
    return prefix + ".weight", prefix + ".bias"


class ModelBinding:
    """Maps a module's parameters onto struct mpo_model.  grads: {param name: tensor} or None (inference)."""

    def __init__(self, module, variant, fusion, omic_sizes, n_classes):
        if fusion not in _FUSION_CODE:
            raise NotImplementedError("fusion=%r is not implemented on the B200 path" % fusion)
        if len(omic_sizes) != Q:
            raise NotImplementedError("the B200 kernels are built for 6 omic signature groups, got %d" % len(omic_sizes))
        self.module = module
        self.variant = variant
        self.fusion = _FUSION_CODE[fusion]
        self.omic_sizes = [int(d) for d in omic_sizes]
        self.n_classes = int(n_classes)
        self.names = [n for n, _ in module.named_parameters()]

    def _build_index(self):
        """(full name, owning module's _parameters dict, key) per parameter, in named_parameters() order, plus the
        (parent _modules dict, key, child) links that prove the module tree is still the one that was indexed."""
        index, links = [], []
        for mod_name, mod in self.module.named_modules():
            for key, p in mod._parameters.items():
                if p is not None:
                    index.append(((mod_name + "." if mod_name else "") + key, mod._parameters, key))
            for key, child in mod._modules.items():
                if child is not None:
                    links.append((mod._modules, key, child))
        if [e[0] for e in index] != self.names:      # shared / re-registered parameters: keep the plain traversal
            return None, None
        return index, links

    def params(self, refresh=False):
        """name -> nn.Parameter.  module.named_parameters() costs ~0.3-0.5 ms per traversal, which was a quarter of the
        per-slide call; the refresh (once per call, in run_slide) instead re-reads every parameter from its owning
        module's _parameters dict through a cached index (~100 dict lookups).  The index is rebuilt when a submodule
        has been replaced; a parameter re-assigned in place of an old one is picked up by the lookup itself."""
        if refresh or getattr(self, "_P", None) is None:
            index = getattr(self, "_index", None)          # None: not built yet; False: do not use an index
            if index and not all(d.get(k) is c for d, k, c in self._links):
                index = None
            if index is None:
                index, self._links = self._build_index()
                self._index = index = index if index else False
            P = None
            if index:
                P = {name: d.get(k) for name, d, k in index}
                if any(p is None for p in P.values()):      # a parameter was deleted or set to None: re-index next time
                    self._index, P = None, None
            self._P = P if P is not None else dict(self.module.named_parameters())
        return self._P

    def build(self, grads=None):
        """struct mpo_model over the current parameter (and gradient) addresses.  The struct of the last call is kept
        per mode (with / without gradients) and reused while every address is unchanged: filling ~200 pointers
        through ctypes and re-validating 100 tensors was 0.17 ms per call, twice per slide step."""
        P = self.params()
        key = (tuple([p.data_ptr() for p in P.values()]),
               None if grads is None else tuple([g.data_ptr() for g in grads.values()]))
        cache = self.__dict__.setdefault("_built", {})
        hit = cache.get(grads is None)
        if hit is not None and hit[0] == key:
            return hit[1]
        m = self._build(P, grads)
        cache[grads is None] = (key, m)
        return m

    def _build(self, P, grads):
        for n, p in P.items():
            if not p.is_cuda:
                raise RuntimeError("parameter %s is on %s: move the model to a CUDA device (no CPU fallback)" % (n, p.device))
            if p.dtype != torch.float32 or not p.is_contiguous():
                raise RuntimeError("parameter %s must be contiguous float32" % n)
        keep = []

        def lin(prefix, wname=None, bname=None):
            wn, bn = (wname, bname) if wname else _lin_names(prefix)
            L = _lib.MpoLin()
            L.w = P[wn].data_ptr()
            L.b = P[bn].data_ptr()
            if grads is not None:
                L.gw = grads[wn].data_ptr()
                L.gb = grads[bn].data_ptr()
            return L

        def norm(prefix):
            N = _lib.MpoNorm()
            N.g = P[prefix + ".weight"].data_ptr()
            N.b = P[prefix + ".bias"].data_ptr()
            if grads is not None:
                N.gg = grads[prefix + ".weight"].data_ptr()
                N.gb = grads[prefix + ".bias"].data_ptr()
            return N

        def enc_layer(prefix):
            E = _lib.MpoEncoderLayer()
            E.in_proj = lin(None, prefix + ".self_attn.in_proj_weight", prefix + ".self_attn.in_proj_bias")
            E.out_proj = lin(prefix + ".self_attn.out_proj")
            E.linear1 = lin(prefix + ".linear1")
            E.linear2 = lin(prefix + ".linear2")
            E.norm1 = norm(prefix + ".norm1")
            E.norm2 = norm(prefix + ".norm2")
            return E

        def pool(head, rho):
            H = _lib.MpoPoolHead()
            H.att_a = lin(head + ".attention_a.0")
            H.att_b = lin(head + ".attention_b.0")
            H.att_c = lin(head + ".attention_c")
            H.rho = lin(rho + ".0")
            return H

        m = _lib.MpoModel()
        m.variant = self.variant
        m.fusion = self.fusion
        m.n_classes = self.n_classes
        for i, d in enumerate(self.omic_sizes):
            m.omic_dims[i] = d
            m.snn[i][0] = lin("G.%d.0.0" % i)
            m.snn[i][1] = lin("G.%d.1.0" % i)
        m.H = lin("H.0")
        m.coattn_in = lin(None, "co_attention.in_proj_weight", "co_attention.in_proj_bias")
        m.coattn_out = lin("co_attention.out_proj")
        if self.variant == VARIANT_NACAGAT:
            c = _lib.MpoCag()
            c.fc1 = lin("co_attention.CAG.fc1.0")
            c.fc2 = lin("co_attention.CAG.fc2.0")
            c.fc3 = lin("co_attention.CAG.fc3.0")
            c.fc_c = lin("co_attention.CAG.fc_c.0")
            c.G = norm("co_attention.CAG.G.1")
            c.E = norm("co_attention.CAG.E.1")
            m.cag = c
        for l in range(2):
            m.path_tr[l] = enc_layer("path_transformer.layers.%d" % l)
            m.omic_tr[l] = enc_layer("omic_transformer.layers.%d" % l)
        m.path_pool = pool("path_attention_head", "path_rho")
        m.omic_pool = pool("omic_attention_head", "omic_rho")
        if self.fusion in (FUSION_CONCAT, FUSION_GATED_CONCAT):
            m.fusion0 = lin("fusion_layer.fusion_layer.0")
            m.fusion2 = lin("fusion_layer.fusion_layer.2")
            if self.fusion == FUSION_GATED_CONCAT:
                # the reference keeps the two gate layers in a plain Python list (fusion.py:25-27): they are not
                # parameters of the module, are never trained and do not follow .to(device); device copies are made here
                dev = next(iter(P.values())).device
                for i, gate in enumerate(self.module.fusion_layer.gates):
                    gw = gate[0].weight.detach().to(device=dev, dtype=torch.float32).contiguous()
                    gb = gate[0].bias.detach().to(device=dev, dtype=torch.float32).contiguous()
                    keep.append((gw, gb))
                    m.gate[i].w, m.gate[i].b = gw.data_ptr(), gb.data_ptr()
        else:
            b = _lib.MpoBilinear()
            b.h1 = lin("fusion_layer.linear_h1.0")
            b.z1 = lin("fusion_layer.linear_z1")
            b.o1 = lin("fusion_layer.linear_o1.0")
            b.h2 = lin("fusion_layer.linear_h2.0")
            b.z2 = lin("fusion_layer.linear_z2")
            b.o2 = lin("fusion_layer.linear_o2.0")
            b.fc1 = lin("fusion_layer.fc1.0")
            b.fc2 = lin("fusion_layer.fc2.0")
            m.bil = b
        m.classifier = lin("classifier")
        keep.append(P)
        keep.append(grads)
        m._keepalive = keep
        return m


class SlideState:
    """Everything one forward pass leaves behind for the backward pass."""
    pass


def _no_engine():
    return None


class SlideEngine:
    # an engine caches device buffers and ctypes structs over one module's parameter addresses: a copy.deepcopy() or
    # pickle of the owning module drops it, and the copy builds its own on first use
    def __deepcopy__(self, memo):
        return None

    def __reduce__(self):
        return (_no_engine, ())

    def __init__(self, binding, bag_dropout=0.25, attn_dropout=0.25):
        self.binding = binding
        self.bag_dropout = float(bag_dropout)
        self.attn_dropout = float(attn_dropout)     # NaCAGaT: PreGatingContextualAttention(dropout_p=0.25), blocks.py:52
        self.tail_dropout = float(bag_dropout)      # the model's `dropout`: SNN / encoder / rho layers (mcat.py:13,38-66)
        self._ws_cache = {}
        self._w_bf16 = None
        self._wk_f16 = None

    # -- helpers
    def _tail_ws(self, model, B, device, reuse):
        """fresh workspace per pass unless the caller guarantees forward/backward alternate (BatchTrainer)."""
        key = (B, str(device))
        if reuse and key in self._ws_cache:
            return self._ws_cache[key]
        n = _lib.lib().mpo_tail_ws_floats(ctypes.byref(model), B)
        if n <= 0:
            raise RuntimeError("mpo_tail_ws_floats failed")
        ws = torch.empty(n, dtype=torch.float32, device=device)
        if reuse:
            self._ws_cache = {key: ws}   # keep one size around
        return ws

    def ws_view(self, model, st, name):
        """named intermediate of the tail workspace (tests / debugging)."""
        n = ctypes.c_int64()
        off = _lib.lib().mpo_tail_ws_lookup(ctypes.byref(model), st.B, name.encode(), ctypes.byref(n))
        if off < 0:
            raise KeyError(name)
        return st.tail_ws[off:off + n.value]

    def _io(self, st):
        io = _lib.MpoTailIo()
        io.num_slides = st.B
        for i in range(Q):
            io.omics[i] = st.omics[i].data_ptr()
        io.ws = st.tail_ws.data_ptr()
        io.qp = st.qp.data_ptr()
        io.qk = st.qk.data_ptr()
        io.kc = st.kc.data_ptr() if st.kc is not None else None
        io.pooled = st.bag_ws.pooled.data_ptr()
        # sum_n a'_in differs from 1 only under attention dropout (NaCAGaT, train mode)
        use_suma = st.bag_ws.nacagat and getattr(st, "attn_p", 0.0) > 0.0
        io.suma = st.bag_ws.suma.data_ptr() if use_suma else None
        io.dsuma = st.dsuma.data_ptr() if (use_suma and st.dsuma is not None) else None
        io.dkc = st.dkc.data_ptr() if st.dkc is not None else None
        io.dtq = st.dtq.data_ptr() if st.dtq is not None else None
        io.dpooled = st.dpooled.data_ptr() if st.dpooled is not None else None
        io.dqk = st.dqk.data_ptr() if st.dqk is not None else None
        io.hazards = st.hazards.data_ptr()
        io.S = st.S.data_ptr()
        io.Y = st.Y.data_ptr()
        io.att_path = st.att_path.data_ptr()
        io.att_omic = st.att_omic.data_ptr()
        # train mode: the tail's dropout layers draw from the same (seed, device seed word) as the bag stage
        io.drop_p = self.tail_dropout if getattr(st, "train", False) else 0.0
        io.seed = getattr(st, "seed", 0) & 0xFFFFFFFF
        io.seed_dev = st.seed_dev.data_ptr() if (getattr(st, "train", False) and st.seed_dev is not None) else None
        io.train = 1 if getattr(st, "train", False) else 0
        return io

    # -- buffers
    def alloc_state(self, model, bag, save_for_backward=True, reuse_ws=False, with_backward_buffers=False):
        """All device buffers one pass over `bag` needs (caller-owned in the C ABI, allocated through torch)."""
        bnd = self.binding
        dev = bag.x.device
        B, K = bag.num_slides, bnd.n_classes
        f32 = dict(dtype=torch.float32, device=dev)
        st = SlideState()
        st.B, st.bag = B, bag
        st.tail_ws = self._tail_ws(model, B, dev, reuse_ws)
        st.qp = torch.empty((B, Q, D), **f32)
        st.qk = torch.empty((B, Q, D), **f32)
        nac = bnd.variant == VARIANT_NACAGAT
        st.kc = torch.empty((B, Q), **f32) if nac else None
        st.dpooled = st.dqk = st.dsuma = st.dkc = st.dtq = None
        st.hazards = torch.empty((B, K), **f32)
        st.S = torch.empty((B, K), **f32)
        st.Y = torch.empty((B, K), **f32)
        st.att_path = torch.empty((B, Q), **f32)
        st.att_omic = torch.empty((B, Q), **f32)
        st.bag_ws = bp.BagWorkspace(bag, save_h=save_for_backward or nac, nacagat=nac,
                                    save_gate=nac and save_for_backward)
        st.seed_dev = None
        if with_backward_buffers:
            st.bag_ws.ensure_bwd(bag)
            st.dpooled = torch.empty((B, Q, D), **f32)
            st.dqk = torch.empty((B, Q, D), **f32)
            if nac:
                st.dsuma = torch.empty((B, Q), **f32)
                st.dkc = torch.empty((B, Q), **f32)
                st.dtq = torch.empty((B, Q, D), **f32)
            st.loss = torch.empty(B, **f32)
            st.dhz = torch.empty((B, K), **f32)
            st.dS = torch.empty((B, K), **f32)
        return st

    # -- forward
    def forward(self, model, bag, omics, train=False, save_for_backward=True, seed=None, reuse_ws=False,
                after_bag=None, st=None, post=None):
        """bag: PackedBag of B slides; omics: 6 tensors [B, d_i] float32 on the GPU.  `st` reuses the buffers of an
        earlier alloc_state() (static addresses: needed for CUDA-graph capture)."""
        bnd = self.binding
        nac = bnd.variant == VARIANT_NACAGAT
        B = bag.num_slides
        if st is None:
            st = self.alloc_state(model, bag, save_for_backward, reuse_ws)
        elif st.B != B or st.bag is not bag:
            raise RuntimeError("the reused slide state was allocated for a different packed bag")
        st.train = bool(train)
        st.omics = []
        for i, o in enumerate(omics):
            require_cuda(o, "omics[%d]" % i)
            o = o.to(torch.float32).reshape(B, -1).contiguous()
            if o.shape[1] != bnd.omic_sizes[i]:
                raise RuntimeError("omics[%d] has width %d, the model expects %d" % (i, o.shape[1], bnd.omic_sizes[i]))
            st.omics.append(o)
        st.drop_p = self.bag_dropout if train else 0.0
        st.attn_p = self.attn_dropout if (train and nac) else 0.0
        st.seed = (_next_seed() if seed is None else seed) if train else 0
        P = bnd.params()
        # bf16 streaming copy of H.0.weight (refreshed every pass: the optimizer updates the fp32 master in place)
        w_h = P["H.0.weight"]
        if self._w_bf16 is None or self._w_bf16.device != w_h.device:
            self._w_bf16 = torch.empty(w_h.shape, dtype=torch.bfloat16, device=w_h.device)
        bp.cast_bf16(w_h.detach(), out=self._w_bf16)
        io = self._io(st)
        s = _stream()
        _lib.call("mpo_tail_pre_fwd", ctypes.byref(model), ctypes.byref(io), s)
        if nac:
            # fp16 streaming copy of the key projection (the gate kernels read it as a tensor-core operand)
            w_in = P["co_attention.in_proj_weight"].detach()
            if self._wk_f16 is None or self._wk_f16.device != w_in.device:
                self._wk_f16 = torch.empty((D, D), dtype=torch.float16, device=w_in.device)
            bp.cast_f16(w_in[D:2 * D], out=self._wk_f16)
            sd = st.seed_dev if train else None
            bp.bag_project(bag, self._w_bf16, P["H.0.bias"].detach(), st.qk, st.bag_ws, seed=st.seed, drop_p=st.drop_p,
                           seed_dev=sd)
            bp.bag_gate_forward(bag, self._wk_f16, P["co_attention.in_proj_bias"].detach()[D:2 * D], st.qp, st.kc,
                                st.bag_ws, seed=st.seed, attn_drop_p=st.attn_p, seed_dev=sd)
        else:
            bp.bag_forward(bag, self._w_bf16, P["H.0.bias"].detach(), st.qk, st.bag_ws, seed=st.seed, drop_p=st.drop_p,
                           seed_dev=st.seed_dev if train else None)
        if after_bag is not None:        # e.g. the cross-GPU log-sum-exp combine of a patch-sharded bag (dp.py)
            after_bag(st)
        if post is not None:             # training step: post forward + loss + post backward as one C-ABI call
            post(st, io, s)
        else:
            _lib.call("mpo_tail_post_fwd", ctypes.byref(model), ctypes.byref(io), s)
        return st

    def attention_map(self, st):
        """normalised co-attention map [6, total_rows] (attention_scores['coattn'])."""
        return bp.attention_map(st.bag, st.bag_ws, seed=st.seed, seed_dev=st.seed_dev if st.train else None,
                                attn_drop_p=getattr(st, "attn_p", 0.0))

    def bag_backward_only(self, model, st):
        """the bag stage of the backward pass (mpo_bag_bwd / mpo_bag_bwd_nacagat) given st.dpooled (and st.dsuma)."""
        dev = st.bag.x.device
        f32 = dict(dtype=torch.float32, device=dev)
        s = _stream()
        if not model.H.gw or not model.H.gb:
            raise RuntimeError("backward needs a model binding with gradient buffers")
        gw, gb = ctypes.c_void_p(model.H.gw), ctypes.c_void_p(model.H.gb)
        st.bag_ws.ensure_bwd(st.bag)
        if st.dqk is None:
            st.dqk = torch.empty((st.B, Q, D), **f32)
        ws = st.bag_ws
        if ws.nacagat:
            if ws.t_saved is None:
                raise RuntimeError("this NaCAGaT forward pass did not keep the gate activations (it ran under no_grad)")
            if st.dkc is None:
                st.dkc = torch.empty((st.B, Q), **f32)
                st.dtq = torch.empty((st.B, Q, D), **f32)
            use_suma = getattr(st, "attn_p", 0.0) > 0.0
            a = _lib.MpoNacagatBwd()
            for name, t in (("h_saved", ws.h_saved), ("t_saved", ws.t_saved), ("scores", ws.scores), ("pgate", ws.pgate),
                            ("lse", ws.lse), ("pooled", ws.pooled), ("suma", ws.suma if use_suma else None),
                            ("pooled_lo", ws.pooled_lo),
                            ("dpooled", st.dpooled), ("dsuma", st.dsuma if use_suma else None),
                            ("d_amap", getattr(st, "d_amap", None)), ("amap_dot", getattr(st, "amap_dot", None)),
                            ("qk", st.qk),
                            ("qp", st.qp), ("w_k_f16", self._wk_f16), ("dz_ws", ws.dz), ("dkg_ws", ws.dkg),
                            ("dg_ws", ws.dg), ("part_dqk", ws.part_dqk), ("part_dtq", ws.part_dtq),
                            ("part_db", ws.part_db), ("part_dbk", ws.part_dbk), ("part_dkc", ws.part_dkc),
                            ("dg_max", ws.dg_max), ("dqk", st.dqk), ("dkc", st.dkc), ("dtq", st.dtq)):
                setattr(a, name, t.data_ptr() if t is not None else None)
            a.grad_w_h, a.grad_b_h = model.H.gw, model.H.gb
            a.grad_w_k = model.coattn_in.gw + D * D * 4           # key block of co_attention.in_proj_weight.grad
            a.grad_b_k = model.coattn_in.gb + D * 4
            a.drop_p, a.attn_drop_p = st.drop_p, getattr(st, "attn_p", 0.0)
            a.seed = st.seed & 0xFFFFFFFF
            a.seed_dev = st.seed_dev.data_ptr() if (st.seed_dev is not None and st.train) else None
            _lib.call("mpo_bag_bwd_nacagat", st.bag.c(), ctypes.byref(a), s)
        else:
            _lib.call("mpo_bag_bwd", st.bag.c(), _ptr(ws.h_saved), _ptr(ws.scores), _ptr(ws.lse), _ptr(ws.pooled),
                      _ptr(st.dpooled), _ptr(st.qk), _ptr(ws.dz), _ptr(ws.part_dqk), _ptr(ws.part_db), _ptr(st.dqk),
                      gw, gb, _ptr(getattr(st, "d_amap", None)), _ptr(getattr(st, "amap_dot", None)),
                      ctypes.c_float(st.drop_p), s)

    # -- backward
    def backward(self, model, st, dhaz, dS, dY, post_done=False):
        """model must carry gradient pointers; they are accumulated into.  post_done: the post stage's backward
        already ran inside mpo_tail_post_step (st.dpooled is filled)."""
        dev = st.bag.x.device
        f32 = dict(dtype=torch.float32, device=dev)
        if st.dpooled is None:
            st.dpooled = torch.empty((st.B, Q, D), **f32)
        if st.bag_ws.nacagat and st.dsuma is None:
            st.dsuma = torch.empty((st.B, Q), **f32)
            st.dkc = torch.empty((st.B, Q), **f32)
            st.dtq = torch.empty((st.B, Q, D), **f32)
        io = self._io(st)
        s = _stream()

        def prep(g):
            if g is None:
                return None
            return g.detach().to(torch.float32).reshape(st.B, -1).contiguous()

        if not post_done:
            dhaz, dS, dY = prep(dhaz), prep(dS), prep(dY)
            _lib.call("mpo_tail_post_bwd", ctypes.byref(model), ctypes.byref(io), _ptr(dhaz), _ptr(dS), _ptr(dY), s)
        self.bag_backward_only(model, st)
        io = self._io(st)
        _lib.call("mpo_tail_pre_bwd", ctypes.byref(model), ctypes.byref(io), s)


# ------------------------------------------------------------------------------------------------ per-slide autograd
class _SlideFn(torch.autograd.Function):
    """One slide through the engine with gradients delivered to nn.Parameter.grad by autograd."""

    @staticmethod
    def forward(ctx, engine, want_map, train, needs_bwd, wsi, n_omics, *rest):
        omics = rest[:n_omics]
        params = rest[n_omics:]
        bnd = engine.binding
        bag = bp.PackedBag.from_slides([wsi.detach()])
        model = bnd.build(grads=None)
        st = engine.forward(model, bag, [o.detach().reshape(1, -1) for o in omics], train=train,
                            save_for_backward=needs_bwd)
        ctx.engine, ctx.state, ctx.n_omics, ctx.n_params = engine, st, n_omics, len(params)
        coattn = engine.attention_map(st) if want_map else torch.empty(0, device=wsi.device)
        st.amap = coattn if want_map else None
        # the outputs are COPIES of the state's buffers: returning st.hazards itself would close a reference cycle
        # (ctx -> state -> hazards -> grad_fn -> ctx) that only the cyclic GC breaks, and every slide's saved
        # activations would stay allocated (fresh cudaMallocs per call) until it runs
        a_path, a_omic = st.att_path.clone(), st.att_omic.clone()
        outs = (st.hazards.clone(), st.S.clone(), st.Y.clone(), coattn, a_path, a_omic)
        # the map is differentiable (CrossEntropySurvivalAttnRegLoss feeds attention_scores['coattn'] to the loss,
        # models/loss.py:88-101 / models/nacagat/main.py:49-50); the pooling logits are not consumed by any loss
        if want_map and needs_bwd:
            ctx.mark_non_differentiable(a_path, a_omic)
        else:
            ctx.mark_non_differentiable(coattn, a_path, a_omic)
        ctx.set_materialize_grads(False)
        return outs

    @staticmethod
    def backward(ctx, dhaz, dS, dY, dA=None, *unused):
        engine, st = ctx.engine, ctx.state
        ctx.state = None
        if dA is not None and st.amap is not None:
            # a gradient arrived on the returned map: dense [6, N] upstream gradient + its softmax-Jacobian dot
            st.d_amap = dA.detach().to(torch.float32).reshape(Q, -1).contiguous()
            st.amap_dot = bp.attention_map_dot(st.bag, st.amap.detach(), st.d_amap)
        bnd = engine.binding
        P = bnd.params()
        if st.bag_ws.h_saved is None:
            raise RuntimeError("this forward pass did not keep activations (it ran under no_grad)")
        none_in = (None, None, None, None, None, None) + (None,) * ctx.n_omics
        if _ACCUMULATE_IN_PLACE and all(p.requires_grad for p in P.values()):
            # The kernels accumulate (+=) like torch's AccumulateGrad, so they write straight into p.grad (created as
            # zeros when absent) and autograd gets no per-parameter gradient back: 100 tensor adds per slide less.
            grads = {}
            for n, p in P.items():
                if p.grad is None:
                    p.grad = torch.zeros_like(p)
                elif not (p.grad.is_contiguous() and p.grad.dtype == torch.float32 and p.grad.is_cuda):
                    raise RuntimeError("parameter %s has a .grad the kernels cannot accumulate into" % n)
                grads[n] = p.grad
            model = bnd.build(grads=grads)
            engine.backward(model, st, dhaz, dS, dY)
            return none_in + (None,) * ctx.n_params
        total = sum(p.numel() for p in P.values())
        flat = torch.zeros(total, dtype=torch.float32, device=st.bag.x.device)
        grads, off = {}, 0
        for n, p in P.items():
            grads[n] = flat[off:off + p.numel()].view_as(p)
            off += p.numel()
        model = bnd.build(grads=grads)
        engine.backward(model, st, dhaz, dS, dY)
        return none_in + tuple(grads[n] for n in bnd.names)


# MPO_AUTOGRAD_RETURN_GRADS=1: hand the parameter gradients back to autograd (hooks on parameters fire, one add per
# parameter and slide) instead of accumulating into .grad in place
_ACCUMULATE_IN_PLACE = os.environ.get("MPO_AUTOGRAD_RETURN_GRADS", "0") != "1"


def run_slide(engine, wsi, omics, want_map, train):
    """Module-level entry used by the drop-in forward()s.  Returns hazards, S, Y [1,K], coattn or None, path, omic."""
    require_cuda(wsi, "wsi")
    bnd = engine.binding
    P = bnd.params(refresh=True)
    params = [P[n] for n in bnd.names]
    needs_bwd = torch.is_grad_enabled() and any(p.requires_grad for p in params)
    outs = _SlideFn.apply(engine, bool(want_map), bool(train), needs_bwd, wsi, len(omics), *omics, *params)
    hazards, S, Y, coattn, a_path, a_omic = outs
    return hazards, S, Y, (coattn if want_map else None), a_path, a_omic


# ------------------------------------------------------------------------------------------------ batched training
def rebind_grad_views(params, grad_views):
    """optimizer.zero_grad() / module.zero_grad() default to set_to_none=True (it is the call in the reference loop,
    models/mcat/main.py:72-74) and drop the .grad views of the flat gradient buffer bound by BatchTrainer; the kernels
    would keep accumulating into the flat buffer while a torch optimizer sees grad None and skips every parameter.
    Re-bind: a parameter whose .grad was set to None gets its (zeroed) slice back, a foreign .grad tensor is copied
    into the slice."""
    for n, p in params.items():
        g = grad_views[n]
        if p.grad is g:
            continue
        if p.grad is None:
            g.zero_()
        elif p.grad.data_ptr() != g.data_ptr():
            g.copy_(p.grad)
        p.grad = g


class BatchTrainer:
    """Forward + loss + backward for B slides per call with gradients accumulated straight into one flat fp32
    buffer whose slices are the parameters' .grad (so a single NCCL all-reduce covers the whole model).

    Loop semantics follow the reference driver (models/mcat/main.py:30-74): loss / grad_acc_step, gradients
    accumulate until the caller steps the optimizer."""

    def __init__(self, module, loss="nll", alpha=None, eps=1e-7, grad_acc_step=32):
        self.module = module
        self.engine = module._engine
        bnd = self.engine.binding
        P = bnd.params()
        dev = next(iter(P.values())).device
        # every parameter (and its gradient) is a 256-byte-aligned slice of one flat fp32 buffer: one all-reduce and
        # one optimizer kernel cover the whole model
        self.offsets, off = {}, 0
        for n, p in P.items():
            self.offsets[n] = off
            off += (p.numel() + 63) // 64 * 64
        total = off
        self.flat_grad = torch.zeros(total, dtype=torch.float32, device=dev)
        self.flat_param = None
        self.grads = {}
        for n, p in P.items():
            g = self.flat_grad[self.offsets[n]:self.offsets[n] + p.numel()].view_as(p)
            p.grad = g
            self.grads[n] = g
        self.kind = {"nll": 0, "ces": 1}[loss]
        self.alpha = float(alpha if alpha is not None else (0.15 if loss == "nll" else 0.75))
        self.eps = float(eps)
        self.grad_acc_step = int(grad_acc_step)
        self.model = bnd.build(grads=self.grads)

    def zero_grad(self, set_to_none=False):
        """zeroes the flat gradient buffer; the parameters' .grad stay views of it (set_to_none is accepted for
        signature compatibility and ignored: the kernels accumulate into this buffer by address)."""
        self.flat_grad.zero_()
        self.check_grad_views()

    def check_grad_views(self):
        rebind_grad_views(self.engine.binding.params(), self.grads)

    # -- optimizer over the flat buffers (mpo_adam_step)
    def use_flat_adam(self, lr=2e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-5, peer=None):
        """Moves the parameters into one flat buffer (their nn.Parameter objects stay, .data becomes a view) and
        sets up Adam state; adam_step() then updates the whole model with one kernel (reference optimizer:
        torch.optim.Adam(lr=2e-4, weight_decay=1e-5), models/mcat/main.py:298-299, config.yaml:60-62).

        peer (a peer.PeerGroup): data-parallel training -- the flat gradient and parameter buffers move into
        IPC-exported memory that every rank maps, and peer_adam_step() replaces all-reduce + adam_step() by the sharded
        reduce-scatter / Adam / all-gather kernels of csrc/peer.cu."""
        P = self.engine.binding.params()
        dev = self.flat_grad.device
        self.peer = peer
        if peer is not None:
            from .peer import PeerBuffer
            n = self.flat_grad.numel()
            self._peer_grad, self._peer_param = PeerBuffer(4 * n, dev), PeerBuffer(4 * n, dev)
            new_grad = self._peer_grad.tensor(torch.float32)
            new_grad.copy_(self.flat_grad)
            self.flat_grad = new_grad
            for nme, p in P.items():
                g = self.flat_grad[self.offsets[nme]:self.offsets[nme] + p.numel()].view_as(p)
                p.grad = g
                self.grads[nme] = g
            self.flat_param = self._peer_param.tensor(torch.float32)
            peer.register_flat_buffers(self._peer_grad, self._peer_param)
        else:
            self.flat_param = torch.zeros_like(self.flat_grad)
        for n, p in P.items():
            view = self.flat_param[self.offsets[n]:self.offsets[n] + p.numel()].view_as(p)
            view.copy_(p.data)
            p.data = view
        self.adam_m = torch.zeros_like(self.flat_grad)
        self.adam_v = torch.zeros_like(self.flat_grad)
        self.adam_step_dev = torch.zeros(1, dtype=torch.int32, device=dev)
        self.adam_hp = (float(lr), float(betas[0]), float(betas[1]), float(eps), float(weight_decay))
        self.model = self.engine.binding.build(grads=self.grads)      # parameter addresses changed

    def adam_step(self, zero_grad=True, lo=0, hi=None, bump=True, side=False):
        """mpo_adam_step over elements [lo, hi) of the flat buffers (the whole model by default).  side=True
        (mpo_tail_side_adam): the bucket is one the post stage completes (see post_bucket_offset) -- the update is queued
        behind the post stage's side-stream weight gradients, next to the bag backward pass, and never bumps the step
        counter (the step's last adam_step does)."""
        lr, b1, b2, eps, wd = self.adam_hp
        hi = self.flat_grad.numel() if hi is None else hi
        flags = (1 if zero_grad else 0) | (0 if (bump and not side) else 2)       # MPO_ADAM_NO_BUMP = 2
        _lib.call("mpo_tail_side_adam" if side else "mpo_adam_step", _ptr(self.flat_param[lo:hi]),
                  _ptr(self.flat_grad[lo:hi]), _ptr(self.adam_m[lo:hi]), _ptr(self.adam_v[lo:hi]), hi - lo,
                  ctypes.c_float(lr), ctypes.c_float(b1), ctypes.c_float(b2), ctypes.c_float(eps), ctypes.c_float(wd),
                  _ptr(self.adam_step_dev), flags, _stream())

    def peer_adam_step(self, lo=0, hi=None, bump=True, slot=2):
        """mpo_peer_adam_step over elements [lo, hi) of the flat buffers (the whole model by default): needs
        use_flat_adam(peer=...).  The gradients must already carry the 1 / global-window scale (grad_acc_step)."""
        if getattr(self, "peer", None) is None:
            raise RuntimeError("peer_adam_step needs use_flat_adam(peer=PeerGroup)")
        lr, b1, b2, eps, wd = self.adam_hp
        hi = self.flat_grad.numel() if hi is None else hi
        _lib.call("mpo_peer_adam_step", self.peer.ref(), slot, lo, hi, _ptr(self.adam_m), _ptr(self.adam_v), 0,
                  ctypes.c_float(lr), ctypes.c_float(b1), ctypes.c_float(b2), ctypes.c_float(eps), ctypes.c_float(wd),
                  ctypes.c_float(1.0), _ptr(self.adam_step_dev), 1 if bump else 0, _stream())

    def _run_fwd(self, st, bag, omics, labels, censor, train, seed, inline_wgrad=False):
        """pre -> bag forward -> post forward + loss + post backward (mpo_tail_post_step).  With inline_wgrad the
        post stage's parameter gradients are complete when this part is (see post_bucket_offset)."""
        eng = self.engine
        if st is not None and st.seed_dev is not None and train:
            _lib.call("mpo_advance_seed", _ptr(st.seed_dev), _stream())
        if st is None:
            st = eng.alloc_state(self.model, bag, save_for_backward=True, reuse_ws=True, with_backward_buffers=True)
        flags = 1 if inline_wgrad else 0          # MPO_POST_STEP_INLINE_WGRAD

        def post(st_, io, s):
            # post forward + loss (models/loss.py) + post backward in one call: one cluster kernel on the fused path
            _lib.call("mpo_tail_post_step", ctypes.byref(self.model), ctypes.byref(io), self.kind, _ptr(labels),
                      _ptr(censor), ctypes.c_float(self.alpha), ctypes.c_float(self.eps),
                      ctypes.c_float(1.0 / self.grad_acc_step), _ptr(st_.loss), _ptr(st_.dhz), _ptr(st_.dS), flags, s)

        return eng.forward(self.model, bag, omics, train=train, save_for_backward=True, seed=seed, reuse_ws=True, st=st,
                           post=post)

    def _run_bwd(self, st):
        """bag backward -> pre backward (the rest of the step's gradients)."""
        self.engine.backward(self.model, st, st.dhz, st.dS, None, post_done=True)
        return st

    def _run(self, st, bag, omics, labels, censor, train, seed):
        return self._run_bwd(self._run_fwd(st, bag, omics, labels, censor, train, seed))

    def post_bucket_offset(self):
        """First element of the flat gradient buffer that belongs to the post-stage-only bucket: every parameter from
        co_attention.out_proj on (out projection, encoders, pooling heads, rho, fusion, classifier) is registered after
        H / G / co_attention.in_proj, and its gradient is final once mpo_tail_post_step has run -- a data-parallel
        trainer all-reduces flat_grad[offset:] while the bag backward pass is still running."""
        names = list(self.offsets.keys())
        first = names.index("co_attention.out_proj.weight")
        pre_only = ("H.", "G.", "co_attention.in_proj")
        if any(n.startswith(pre_only) for n in names[first:]):
            raise RuntimeError("unexpected parameter order: the post-stage gradient bucket is not a contiguous tail")
        return self.offsets[names[first]]

    def step(self, bag, omics, labels, censor, train=True, seed=None):
        """Returns (loss [B], hazards [B,K], S [B,K]).  labels int64 [B], censor float32 [B], on the GPU."""
        self.check_grad_views()
        st = self._run(None, bag, omics, labels, censor, train, seed)
        self.last_state = st
        return st.loss, st.hazards, st.S

    def evaluate(self, bag, omics, labels, censor, want_map=False):
        """Forward + loss for B slides, no backward (models/mcat/main.py:115-148 `validate`, :159-183 `test`):
        returns (loss [B], hazards [B,K], S [B,K], Y [B,K], map [6, rows] or None).  Eval mode (no dropout)."""
        eng = self.engine
        model = eng.binding.build(grads=None)
        st = eng.forward(model, bag, omics, train=False, save_for_backward=False)
        B, K = st.B, st.hazards.shape[1]
        loss = torch.empty(B, dtype=torch.float32, device=bag.x.device)
        scratch = torch.empty((2, B, K), dtype=torch.float32, device=bag.x.device)
        _lib.call("mpo_surv_loss", self.kind, _ptr(st.hazards), _ptr(st.S), _ptr(labels), _ptr(censor),
                  ctypes.c_float(self.alpha), ctypes.c_float(self.eps), ctypes.c_float(1.0), _ptr(loss), _ptr(scratch[0]),
                  _ptr(scratch[1]), B, K, _stream())
        amap = eng.attention_map(st) if want_map else None
        return loss, st.hazards, st.S, st.Y, amap

    def capture(self, bag, omics, labels, censor, train=True, with_adam=False, split=False, allreduce=False,
                peer_step=False, peer_overlap=True):
        """Record one step (forward + loss + backward) over these STATIC buffers into a CUDA graph.

        The caller refreshes the contents of bag.x / omics / labels / censor in place and calls replay(); the
        small tail launches then cost one graph launch.  Dropout masks change on every replay
        through a device-side seed.  split=True records the step as TWO graphs -- everything up to and including the
        post stage's gradients, then bag backward + pre backward -- so that a data-parallel caller can start the
        all-reduce of flat_grad[post_bucket_offset():] between them (replay_first() / replay_second()).
        allreduce=True (torch.distributed initialised, NCCL) records the whole data-parallel step as ONE graph instead:
        forward part, all-reduce of the post-stage bucket on NCCL's stream next to the bag backward part, all-reduce
        of the rest, then (with_adam) the optimizer step.
        peer_step=True (use_flat_adam(peer=...)): the whole data-parallel step INCLUDING its communication is one graph
        of plain kernels: forward part -> [side branch: mpo_peer_adam_step over the post-stage bucket] next to the bag
        backward part -> mpo_peer_adam_step over the rest.  Every rank must replay the same number of times."""
        eng = self.engine
        dev = bag.x.device
        st = eng.alloc_state(self.model, bag, save_for_backward=True, reuse_ws=False, with_backward_buffers=True)
        st.seed_dev = torch.tensor([_next_seed()], dtype=torch.int64, device=dev).to(torch.int32)
        omics = [o.to(torch.float32).reshape(bag.num_slides, -1).contiguous() for o in omics]
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):                         # warm-up outside the capture (lazy kernel attributes etc.)
                self._run_bwd(self._run_fwd(st, bag, omics, labels, censor, train, 0,
                                            inline_wgrad=split or allreduce or peer_step))
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        if allreduce:
            import torch.distributed as dist
            off = self.post_bucket_offset()
            for _ in range(2):                         # communicator / channel set-up outside the capture
                dist.all_reduce(self.flat_grad[off:])
                dist.all_reduce(self.flat_grad[:off])
            torch.cuda.synchronize()
        self.flat_grad.zero_()
        graph = torch.cuda.CUDAGraph()
        graph2 = torch.cuda.CUDAGraph() if split else None
        _lib.lib().mpo_launch_count(1)
        if peer_step:
            off = self.post_bucket_offset()
            branch = torch.cuda.Stream(device=dev)
            with torch.cuda.graph(graph):
                self._run_fwd(st, bag, omics, labels, censor, train, 0, inline_wgrad=True)
                if peer_overlap:
                    # the post-stage bucket (~2/3 of the model) is final here: its sharded optimizer step runs on a
                    # branch of the graph next to the bag backward pass, which only produces pre-stage gradients
                    branch.wait_stream(torch.cuda.current_stream())
                    with torch.cuda.stream(branch):
                        self.peer_adam_step(off, None, bump=False, slot=2)
                    self._run_bwd(st)
                    torch.cuda.current_stream().wait_stream(branch)
                    self.peer_adam_step(0, off, bump=True, slot=4)
                else:
                    self._run_bwd(st)
                    self.peer_adam_step(0, None, bump=True, slot=2)
        elif allreduce:
            with torch.cuda.graph(graph):
                self._run_fwd(st, bag, omics, labels, censor, train, 0, inline_wgrad=True)
                w1 = dist.all_reduce(self.flat_grad[off:], async_op=True)
                self._run_bwd(st)
                w2 = dist.all_reduce(self.flat_grad[:off], async_op=True)
                w1.wait()
                w2.wait()
                if with_adam:
                    self.adam_step(zero_grad=True)
        elif split:
            with torch.cuda.graph(graph):
                self._run_fwd(st, bag, omics, labels, censor, train, 0, inline_wgrad=True)
            with torch.cuda.graph(graph2, pool=graph.pool()):
                self._run_bwd(st)
        else:
            with torch.cuda.graph(graph):
                if with_adam:
                    # single GPU: the optimizer step (and the gradient reset) ride in the same graph; the post-stage
                    # bucket (~2/3 of the model) is updated behind its side-stream weight gradients, next to the bag
                    # backward pass, the rest after the pre-stage backward
                    off = self.post_bucket_offset()
                    self._run_fwd(st, bag, omics, labels, censor, train, 0)
                    self.adam_step(zero_grad=True, lo=off, side=True)
                    self._run_bwd(st)
                    self.adam_step(zero_grad=True, lo=0, hi=off)
                else:
                    self._run(st, bag, omics, labels, censor, train, 0)
        launches = int(_lib.lib().mpo_launch_count(1))     # kernels recorded into the graph(s) = launched per replay
        self.flat_grad.zero_()
        self.last_state = st
        return GraphedStep(graph, st, (bag, omics, labels, censor), launches, graph2, trainer=self)


class GraphedStep:
    """A captured train step over static buffers: refresh the buffers in place, then replay()."""

    def __init__(self, graph, state, static_inputs, launches=0, graph2=None, trainer=None):
        self.graph, self.graph2, self.state, self.static_inputs = graph, graph2, state, static_inputs
        self.trainer = trainer
        self.launches_per_replay = launches     # kernels of libmpo_b200.so inside the captured step
        self.replays = 0

    def replay_first(self):
        """split capture: pre, bag forward, post forward + loss + post backward (post-stage gradients complete)."""
        if self.trainer is not None and self.trainer.flat_param is None:
            self.trainer.check_grad_views()       # a torch optimizer reads p.grad: keep the views bound (see BatchTrainer)
        self.graph.replay()

    def replay_second(self):
        """split capture: bag backward + pre backward."""
        self.graph2.replay()
        self.replays += 1
        st = self.state
        return st.loss, st.hazards, st.S

    def replay(self):
        self.replay_first()
        if self.graph2 is not None:
            return self.replay_second()
        self.replays += 1
        st = self.state
        return st.loss, st.hazards, st.S
