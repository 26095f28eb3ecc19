"""Batched training / validation / test drivers with the loop semantics of the reference.

Reference: models/mcat/main.py:19-183 (identical in models/nacagat/main.py): one slide per iteration, `loss /
grad_acc_step` back-propagated per slide, `optimizer.step(); optimizer.zero_grad()` every grad_acc_step slides (a short
last window keeps its gradient: the next epoch's first step includes it), risk = -sum(survs) per slide, epoch loss =
mean of the per-slide losses (+ lambda * l1_reg), c-index over the epoch (`concordance_index_censored` of
scikit-survival, restated in `concordance_index` below because sksurv is not a dependency of the hot path).

Here a whole accumulation window is ONE packed step (BatchTrainer): the slides of a window see the same weights, exactly
as in the reference, so the result is the same gradient, computed B slides at a time.
"""
import datetime
import os

import numpy as np
import torch

from . import bagpass as bp
from . import slidepath


def concordance_index(event, time, risk, tied_tol=1e-8):
    """Harrell's C as sksurv.metrics.concordance_index_censored computes it (models/mcat/main.py:81):
    pairs (i, j) are comparable when i had the event and time_i < time_j, or time_i == time_j and j is censored;
    concordant when risk_i > risk_j, tied when |risk_i - risk_j| <= tied_tol (counted 1/2)."""
    event = np.asarray(event, bool)
    time = np.asarray(time, np.float64)
    risk = np.asarray(risk, np.float64)
    conc = tied = comparable = 0
    for i in np.nonzero(event)[0]:
        mask = (time > time[i]) | ((time == time[i]) & ~event)
        mask[i] = False
        d = risk[i] - risk[mask]
        comparable += int(mask.sum())
        tied += int((np.abs(d) <= tied_tol).sum())
        conc += int((d > tied_tol).sum())
    if comparable == 0:
        raise ValueError("no comparable pairs: the c-index is undefined")
    return (conc + 0.5 * tied) / comparable


def _as_sample(item):
    """dataset item (survival_months, survival_class, censorship, omics, patches) (dataset/dataset.py:143) or a dict."""
    if isinstance(item, dict):
        return item
    months, label, censor, omics, bag = item
    return dict(months=float(months), label=int(label), censor=float(censor), omics=list(omics), bag=bag)


class EpochRunner:
    """train_epoch / validate / test over an iterable of slides (anything that yields the reference dataset's items)."""

    def __init__(self, module, optimizer=None, loss="ces", grad_acc_step=32, lambda_l1=0.0, alpha=None,
                 lr=2e-4, weight_decay=1e-5, group=None):
        self.module = module
        self.trainer = slidepath.BatchTrainer(module, loss=loss, alpha=alpha, grad_acc_step=grad_acc_step)
        self.grad_acc_step = int(grad_acc_step)
        self.lambda_l1 = float(lambda_l1)
        self.group = group
        self.optimizer = optimizer
        if optimizer is None:          # the reference's default optimizer (mcat/main.py:298-299) as one fused kernel
            self.trainer.use_flat_adam(lr=lr, weight_decay=weight_decay)
        self.pending = 0               # slides whose gradient sits in the buffer since the last optimizer step
        self.device = self.trainer.flat_grad.device

    # -- helpers
    def _to_window(self, samples):
        dev = self.device
        bags = [s["bag"].to(dev, non_blocking=True) for s in samples]
        pb = bp.PackedBag.from_slides(bags)
        nq = len(samples[0]["omics"])
        om = [torch.stack([s["omics"][i].reshape(-1) for s in samples]).to(dev, dtype=torch.float32) for i in range(nq)]
        labels = torch.tensor([s["label"] for s in samples], dtype=torch.int64, device=dev)
        cens = torch.tensor([s["censor"] for s in samples], dtype=torch.float32, device=dev)
        return pb, om, labels, cens

    def _l1(self):
        if not self.lambda_l1:
            return 0.0
        from .utils import l1_reg
        return float(l1_reg(self.module).item()) * self.lambda_l1

    def _l1_backward(self, n_slides):
        """the reference adds `loss_reg` to every slide's loss before backward (main.py:58-70): n_slides times
        lambda * sign(W) accumulate over a window."""
        if not self.lambda_l1:
            return
        import ctypes
        from . import _lib
        tr = self.trainer
        if tr.flat_param is not None:
            _lib.call("mpo_l1_grad", bp._ptr(tr.flat_param), bp._ptr(tr.flat_grad), tr.flat_grad.numel(),
                      ctypes.c_float(self.lambda_l1 * n_slides), bp._stream())
        else:
            for n, p in tr.engine.binding.params().items():
                _lib.call("mpo_l1_grad", bp._ptr(p.detach()), bp._ptr(tr.grads[n]), p.numel(),
                          ctypes.c_float(self.lambda_l1 * n_slides), bp._stream())

    def _optimizer_step(self):
        tr = self.trainer
        if self.group is not None:
            from . import dp
            dp.all_reduce_gradients(tr.flat_grad, self.group)
        if self.optimizer is None:
            tr.adam_step(zero_grad=True)
        else:
            tr.check_grad_views()
            self.optimizer.step()
            tr.zero_grad()
        self.pending = 0

    # -- epochs
    def train_epoch(self, slides, train_mode=True):
        """-> dict(loss, c_index, risk, censorship, event_time, optimizer_steps).  models/mcat/main.py:19-86."""
        self.module.train(train_mode)
        risks, cens, times, losses = [], [], [], []
        steps = 0
        window = []

        def flush():
            nonlocal steps
            if not window:
                return
            room = self.grad_acc_step - self.pending
            part, rest = window[:room], window[room:]
            pb, om, labels, c = self._to_window(part)
            loss, hz, S = self.trainer.step(pb, om, labels, c, train=train_mode)
            l1 = self._l1()
            self._l1_backward(len(part))
            losses.extend((loss.detach().cpu().numpy() + l1).tolist())
            risks.extend((-S.sum(dim=1)).detach().cpu().numpy().tolist())
            self.pending += len(part)
            if self.pending == self.grad_acc_step:
                self._optimizer_step()
                steps += 1
            window[:] = rest

        for item in slides:
            s = _as_sample(item)
            window.append(s)
            cens.append(s["censor"])
            times.append(s.get("months", 0.0))
            if self.pending + len(window) >= self.grad_acc_step:
                flush()
        while window:
            flush()
        out = dict(loss=float(np.mean(losses)) if losses else float("nan"), risk=np.asarray(risks),
                   censorship=np.asarray(cens), event_time=np.asarray(times), optimizer_steps=steps)
        try:
            out["c_index"] = concordance_index((1 - out["censorship"]).astype(bool), out["event_time"], out["risk"])
        except ValueError:
            out["c_index"] = float("nan")
        return out

    def validate(self, slides, window=32):
        """-> dict(loss, c_index, risk, ...).  models/mcat/main.py:115-155 (eval mode, no gradients)."""
        from .ingest import iterate_windows
        self.module.eval()
        risks, cens, times, losses = [], [], [], []
        l1 = self._l1()
        with torch.no_grad():
            for samples in iterate_windows((_as_sample(i) for i in slides), window):
                pb, om, labels, c = self._to_window(samples)
                loss, hz, S, Y, _ = self.trainer.evaluate(pb, om, labels, c)
                losses.extend((loss.cpu().numpy() + l1).tolist())
                risks.extend((-S.sum(dim=1)).cpu().numpy().tolist())
                cens.extend(s["censor"] for s in samples)
                times.extend(s.get("months", 0.0) for s in samples)
        out = dict(loss=float(np.mean(losses)) if losses else float("nan"), risk=np.asarray(risks),
                   censorship=np.asarray(cens), event_time=np.asarray(times))
        try:
            out["c_index"] = concordance_index((1 - out["censorship"]).astype(bool), out["event_time"], out["risk"])
        except ValueError:
            out["c_index"] = float("nan")
        return out

    def test(self, slides, output_dir=None, model_name="MCAT", patient="", epoch=0, window=8):
        """Inference with the co-attention maps (models/mcat/main.py:159-183): returns a list of dicts (hazards, survs,
        risk, Y, coattn [6, N_b]) and, with output_dir, writes `ATTN_<model>_<patient>_<now>_E<epoch>_<index>.pt` files
        holding the [6, N] map exactly as `torch.save(attention_scores['coattn'], ...)` does there."""
        from .ingest import iterate_windows
        self.module.eval()
        now = datetime.datetime.now().strftime('%Y%m%d%H%M%S')
        out, index = [], 0
        with torch.no_grad():
            for samples in iterate_windows((_as_sample(i) for i in slides), window):
                pb, om, labels, c = self._to_window(samples)
                _, hz, S, Y, amap = self.trainer.evaluate(pb, om, labels, c, want_map=True)
                for b in range(len(samples)):
                    r0, r1 = pb.slide_rows(b)
                    rec = dict(hazards=hz[b:b + 1].clone(), survs=S[b:b + 1].clone(), Y=Y[b:b + 1].clone(),
                               risk=float(-S[b].sum().item()), coattn=amap[:, r0:r1].clone())
                    if output_dir is not None:
                        save_attention_map(rec["coattn"], output_dir, model_name, patient, now, epoch, index)
                    out.append(rec)
                    index += 1
        return out


def save_attention_map(coattn, output_dir, model_name, patient, now, epoch, index):
    """models/mcat/main.py:180-183."""
    os.makedirs(output_dir, exist_ok=True)
    path = os.path.join(output_dir, f'ATTN_{model_name}_{patient}_{now}_E{epoch}_{index}.pt')
    torch.save(coattn.detach().cpu(), path)
    return path
