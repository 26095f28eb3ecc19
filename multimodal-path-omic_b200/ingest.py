"""Bag ingestion: slide files -> pinned bf16 host staging -> HBM.

The reference reads one `<slide_id>.pt` tensor (or one HDF5 dataset) per slide inside `Dataset.__getitem__` and ships the
fp32 [N, 1024] tensor to the device with a blocking `.to(device)` (dataset/dataset.py:119-143, models/mcat/main.py:37).
With a ~35 us slide step the copy is the end-to-end bound (33.5 MB per 16 384-patch slide as bf16, twice that as fp32),
so this module
  * converts fp32 -> bf16 ONCE on the host, straight into pinned staging memory (half the PCIe bytes of the reference),
  * stages whole accumulation windows (B slides packed back to back: the layout the kernels stream) through TWO pinned
    buffers and two device buffers, H2D on a copy stream overlapped with the previous window's step, and
  * can keep the converted bags resident (pinned host cache, or in HBM: 512 slides x 16 384 patches are 17 GB of the
    180 GB) so that later epochs do no file I/O and -- for the HBM store -- no PCIe traffic at all.
PyTorch owns the memory and the streams; the step itself runs through libmpo_b200.so.
"""
import os

import numpy as np
import torch

from . import bagpass as bp

D_IN = bp.D_IN


# ------------------------------------------------------------------------------------------------ slide files
class SlideFileSource:
    """Random access to the patch-embedding bags of a cohort, as the reference stores them.

    kind 'pt'  : `<dir>/<slide_id>.pt`  torch.save'd [N, 1024] tensors            (dataset/dataset.py:126)
    kind 'npy' : `<dir>/<slide_id>.npy`
    kind 'h5'  : one HDF5 file, dataset `<slide_id>` -> [N, 1024]                   (dataset/dataset.py:129)
                 needs h5py, which this image does not ship: a clear ImportError is raised, nothing is emulated."""

    def __init__(self, path, kind=None):
        self.path = path
        if kind is None:
            kind = "h5" if (os.path.isfile(path) and path.endswith((".h5", ".hdf5"))) else None
            if kind is None:
                names = os.listdir(path)
                kind = "pt" if any(n.endswith(".pt") for n in names) else "npy"
        self.kind = kind
        self._h5 = None
        if kind == "h5":
            try:
                import h5py
            except ImportError as exc:
                raise ImportError("an HDF5 bag file needs h5py, which is not installed in this environment; "
                                  "convert the bags to .pt / .npy files or install h5py") from exc
            self._h5 = h5py.File(path, "r")

    @staticmethod
    def slide_key(slide_id):
        return slide_id[:-4] if slide_id.endswith(".svs") else slide_id      # dataset.py:125,128

    def has(self, slide_id):
        key = self.slide_key(slide_id)
        if self.kind == "h5":
            return key in self._h5
        return os.path.exists(os.path.join(self.path, key + "." + self.kind))

    def load(self, slide_id):
        """-> CPU tensor [N, 1024], fp32 or bf16, contiguous."""
        key = self.slide_key(slide_id)
        if self.kind == "pt":
            t = torch.load(os.path.join(self.path, key + ".pt"), map_location="cpu")
        elif self.kind == "npy":
            t = torch.from_numpy(np.load(os.path.join(self.path, key + ".npy")))
        else:
            t = torch.from_numpy(np.asarray(self._h5[key]))
        if t.dim() == 3 and t.shape[0] == 1:
            t = t[0]
        if t.dim() != 2 or t.shape[1] != D_IN:
            raise RuntimeError("bag of slide %s has shape %s, expected [N, %d]" % (slide_id, tuple(t.shape), D_IN))
        if t.dtype not in (torch.float32, torch.bfloat16):
            t = t.to(torch.float32)
        return t.contiguous()


def to_bf16_into(dst, src):
    """fp32/bf16 CPU bag -> rows of a (pinned) bf16 host buffer, converting on the host cores (round to nearest even)."""
    dst.copy_(src)          # torch's CPU cast is vectorised and multi-threaded; dst decides the dtype
    return dst


# ------------------------------------------------------------------------------------------------ NUMA placement
def bind_to_gpu_numa_node(device_index):
    """Pin this process to the CPU cores next to its GPU BEFORE pinned staging memory is allocated, so first-touch puts
    the staging buffers on the GPU's own NUMA node (one process per GPU: without this all eight ranks of a node share the
    cores and the memory controller of node 0 -- the round-1 8-GPU end-to-end line reached 3.3x of one GPU).
    Returns the CPU list, or None when the topology is not exposed (containers often hide /sys)."""
    try:
        props = torch.cuda.get_device_properties(device_index)
        bdf = "%04x:%02x:%02x.0" % (props.pci_domain_id, props.pci_bus_id, props.pci_device_id)
        with open("/sys/bus/pci/devices/%s/local_cpulist" % bdf) as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        allowed = os.sched_getaffinity(0)
        cpus = sorted(cpus & allowed) if cpus & allowed else None
        if cpus:
            os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:
        return None


# ------------------------------------------------------------------------------------------------ window staging
class WindowStager:
    """Double-buffered host -> device staging of accumulation windows.

    A window is a list of slides (bag [N_b, 1024] CPU tensors + omics + label + censorship).  `stage(k, window)` packs
    the bags as bf16 into pinned slot k % 2 and enqueues the H2D copies on the copy stream; `acquire(k)` makes the compute
    stream wait for them and returns (PackedBag, omics, labels, censor) over device slot k % 2; `release(k)` lets the
    copy stream overwrite the slot once the step that read it has been enqueued."""

    def __init__(self, device, max_rows, max_slides, omic_sizes):
        self.device = torch.device(device)
        self.max_rows, self.max_slides = int(max_rows), int(max_slides)
        self.omic_sizes = [int(d) for d in omic_sizes]
        self.host_x = [torch.empty((self.max_rows, D_IN), dtype=torch.bfloat16).pin_memory() for _ in range(2)]
        self.host_om = [[torch.empty((self.max_slides, d), dtype=torch.float32).pin_memory() for d in self.omic_sizes]
                        for _ in range(2)]
        self.host_lab = [torch.empty(self.max_slides, dtype=torch.int64).pin_memory() for _ in range(2)]
        self.host_cen = [torch.empty(self.max_slides, dtype=torch.float32).pin_memory() for _ in range(2)]
        self.dev_x = [torch.empty((self.max_rows, D_IN), dtype=torch.bfloat16, device=self.device) for _ in range(2)]
        self.dev_om = [[torch.empty((self.max_slides, d), dtype=torch.float32, device=self.device)
                        for d in self.omic_sizes] for _ in range(2)]
        self.dev_lab = [torch.empty(self.max_slides, dtype=torch.int64, device=self.device) for _ in range(2)]
        self.dev_cen = [torch.empty(self.max_slides, dtype=torch.float32, device=self.device) for _ in range(2)]
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self.ready = [torch.cuda.Event() for _ in range(2)]
        self.freed = [torch.cuda.Event() for _ in range(2)]
        self.host_done = [torch.cuda.Event() for _ in range(2)]     # the H2D copy has read the pinned slot
        for e in self.freed + self.host_done:
            e.record(torch.cuda.current_stream(self.device))
        self.meta = [None, None]
        self.h2d_bytes = 0

    def stage(self, k, window):
        slot = k % 2
        lengths = [int(s["bag"].shape[-2]) for s in window]
        rows, B = sum(lengths), len(window)
        if rows > self.max_rows or B > self.max_slides:
            raise RuntimeError("window of %d slides / %d rows exceeds the stager's %d / %d" %
                               (B, rows, self.max_slides, self.max_rows))
        self.host_done[slot].synchronize()           # the previous copy out of this pinned slot has finished
        r = 0
        for b, s in enumerate(window):
            bag = s["bag"][0] if s["bag"].dim() == 3 else s["bag"]
            to_bf16_into(self.host_x[slot][r:r + lengths[b]], bag)
            r += lengths[b]
            for i, o in enumerate(s["omics"]):
                self.host_om[slot][i][b].copy_(o.reshape(-1))
            self.host_lab[slot][b] = int(s["label"])
            self.host_cen[slot][b] = float(s["censor"])
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.freed[slot])
            self.dev_x[slot][:rows].copy_(self.host_x[slot][:rows], non_blocking=True)
            for i in range(len(self.omic_sizes)):
                self.dev_om[slot][i][:B].copy_(self.host_om[slot][i][:B], non_blocking=True)
            self.dev_lab[slot][:B].copy_(self.host_lab[slot][:B], non_blocking=True)
            self.dev_cen[slot][:B].copy_(self.host_cen[slot][:B], non_blocking=True)
            self.host_done[slot].record(self.copy_stream)
            self.ready[slot].record(self.copy_stream)
        self.meta[slot] = (tuple(lengths), rows, B)
        self.h2d_bytes += rows * D_IN * 2 + B * (sum(self.omic_sizes) * 4 + 12)

    def acquire(self, k):
        slot = k % 2
        lengths, rows, B = self.meta[slot]
        torch.cuda.current_stream(self.device).wait_event(self.ready[slot])
        bag = bp.PackedBag(self.dev_x[slot][:rows], lengths)
        return bag, [o[:B] for o in self.dev_om[slot]], self.dev_lab[slot][:B], self.dev_cen[slot][:B]

    def release(self, k):
        self.freed[k % 2].record(torch.cuda.current_stream(self.device))


def iterate_windows(samples, window):
    """groups an iterable of per-slide samples into lists of `window` slides (the last one may be shorter)."""
    cur = []
    for s in samples:
        cur.append(s)
        if len(cur) == window:
            yield cur
            cur = []
    if cur:
        yield cur


class ResidentBagStore:
    """All bags of a cohort as ONE bf16 [rows, 1024] buffer in HBM, loaded and converted once.

    `window(ids)` returns a PackedBag over the slides `ids`: a zero-copy view when they are stored back to back in that
    order (sequential epochs, validation), otherwise a device-side gather into a scratch buffer (shuffled epochs: one D2D
    copy per window instead of file I/O + PCIe)."""

    def __init__(self, device, bags):
        self.device = torch.device(device)
        self.lengths = [int(b.shape[-2]) for b in bags]
        self.offsets = np.concatenate([[0], np.cumsum(self.lengths)]).astype(np.int64)
        self.x = torch.empty((int(self.offsets[-1]), D_IN), dtype=torch.bfloat16, device=self.device)
        stage = None
        for i, b in enumerate(bags):
            b = b[0] if b.dim() == 3 else b
            if stage is None or stage.shape[0] < b.shape[0]:
                stage = torch.empty((b.shape[0], D_IN), dtype=torch.bfloat16).pin_memory()
            to_bf16_into(stage[:b.shape[0]], b)
            self.x[int(self.offsets[i]):int(self.offsets[i + 1])].copy_(stage[:b.shape[0]])      # blocking: stage is reused
        self._scratch = None

    def window(self, ids):
        ids = [int(i) for i in ids]
        lengths = [self.lengths[i] for i in ids]
        if all(ids[j] + 1 == ids[j + 1] for j in range(len(ids) - 1)):
            a, b = int(self.offsets[ids[0]]), int(self.offsets[ids[-1] + 1])
            return bp.PackedBag(self.x[a:b], lengths)
        rows = sum(lengths)
        if self._scratch is None or self._scratch.shape[0] < rows:
            self._scratch = torch.empty((rows, D_IN), dtype=torch.bfloat16, device=self.device)
        r = 0
        for i, n in zip(ids, lengths):
            self._scratch[r:r + n].copy_(self.x[int(self.offsets[i]):int(self.offsets[i + 1])])
            r += n
        return bp.PackedBag(self._scratch[:rows], lengths)
