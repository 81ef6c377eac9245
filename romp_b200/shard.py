"""Frame-sharded multi-GPU execution.

Frames are independent, so rank r simply runs the whole hot path on its contiguous slice of the batch; no
collective is needed to compute.  The single collective is the all-gather that *collects* the per-person
outputs: one all_gather of the person counts, then one padded all_gather of a packed float record and one of
a packed int64 record.  ``pred_batch_ids`` are offset by the rank's first frame (the reference does the same
bookkeeping for nn.DataParallel: romp/lib/maps_utils/result_parser.py:59-64).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

FLOAT_KEYS = ["cam", "smpl_thetas", "smpl_betas", "center_confs", "cam_trans", "joints", "pj2d_org", "verts"]
INT_KEYS = ["center_preds", "pred_batch_ids"]


def shard_range(total_frames: int, rank: int, world: int):
    """Contiguous frame range of `rank` (remainder frames go to the first ranks)."""
    base, rem = divmod(total_frames, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def pack(out, frame_offset, device):
    """dict (or None) -> (float record [n, F], int record [n, 3], layout)."""
    if out is None:
        return torch.zeros(0, 0, device=device), torch.zeros(0, 3, dtype=torch.int64, device=device), None
    t = {k: (torch.from_numpy(v) if isinstance(v, np.ndarray) else v).to(device) for k, v in out.items()}
    n = t["cam"].shape[0]
    fkeys = [k for k in FLOAT_KEYS if k in t]
    layout = [(k, tuple(t[k].shape[1:])) for k in fkeys]
    frec = torch.cat([t[k].reshape(n, -1).float() for k in fkeys], 1)
    irec = torch.cat([t["center_preds"].reshape(n, 2), (t["pred_batch_ids"] + frame_offset).reshape(n, 1)], 1)
    return frec.contiguous(), irec.contiguous(), layout


def unpack(frec, irec, layout):
    out, col = {}, 0
    n = frec.shape[0]
    for k, shp in layout:
        w = int(np.prod(shp)) if len(shp) else 1
        out[k] = frec[:, col:col + w].reshape((n,) + shp)
        col += w
    out["center_preds"] = irec[:, :2]
    out["pred_batch_ids"] = irec[:, 2]
    out["global_orient"] = out["smpl_thetas"][:, :3]
    out["body_pose"] = out["smpl_thetas"][:, 3:]
    return out


def all_gather_outputs(out, frame_offset: int, world: int, to_numpy: bool = True, group=None):
    """All ranks end up with the outputs of every rank, ordered by (global frame asc, score desc)."""
    backend = dist.get_backend(group)
    device = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    frec, irec, layout = pack(out, frame_offset, device)
    n = torch.tensor([frec.shape[0], frec.shape[1]], dtype=torch.int64, device=device)
    counts = torch.zeros(world * 2, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(counts, n, group=group)
    counts = counts.cpu().view(world, 2)
    nmax, width = int(counts[:, 0].max()), int(counts[:, 1].max())
    if nmax == 0:
        return None
    if layout is None and width == DEFAULT_WIDTH:
        layout = DEFAULT_LAYOUT                       # a rank without persons still knows the standard record
    widths = set(int(w) for w in counts[:, 1].tolist() if int(w) > 0)
    if layout is None or len(widths) > 1:             # unusual key set: agree on it explicitly (pickled, slow path)
        layouts = [None] * world
        dist.all_gather_object(layouts, layout, group=group)
        layout = next(l for l in layouts if l is not None)
    fpad = torch.zeros(nmax, width, device=device)
    ipad = torch.zeros(nmax, 3, dtype=torch.int64, device=device)
    if frec.shape[0]:
        fpad[:frec.shape[0]] = frec
        ipad[:irec.shape[0]] = irec
    fall = torch.empty(world * nmax, width, device=device)
    iall = torch.empty(world * nmax, 3, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(fall, fpad, group=group)
    dist.all_gather_into_tensor(iall, ipad, group=group)
    ntot = int(counts[:, 0].sum())
    if not to_numpy:
        keep = torch.cat([torch.arange(r * nmax, r * nmax + int(counts[r, 0])) for r in range(world)]).to(device)
        return unpack(fall[keep], iall[keep], layout)
    # D2H of the valid rows of every rank's slab into cached pinned mirrors, then host-side views
    fh, ih = _pinned("f", ntot, width, torch.float32), _pinned("i", ntot, 3, torch.int64)
    row = 0
    for r in range(world):
        c = int(counts[r, 0])
        if c:
            fh[row:row + c].copy_(fall[r * nmax:r * nmax + c], non_blocking=True)
            ih[row:row + c].copy_(iall[r * nmax:r * nmax + c], non_blocking=True)
            row += c
    if device.type == "cuda":
        torch.cuda.current_stream().synchronize()
    return {k: v.numpy() for k, v in unpack(fh[:ntot], ih[:ntot], layout).items()}


class GatherPipeline:
    """Pipelined all_gather_outputs for a stream of steps (the multi-GPU `forward_batches` loop).

    submit() enqueues pack + the padded NCCL all-gathers + the device->host copies of the gathered rows on a side stream
    and returns at once (only the tiny person-count exchange is synchronous); result() of the PREVIOUS step is collected
    while the current step computes.  Host mirrors are double-buffered, so a result stays valid until the submit after
    the next.  `host_rank`: rank that wants numpy results (None = every rank); other ranks only take part in the
    collective and get None from result().  On a non-CUDA backend (gloo tests) everything runs synchronously.
    """

    def __init__(self, world: int, group=None, host_rank=None):
        self.world, self.group, self.host_rank = world, group, host_rank
        self.cuda = dist.get_backend(group) == "nccl"
        self.stream = torch.cuda.Stream() if self.cuda else None
        self.slot = 0

    def submit(self, out, frame_offset: int):
        if not self.cuda:
            return ("done", all_gather_outputs(out, frame_offset, self.world, True, self.group))
        rank = dist.get_rank(self.group)
        want_host = self.host_rank is None or rank == self.host_rank
        device = torch.device("cuda", torch.cuda.current_device())
        frec, irec, layout = pack(out, frame_offset, device)
        n = torch.tensor([frec.shape[0], frec.shape[1]], dtype=torch.int64, device=device)
        counts = torch.zeros(self.world * 2, dtype=torch.int64, device=device)
        dist.all_gather_into_tensor(counts, n, group=self.group)
        counts = counts.cpu().view(self.world, 2)                  # the one synchronisation point (16 B per rank)
        nmax, width = int(counts[:, 0].max()), int(counts[:, 1].max())
        if nmax == 0:
            return ("done", None)
        if layout is None or width != DEFAULT_WIDTH:
            # unusual key set: fall back to the synchronous path (agrees on the layout explicitly)
            return ("done", all_gather_outputs(out, frame_offset, self.world, True, self.group))
        ready = torch.cuda.Event()
        ready.record()
        slot, self.slot = self.slot, self.slot ^ 1
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(ready)
            fpad = torch.zeros(nmax, width, device=device)
            ipad = torch.zeros(nmax, 3, dtype=torch.int64, device=device)
            if frec.shape[0]:
                fpad[:frec.shape[0]] = frec
                ipad[:irec.shape[0]] = irec
            fall = torch.empty(self.world * nmax, width, device=device)
            iall = torch.empty(self.world * nmax, 3, dtype=torch.int64, device=device)
            dist.all_gather_into_tensor(fall, fpad, group=self.group)
            dist.all_gather_into_tensor(iall, ipad, group=self.group)
            ntot = int(counts[:, 0].sum())
            fh = ih = None
            if want_host:
                fh, ih = _pinned(("f", slot), ntot, width, torch.float32), _pinned(("i", slot), ntot, 3, torch.int64)
                row = 0
                for r in range(self.world):
                    c = int(counts[r, 0])
                    if c:
                        fh[row:row + c].copy_(fall[r * nmax:r * nmax + c], non_blocking=True)
                        ih[row:row + c].copy_(iall[r * nmax:r * nmax + c], non_blocking=True)
                        row += c
            done = torch.cuda.Event()
            done.record(self.stream)
        for t in (frec, irec, fpad, ipad, fall, iall):
            t.record_stream(self.stream)
        return ("pending", done, fh, ih, ntot, (fall, iall))

    def result(self, handle):
        if handle[0] == "done":
            return handle[1]
        _, done, fh, ih, ntot, _keep = handle
        done.synchronize()
        if fh is None:
            return None
        return {k: v.numpy() for k, v in unpack(fh[:ntot], ih[:ntot], DEFAULT_LAYOUT).items()}


DEFAULT_LAYOUT = [("cam", (3,)), ("smpl_thetas", (72,)), ("smpl_betas", (10,)), ("center_confs", (1,)), ("cam_trans", (3,)),
                  ("joints", (71, 3)), ("pj2d_org", (71, 2)), ("verts", (6890, 3))]
DEFAULT_WIDTH = sum(int(np.prod(s)) for _, s in DEFAULT_LAYOUT)
_PIN = {}


def _pinned(tag, rows, width, dtype):
    """Grow-only pinned host mirror (valid until the next gather that needs it)."""
    key = (tag, width, dtype)
    buf = _PIN.get(key)
    if buf is None or buf.shape[0] < rows:
        buf = torch.empty(max(rows, 64) * 2, width, dtype=dtype)
        if torch.cuda.is_available():
            buf = buf.pin_memory()
        _PIN[key] = buf
    return buf
