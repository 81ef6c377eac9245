"""Frame-sharded multi-GPU execution.

Frames are independent, so rank r simply runs the whole hot path on its contiguous slice of the batch; no
collective is needed to compute.  The single collective is the all-gather that *collects* the per-person outputs.

Design (round 2):
* The per-person record has a FIXED layout that follows from the model configuration (ROMP / BEV, calc_smpl, number of
  betas) - never from the data - so every rank knows it even when its shard detected nobody.  The sequence of
  collectives therefore never depends on a rank-local condition.
* Every rank packs ``[header row | rows]`` into a persistent send buffer with one kernel (``b200romp_pack_rows``; the
  person count is read on the device) and ONE ``all_gather_into_tensor`` ships ``1 + rows_hint`` rows per rank.  The
  header carries the rank's count and first frame, so no separate count exchange and no host synchronisation sit in
  front of the collective.  ``rows_hint`` is a host-side upper bound (previous step's maximum with slack); if a rank
  overflows it, every rank sees that in the gathered headers and all of them repeat the gather with the exact maximum
  (rare slow path, collectively decided).
* ``pred_batch_ids`` are offset by the rank's first frame on unpack (the reference does the same bookkeeping for
  nn.DataParallel: romp/lib/maps_utils/result_parser.py:59-64).
* Each rank reads back ITS OWN shard to its host (``ROMP.forward_batches``); the gathered records stay on the device
  unless a caller asks for them (``result(to_numpy=True)``).

On a non-CUDA backend (the world-size-2 ``gloo`` tests of the host logic) the same code runs on CPU tensors with a
torch implementation of the pack step.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

MAGIC = 0x0B200B20
HEADER_WORDS = 8


class RecordLayout:
    """Ordered fields (name, per-person shape, torch dtype) -> byte offsets inside one fixed-width record."""

    def __init__(self, fields):
        self.fields = [(n, tuple(s), d) for n, s, d in fields]
        self.offsets, off = [], 0
        for _, shp, dt in self.fields:
            nbytes = int(np.prod(shp, dtype=np.int64)) * torch.empty(0, dtype=dt).element_size() if len(shp) else torch.empty(0, dtype=dt).element_size()
            assert nbytes % 4 == 0
            self.offsets.append((off, nbytes))
            off += nbytes
        self.row_bytes = max(32, (off + 15) // 16 * 16)

    def __eq__(self, other):
        return isinstance(other, RecordLayout) and self.fields == other.fields


def romp_layout(calc_smpl=True, n_betas=10, n_verts=6890):
    """Record of ROMP.forward's output dict + pred_batch_ids."""
    f32, i64 = torch.float32, torch.int64
    fields = [("cam", (3,), f32), ("smpl_thetas", (72,), f32), ("smpl_betas", (n_betas,), f32), ("center_confs", (1,), f32),
              ("cam_trans", (3,), f32)]
    if calc_smpl:
        fields += [("joints", (71, 3), f32), ("pj2d_org", (71, 2), f32), ("verts", (n_verts, 3), f32)]
    fields += [("center_preds", (2,), i64), ("pred_batch_ids", (), i64)]
    return RecordLayout(fields)


def shard_range(total_frames: int, rank: int, world: int):
    """Contiguous frame range of `rank` (remainder frames go to the first ranks)."""
    base, rem = divmod(total_frames, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def _as_tensor(v, device):
    t = torch.from_numpy(v) if isinstance(v, np.ndarray) else v
    return t.to(device)


def pack_rows(layout: RecordLayout, fields: dict, count, frame_offset: int, dst: torch.Tensor, stream=None):
    """dst uint8 [1 + capacity, row_bytes] <- header + records.  `count`: int, or int32 device tensor on CUDA.
    CUDA tensors: one b200romp_pack_rows launch (no host sync); CPU tensors (gloo tests): torch copies."""
    cap = dst.shape[0] - 1
    if dst.is_cuda:
        from . import _lib
        lib = _lib.load()
        n = len(layout.fields)
        srcs, sizes = (C.c_void_p * n)(), (C.c_int * n)()
        for i, ((name, _, dt), (_, nbytes)) in enumerate(zip(layout.fields, layout.offsets)):
            t = fields[name]
            assert t.is_cuda and t.is_contiguous() and t.dtype == dt, name
            srcs[i], sizes[i] = t.data_ptr(), nbytes
        dcount = count if isinstance(count, torch.Tensor) else None
        st = torch.cuda.current_stream().cuda_stream if stream is None else stream
        _lib.check(lib.b200romp_pack_rows(srcs, sizes, n, None if dcount is None else C.c_void_p(dcount.data_ptr()),
                                          0 if dcount is not None else int(count), cap, int(frame_offset), 0,
                                          C.c_void_p(dst.data_ptr()), layout.row_bytes, C.c_void_p(st)), "pack_rows")
        return
    n = int(count)
    assert n <= cap
    hdr = torch.zeros(HEADER_WORDS, dtype=torch.int32)
    hdr[0], hdr[1], hdr[2], hdr[4] = MAGIC, n, int(frame_offset), layout.row_bytes
    dst[0, :HEADER_WORDS * 4] = hdr.view(torch.uint8)
    for (name, _, dt), (off, nbytes) in zip(layout.fields, layout.offsets):
        if n:
            src = _as_tensor(fields[name], dst.device)[:n].to(dt).contiguous().reshape(n, -1)
            dst[1:1 + n, off:off + nbytes] = src.view(torch.uint8).reshape(n, nbytes)


def unpack_rows(layout: RecordLayout, rows: torch.Tensor, frame_offsets: torch.Tensor):
    """rows uint8 [n, row_bytes] (+ per-row first-frame offsets, int64 [n]) -> output dict of tensors."""
    n = rows.shape[0]
    out = {}
    for (name, shp, dt), (off, nbytes) in zip(layout.fields, layout.offsets):
        out[name] = rows[:, off:off + nbytes].contiguous().view(dt).reshape((n,) + shp)
    out["pred_batch_ids"] = out["pred_batch_ids"] + frame_offsets.to(out["pred_batch_ids"].device)
    if "smpl_thetas" in out:
        out["global_orient"], out["body_pose"] = out["smpl_thetas"][:, :3], out["smpl_thetas"][:, 3:]
    return out


class ShardGather:
    """The collective of the frame-sharded path, pipelined for a stream of steps.

    submit() enqueues pack + ONE all-gather on a side stream (after the caller's current stream) and returns at once
    without any host synchronisation; result() of an earlier step is collected while later steps compute.  Send /
    receive buffers are persistent and double-buffered (a result stays valid until the second-next submit).
    """

    def __init__(self, world: int, layout: RecordLayout, capacity: int, group=None, rows_hint: int = 64):
        self.world, self.layout, self.capacity, self.group = world, layout, int(capacity), group
        self.cuda = dist.get_backend(group) == "nccl"
        self.device = torch.device("cuda", torch.cuda.current_device()) if self.cuda else torch.device("cpu")
        self.stream = torch.cuda.Stream() if self.cuda else None
        self.rows_hint = max(1, min(int(rows_hint), self.capacity))
        rb = layout.row_bytes
        self.send = [torch.zeros(1 + self.capacity, rb, dtype=torch.uint8, device=self.device) for _ in range(2)]
        self.recv = [None, None]
        self._hdr_host = [None, None]
        self.slot = 0
        self.collectives = 0           # all-gathers issued so far (tests / gpu_launches accounting)

    def _recv(self, slot, rows):
        need = self.world * (1 + rows)
        buf = self.recv[slot]
        if buf is None or buf.shape[0] < need:
            buf = self.recv[slot] = torch.empty(need + self.world * max(16, rows // 4), self.layout.row_bytes, dtype=torch.uint8, device=self.device)
        return buf[:need]

    def _send(self, slot, rows):
        buf = self.send[slot]
        if buf.shape[0] < 1 + rows:      # only the one-shot form (rank-local capacity) ever grows a send buffer
            big = torch.zeros(1 + rows, self.layout.row_bytes, dtype=torch.uint8, device=self.device)
            big[:buf.shape[0]] = buf
            buf = self.send[slot] = big
        return buf[:1 + rows]

    def _gather(self, slot, rows):
        recv = self._recv(slot, rows)
        dist.all_gather_into_tensor(recv.view(-1), self._send(slot, rows).view(-1), group=self.group)
        self.collectives += 1
        return recv

    def _headers_to_host(self, slot, recv, rows):
        """enqueue (on the gather stream) the copy of the world x 32-byte headers into this slot's pinned host mirror"""
        hdr = recv.view(self.world, 1 + rows, self.layout.row_bytes)[:, 0, :HEADER_WORDS * 4].contiguous()
        if self._hdr_host[slot] is None:
            t = torch.zeros(self.world, HEADER_WORDS * 4, dtype=torch.uint8)
            self._hdr_host[slot] = t.pin_memory() if self.cuda else t
        self._hdr_host[slot].copy_(hdr, non_blocking=True)

    def submit(self, fields: dict, count, frame_offset: int, rows_hint: int | None = None):
        """fields: name -> tensor [cap, ...] (device tensors of the ROMP slot, or CPU tensors under gloo);
        count: python int or int32 device tensor.  Returns a handle for counts() / result().  Everything (pack kernel, the
        all-gather, the copy of the gathered headers to pinned host memory) is enqueued on the gather's own stream after
        the caller's current stream: no host synchronisation, nothing on the legacy default stream."""
        slot, self.slot = self.slot, self.slot ^ 1
        rows = min(self.capacity, max(1, int(self.rows_hint if rows_hint is None else rows_hint)))
        if self.cuda:
            ready = torch.cuda.Event()
            ready.record()
            with torch.cuda.stream(self.stream):
                self.stream.wait_event(ready)
                pack_rows(self.layout, fields, count, frame_offset, self.send[slot], self.stream.cuda_stream)
                packed = torch.cuda.Event()
                packed.record(self.stream)          # from here on the caller may overwrite `fields`
                recv = self._gather(slot, rows)
                self._headers_to_host(slot, recv, rows)
                done = torch.cuda.Event()
                done.record(self.stream)
            for t in fields.values():
                if isinstance(t, torch.Tensor) and t.is_cuda:
                    t.record_stream(self.stream)
        else:
            pack_rows(self.layout, fields, count, frame_offset, self.send[slot])
            recv, done, packed = self._gather(slot, rows), None, None
            self._headers_to_host(slot, recv, rows)
        return dict(slot=slot, rows=rows, recv=recv, done=done, packed=packed)

    def wait(self, handle, stream=None):
        """Make `stream` (default: the current stream) wait for the gather of `handle` - device-side join, no host sync."""
        if handle["done"] is not None:
            (torch.cuda.current_stream() if stream is None else stream).wait_event(handle["done"])

    def counts(self, handle):
        """(person counts, first-frame offsets) of every rank for this step.  The one host wait: the step's gather event;
        the headers were already copied to pinned host memory on the gather stream.  If some rank held more persons than
        the rows hint, every rank sees it here and all of them repeat the gather with the exact maximum (collectively
        decided: the headers are identical on all ranks)."""
        if "counts" in handle:
            return handle["counts"], handle["offsets"]
        slot = handle["slot"]
        if handle["done"] is not None:
            handle["done"].synchronize()
        rb = self.layout.row_bytes
        hdr = self._hdr_host[slot].view(torch.int32).view(self.world, HEADER_WORDS)
        assert bool((hdr[:, 0] == MAGIC).all()) and bool((hdr[:, 4] == rb).all()), "record layout differs between ranks"
        counts, offsets = hdr[:, 1].tolist(), hdr[:, 2].tolist()
        nmax = max(counts)
        self.rows_hint = min(self.capacity, max(16, int(nmax * 1.25) + 8))      # next step's bound
        if nmax > handle["rows"]:
            if self.cuda:
                with torch.cuda.stream(self.stream):
                    handle["recv"] = self._gather(slot, nmax)
                self.stream.synchronize()
            else:
                handle["recv"] = self._gather(slot, nmax)
            handle["rows"] = nmax
        handle["counts"], handle["offsets"] = counts, offsets
        return counts, offsets

    def result(self, handle, to_numpy=False):
        """Every rank's persons in global frame order as a dict of tensors on the gather device (or numpy arrays), or None
        when nobody was detected anywhere.  The concatenation runs on the gather stream."""
        counts, offsets = self.counts(handle)
        if max(counts) == 0:
            return None
        rows, recv = handle["rows"], handle["recv"]
        v = recv.view(self.world, 1 + rows, self.layout.row_bytes)
        offs = torch.cat([torch.full((c,), o, dtype=torch.int64) for c, o in zip(counts, offsets) if c])
        if self.cuda:
            with torch.cuda.stream(self.stream):
                out = unpack_rows(self.layout, torch.cat([v[r, 1:1 + c] for r, c in enumerate(counts) if c], 0), offs.to(self.device, non_blocking=True))
                if to_numpy:
                    out = {k: t.cpu() for k, t in out.items()}
            self.stream.synchronize()
        else:
            out = unpack_rows(self.layout, torch.cat([v[r, 1:1 + c] for r, c in enumerate(counts) if c], 0), offs)
        if to_numpy:
            out = {k: t.numpy() for k, t in out.items()}
        return out


def all_gather_outputs(out, frame_offset: int, world: int, to_numpy: bool = True, group=None, layout: RecordLayout | None = None,
                       rows_hint: int = 64):
    """One-shot convenience form: `out` is a result dict of ROMP.forward_batch (numpy or tensors) or None.  All ranks
    end up with the outputs of every rank, ordered by (global frame asc, score desc).  `layout` defaults to the ROMP
    record matching `out` (a rank whose `out` is None must pass the layout of its configuration unless it is the
    default one) - it is never negotiated at run time.  `rows_hint` must be the same on every rank."""
    if layout is None:
        if out is None:
            layout = romp_layout()
        else:
            nb = int(np.asarray(out["smpl_betas"]).shape[1])
            layout = romp_layout("verts" in out, nb, int(np.asarray(out["verts"]).shape[1]) if "verts" in out else 6890)
    n = 0 if out is None else int(np.asarray(out["cam"]).shape[0]) if not isinstance(out["cam"], torch.Tensor) else int(out["cam"].shape[0])
    g = ShardGather(world, layout, capacity=max(n, rows_hint), group=group, rows_hint=rows_hint)
    fields = {}
    if out is not None:
        fields = {name: _as_tensor(out[name], g.device).to(dt).contiguous() for name, _, dt in layout.fields}
    else:
        fields = {name: torch.zeros((1,) + shp, dtype=dt, device=g.device) for name, shp, dt in layout.fields}
    return g.result(g.submit(fields, n, frame_offset), to_numpy=to_numpy)
