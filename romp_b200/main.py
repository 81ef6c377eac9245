"""Drop-in call surface of ``simple_romp``'s ROMP for the per-frame inference hot path.

Mirrors simple_romp/romp/main.py: ``romp_settings`` (:17-60, same flags and defaults, minus the
import-time download side effects), ``class ROMP`` (:64-176) with ``ROMP(settings)(image_bgr) -> dict | None``
and the same output dict.  ``forward_batch`` is the batched entry point the reference
lacks (its ``forward`` takes one image per call, main.py:106-107).

Python/PyTorch here only allocates device buffers, owns the CUDA stream and moves bytes; every per-frame
FLOP runs in libb200romp.so (hand-written sm_100a CUDA).  There is no CPU or PyTorch compute fallback:
without the library or without a GPU, constructing ``ROMP`` raises.
"""
from __future__ import annotations

import argparse
import ctypes as C
import os
import os.path as osp
import sys

import numpy as np
import torch

from . import _lib, graph
from ._lib import BF16, F32, U8

MAX_PERSON = 64          # CenterMap.max_person, post_parser.py:11
N_PARAMS = 145           # 3 cam + 22*6 rot6d + 10 betas, model.py:429


def romp_settings(input_args=sys.argv[1:]):
    """Same flags/defaults as the reference's romp_settings (main.py:17-60); no downloads, no prints."""
    parser = argparse.ArgumentParser(description="ROMP (B200-native hot path)")
    parser.add_argument("-m", "--mode", type=str, default="image")
    parser.add_argument("-i", "--input", type=str, default=None)
    parser.add_argument("-o", "--save_path", type=str, default=osp.join(osp.expanduser("~"), "ROMP_results"))
    parser.add_argument("--GPU", type=int, default=0)
    parser.add_argument("--onnx", action="store_true")
    parser.add_argument("-t", "--temporal_optimize", action="store_true")
    parser.add_argument("--center_thresh", type=float, default=0.25)
    parser.add_argument("--show_largest", action="store_true")
    parser.add_argument("-sc", "--smooth_coeff", type=float, default=3.0)
    parser.add_argument("--calc_smpl", action="store_false")
    parser.add_argument("--render_mesh", action="store_true")
    parser.add_argument("--renderer", type=str, default="sim3dr")
    parser.add_argument("--show", action="store_true")
    parser.add_argument("--show_items", type=str, default="mesh")
    parser.add_argument("--save_video", action="store_true")
    parser.add_argument("--frame_rate", type=int, default=24)
    parser.add_argument("--smpl_path", type=str, default=osp.join(osp.expanduser("~"), ".romp", "SMPL_NEUTRAL.pth"))
    parser.add_argument("--model_path", type=str, default=osp.join(osp.expanduser("~"), ".romp", "ROMP.pkl"))
    parser.add_argument("--model_onnx_path", type=str, default=osp.join(osp.expanduser("~"), ".romp", "ROMP.onnx"))
    parser.add_argument("--root_align", type=bool, default=False)
    parser.add_argument("--webcam_id", type=int, default=0)
    # --- additions of this implementation (the reference has no equivalents)
    parser.add_argument("--precision", type=str, default="bf16", choices=["bf16", "tf32", "fp32"],
                        help="conv arithmetic: bf16 tensor cores (fast), tf32 tensor cores on fp32 tensors (the reference's "
                             "default GPU arithmetic, cudnn.allow_tf32) or fp32 CUDA cores (strict parity)")
    parser.add_argument("--max_batch", type=int, default=64, help="largest batch forward_batch will be given")
    parser.add_argument("--backbone", type=str, default="hrnet32", choices=["hrnet32", "resnet50"],
                        help="hrnet32 = simple_romp's ROMPv1 (model.py); resnet50 = the training package's ResNet-50 variant "
                             "(romp/lib/models/resnet_50.py; BASELINE cfg1), state dict with the same head keys")
    parser.add_argument("--cam_trans", type=str, default="lsq", choices=["lsq", "pnp"],
                        help="cam_trans estimator: lsq = closed-form least squares on the GPU (the reference's fallback, "
                             "utils.py:347-389); pnp = the reference's default cv2.solvePnPRansac per person on the host")
    args = parser.parse_args(input_args)
    if not os.path.exists(args.smpl_path):
        alt = args.smpl_path.replace("SMPL_NEUTRAL.pth", "smpl_packed_info.pth")   # main.py:50-52
        if os.path.exists(alt):
            args.smpl_path = alt
    return args


def padding_image(image):
    """utils.py:16-24."""
    h, w = image.shape[:2]
    side = max(h, w)
    pad = np.zeros((side, side, 3), dtype=np.uint8)
    top, left = int((side - h) // 2), int((side - w) // 2)
    pad[top:top + h, left:left + w] = image
    return pad, np.array([top, top + h, left, left + w, h, w], dtype=np.float32)


def img_preprocess(image, input_size=512):
    """utils.py:26-30 restated on the host with OpenCV exactly like the reference (BGR->RGB, square zero pad, cubic resize;
    returns uint8 [1,512,512,3]).  ``ROMP.forward`` does NOT use it - it runs ``ROMP.preprocess`` (one CUDA kernel); this
    host mirror exists for tests and for callers that batch pre-sized frames themselves."""
    import cv2
    image = cv2.cvtColor(image, cv2.COLOR_BGR2RGB)
    pad, info = padding_image(image)
    return cv2.resize(pad, (input_size, input_size), interpolation=cv2.INTER_CUBIC)[None], info


INVALID_TRANS = np.array([-1.0, -1.0, -1.0], np.float32)      # utils.py:295


def _translation_lsq_host(S, p2, focal, center):
    """estimate_translation_np (utils.py:347-389, the reference's own fallback), all joints valid, fp64 on the host."""
    n = S.shape[0]
    Z = np.reshape(np.tile(S[:, 2], (2, 1)).T, -1)
    XY = np.reshape(S[:, :2], -1)
    O_ = np.tile(center, n)
    Fv = np.tile(np.array([focal, focal], np.float64), n)
    w = np.ones(2 * n)
    Q = np.array([Fv * np.tile(np.array([1, 0]), n), Fv * np.tile(np.array([0, 1]), n), O_ - np.reshape(p2, -1)]).T
    c = (np.reshape(p2, -1) - O_) * Z - Fv * XY
    W = np.diagflat(w)
    Q, c = W @ Q, W @ c
    return np.linalg.solve(Q.T @ Q, Q.T @ c)


def estimate_translation_pnp(joints, cam, focal_length=443.4, img_size=512.0):
    """``--cam_trans pnp``: the reference's DEFAULT cam_trans (convert_cam_to_3d_trans2 post_parser.py:96-101 ->
    estimate_translation utils.py:391-436 -> estimate_translation_cv2 :331-345): per person
    cv2.solvePnPRansac(EPnP, reprojectionError 20, 100 iterations) on the 24 SMPL joints against their weak-perspective
    projection (pj2d + 1) * 256, K = diag(443.4) with principal point 256; INVALID_TRANS when RANSAC finds no inliers,
    the closed-form least squares (:347-389) when OpenCV raises.  Host side (OpenCV), like the reference; the device
    default (``--cam_trans lsq``) is that closed form for every person (b200romp_project)."""
    import cv2
    joints, cam = np.asarray(joints, np.float32), np.asarray(cam, np.float32)
    j3 = np.ascontiguousarray(joints[:, :24])
    p2 = (j3[:, :, :2] * cam[:, None, 0:1] + cam[:, None, 1:3] + 1.0) * (img_size / 2.0)      # batch_orth_proj utils.py:309-315
    K = np.eye(3)
    K[0, 0] = K[1, 1] = focal_length
    K[:2, 2] = img_size // 2
    out = np.zeros((len(j3), 3), np.float32)
    for i in range(len(j3)):
        try:
            _, _, tvec, inliers = cv2.solvePnPRansac(j3[i], p2[i].astype(np.float32), K, None, flags=cv2.SOLVEPNP_EPNP,
                                                     reprojectionError=20, iterationsCount=100)
            out[i] = INVALID_TRANS if inliers is None else tvec[:, 0]
        except Exception:
            out[i] = _translation_lsq_host(j3[i].astype(np.float64), p2[i].astype(np.float64), focal_length,
                                           np.array([img_size / 2.0, img_size / 2.0]))
    return out


def _ptr(t):
    return C.c_void_p(t.data_ptr())


class SMPLParser:
    """Seam S3 (post_parser.SMPL_parser, post_parser.py:116-125) on libb200romp's fused SMPL kernels."""

    def __init__(self, pack, device, n_betas=10, shape_key="shapedirs"):
        self.lib = _lib.load()
        g = lambda k, dt: np.ascontiguousarray(pack[k].detach().cpu().numpy() if isinstance(pack[k], torch.Tensor) else pack[k], dtype=dt)
        self._keep = [g("v_template", np.float32), g(shape_key, np.float32), g("posedirs", np.float32),
                      g("J_regressor", np.float32), g("weights", np.float32), g("kintree_table", np.int64),
                      g("extra_joints_index", np.int64), g("J_regressor_extra9", np.float32),
                      g("J_regressor_h36m17", np.float32)]
        k = self._keep
        assert k[0].shape == (6890, 3) and k[1].shape == (6890, 3, n_betas) and k[2].shape == (207, 20670)
        fp, lp = C.POINTER(C.c_float), C.POINTER(C.c_longlong)
        a = lambda x, t: x.ctypes.data_as(t)
        self.h = self.lib.b200romp_smpl_create(device, n_betas, a(k[0], fp), a(k[1], fp), a(k[2], fp), a(k[3], fp),
                                               a(k[4], fp), a(k[5], lp), a(k[6], lp), a(k[7], fp), a(k[8], fp))
        if not self.h:
            raise RuntimeError("b200romp_smpl_create: " + self.lib.b200romp_last_error().decode())
        self.n_betas = n_betas
        self.ws_floats = self.lib.b200romp_smpl_workspace_floats()
        self.faces = torch.from_numpy(np.ascontiguousarray(g("f", np.int64))) if "f" in pack else None

    def forward(self, betas, thetas, n, d_count, root_align, ws, verts, joints, stream):
        _lib.check(self.lib.b200romp_smpl_forward(self.h, _ptr(betas), betas.shape[1], _ptr(thetas), n,
                                                  None if d_count is None else _ptr(d_count), int(root_align),
                                                  _ptr(ws), _ptr(verts), _ptr(joints), C.c_void_p(stream)), "smpl_forward")

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.lib.b200romp_smpl_destroy(self.h)
        except Exception:
            pass


class MapsModule(torch.nn.Module):
    """``ROMP.model``: seam S1 as a module, like the reference's ``self.model`` (ROMPv1 wrapped in nn.DataParallel,
    main.py:74-77): ``model(frames)`` with frames [B,512,512,3] (uint8 / float32 0..255, device) returns
    ``(center_maps [B,1,64,64], params_maps [B,145,64,64])``.  Difference to the reference, by design: the cam-scale
    ``1.1**x`` of main.py:113 is already applied to ``params_maps[:,0]`` (it is fused into the head conv's epilogue)."""

    def __init__(self, owner):
        super().__init__()
        object.__setattr__(self, "_owner", owner)       # not a sub-module: no parameter / state-dict recursion

    def forward(self, frames):
        o = self._owner
        o._after_producers(frames)
        with torch.cuda.stream(o.stream):
            c, p = o.run_maps(frames.contiguous())
        torch.cuda.current_stream(o.tdevice).wait_stream(o.stream)
        return c, p


class ROMP(torch.nn.Module):
    """``ROMP(settings)(image_bgr)`` - same contract as simple_romp/romp/main.py:64-176."""

    def __init__(self, romp_settings, state_dict=None, smpl_pack=None):
        super().__init__()
        self.settings = s = romp_settings
        if not torch.cuda.is_available() or s.GPU < 0:
            raise RuntimeError("romp_b200.ROMP needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        for flag in ("render_mesh", "show"):
            if getattr(s, flag, False):
                raise NotImplementedError(f"--{flag} is outside the B200 hot path (SURVEY.md section 2: out of scope)")
        if getattr(s, "onnx", False):
            raise NotImplementedError("--onnx: onnxruntime is not available offline; seam S1 is ROMP.model (MapsModule), the same "
                                      "seam the reference swaps for its ONNX session (main.py:78-91,108-110)")
        if getattr(s, "show_largest", False) and not getattr(s, "temporal_optimize", False):
            raise NotImplementedError("--show_largest only selects the smoothed person of --temporal_optimize (main.py:128-134)")
        self.lib = _lib.load()
        self.device_index = int(s.GPU)
        self.tdevice = torch.device("cuda", self.device_index)
        torch.cuda.set_device(self.tdevice)
        self.precision = getattr(s, "precision", "bf16")
        self.max_batch = int(getattr(s, "max_batch", 64))
        if state_dict is None:
            state_dict = torch.load(s.model_path, map_location="cpu")          # main.py:75
        self._nets = {}
        self._state_dict = state_dict
        self.stream = torch.cuda.Stream(device=self.tdevice)
        self.calc_smpl = bool(s.calc_smpl)
        if self.calc_smpl:
            if smpl_pack is None:
                smpl_pack = torch.load(s.smpl_path, map_location="cpu")        # smpl.py:41
            self.smpl = SMPLParser(smpl_pack, self.device_index)
        self._alloc(self.max_batch)
        self.model = MapsModule(self)
        self.temporal = None
        if getattr(s, "temporal_optimize", False):                                 # main.py:117-125
            from .temporal import TemporalState
            self.temporal = TemporalState(bool(getattr(s, "show_largest", False)))
            self._tracks = self.lib.b200romp_tracks_create(self.device_index, self.temporal.n_slots)
            if not self._tracks:
                raise RuntimeError("b200romp_tracks_create: " + self.lib.b200romp_last_error().decode())
            self._slot_host = torch.zeros(self.cap, dtype=torch.int32).pin_memory()
            self._slot_dev = torch.zeros(self.cap, dtype=torch.int32, device=self.tdevice)

    # ------------------------------------------------------------------------------------------
    def _net(self, in_dtype):
        if in_dtype not in self._nets:
            build = graph.build_romp_resnet50 if getattr(self.settings, "backbone", "hrnet32") == "resnet50" else graph.build_romp
            self._nets[in_dtype] = build(self._state_dict, self.device_index, self.precision, in_dtype, self.max_batch)
        return self._nets[in_dtype]

    def _alloc(self, B):
        dev, cap = self.tdevice, B * MAX_PERSON
        f32, i64 = torch.float32, torch.int64
        self.cap = cap
        z = lambda *shape, dtype=f32: torch.zeros(*shape, dtype=dtype, device=dev)
        # shared scratch: only touched by kernels on self.stream, in order
        self.shared = dict(
            center_maps=z(B, 1, 64, 64), params_maps=z(B, N_PARAMS, 64, 64), flat_inds=z(cap, dtype=i64),
            params_pred=z(cap, N_PARAMS),
            parse_ws=torch.zeros(int(self.lib.b200romp_parse_workspace_bytes(B)), dtype=torch.uint8, device=dev))
        if self.calc_smpl:
            self.shared["smpl_ws"] = z(cap, self.smpl.ws_floats)
        # two slots of per-person outputs (+ pinned host mirrors) so that batch i+1 can be computed while batch i
        # is still being read back (forward_batches)
        self.slots = []
        for _ in range(2):
            d = dict(count=z(1, dtype=torch.int32), batch_ids=z(cap, dtype=i64), center_confs=z(cap, 1), cam=z(cap, 3),
                     thetas=z(cap, 72), betas=z(cap, 10), center_preds=z(cap, 2, dtype=i64), cam_trans=z(cap, 3),
                     pj2d_org=z(cap, 71, 2))
            if self.calc_smpl:
                d.update(verts=z(cap, 6890, 3), joints=z(cap, 71, 3))
            self.slots.append(dict(dev=d, host=None, count_host=torch.zeros(1, dtype=torch.int32).pin_memory(),
                                   done=torch.cuda.Event(), frames={}, h2d=torch.cuda.Event()))
        self._slot = 0
        self._raw_host = self._raw_dev = None
        self.copy_stream = torch.cuda.Stream(device=dev)
        self.d2h_stream = torch.cuda.Stream(device=dev)

    @property
    def buf(self):
        """device buffers of the slot used by the most recent batch (+ the shared maps)"""
        return {**self.shared, **self.slots[self._slot]["dev"]}

    def _host(self, slot):
        if slot["host"] is None:      # pinned mirrors, allocated on first read-back
            slot["host"] = {k: torch.zeros(v.shape, dtype=v.dtype).pin_memory() for k, v in slot["dev"].items() if k != "count"}
        return slot["host"]

    # ------------------------------------------------------------------------------------------
    @torch.no_grad()
    def run_maps(self, frames_dev):
        """Seam S1 only: frames [B,512,512,3] uint8/float32 on the device -> (center_maps, params_maps) views."""
        B = frames_dev.shape[0]
        assert frames_dev.is_cuda and frames_dev.is_contiguous() and tuple(frames_dev.shape[1:]) == (512, 512, 3)
        assert B <= self.max_batch, f"batch {B} > max_batch {self.max_batch}"
        in_dtype = {torch.uint8: U8, torch.float32: F32}[frames_dev.dtype]
        nb, io = self._net(in_dtype)
        lib, sp = self.lib, C.c_void_p(self.stream.cuda_stream)
        _lib.check(lib.b200romp_net_bind(nb.net, io["frames"], _ptr(frames_dev)))
        _lib.check(lib.b200romp_net_bind(nb.net, io["center_maps"], _ptr(self.shared["center_maps"])))
        _lib.check(lib.b200romp_net_bind(nb.net, io["params_maps"], _ptr(self.shared["params_maps"])))
        _lib.check(lib.b200romp_net_run(nb.net, B, sp), "net_run")
        return self.shared["center_maps"][:B], self.shared["params_maps"][:B]

    @torch.no_grad()
    def run_parse(self, B, center_override=None, slot=None):
        """Seam S2 on the maps currently in the buffers (enqueued on self.stream, no host sync)."""
        sh, lib, sp = self.shared, self.lib, C.c_void_p(self.stream.cuda_stream)
        b = (self.slots[self._slot] if slot is None else slot)["dev"]
        center = sh["center_maps"] if center_override is None else center_override
        _lib.check(lib.b200romp_parse(_ptr(center), _ptr(sh["params_maps"]), B, 64, 10, float(self.settings.center_thresh),
                                      self.cap, _ptr(b["count"]), _ptr(b["batch_ids"]), _ptr(sh["flat_inds"]),
                                      _ptr(b["center_confs"]), _ptr(sh["params_pred"]), _ptr(b["cam"]), _ptr(b["thetas"]),
                                      _ptr(b["betas"]), _ptr(b["center_preds"]), _ptr(sh["parse_ws"]), sp), "parse")

    @torch.no_grad()
    def run_smpl_project(self, cap, offsets, slot=None, count_on_device=True):
        """Seams S3-S4 for up to ``cap`` persons (the device-side count limits it further unless count_on_device=False)."""
        sh, lib, sp = self.shared, self.lib, C.c_void_p(self.stream.cuda_stream)
        b = (self.slots[self._slot] if slot is None else slot)["dev"]
        cnt = b["count"] if count_on_device else None
        cp = None if cnt is None else _ptr(cnt)
        off = (C.c_float * 6)(*[float(v) for v in offsets])
        if self.calc_smpl:
            self.smpl.forward(b["betas"], b["thetas"], cap, cnt, self.settings.root_align, sh["smpl_ws"],
                              b["verts"], b["joints"], self.stream.cuda_stream)
            _lib.check(lib.b200romp_project(_ptr(b["joints"]), None, _ptr(b["cam"]), cap, cp, off,
                                            _ptr(b["pj2d_org"]), None, None, _ptr(b["cam_trans"]), sp), "project")
        else:   # without SMPL the reference keeps the weak-perspective translation of main.py:166
            _lib.check(lib.b200romp_project(_ptr(b["cam"]), None, _ptr(b["cam"]), cap, cp, off, None, None,
                                            _ptr(b["cam_trans"]), None, sp), "project")

    @torch.no_grad()
    def run_post(self, B, offsets, center_override=None, slot=None):
        """Seams S2-S4 on the maps currently in the buffers; everything enqueued on self.stream, no host sync."""
        self.run_parse(B, center_override, slot)
        self.run_smpl_project(B * MAX_PERSON, offsets, slot)

    def _views(self, src, n):
        out = {"cam": src["cam"][:n], "global_orient": src["thetas"][:n, :3], "body_pose": src["thetas"][:n, 3:],
               "smpl_betas": src["betas"][:n], "smpl_thetas": src["thetas"][:n], "center_preds": src["center_preds"][:n],
               "center_confs": src["center_confs"][:n], "cam_trans": src["cam_trans"][:n]}
        if self.calc_smpl:
            out.update(verts=src["verts"][:n], joints=src["joints"][:n], pj2d_org=src["pj2d_org"][:n])
        out["pred_batch_ids"] = src["batch_ids"][:n]
        return out

    def record_layout(self):
        """Fixed per-person record of this configuration for the sharded path's single all-gather (shard.ShardGather)."""
        from . import shard
        return shard.romp_layout(self.calc_smpl, 10)

    def record_fields(self, slot=None):
        """name -> device tensor [cap, ...] of the slot used by the most recent batch, keyed like record_layout()."""
        d = (self.slots[self._slot] if slot is None else slot)["dev"]
        f = {"cam": d["cam"], "smpl_thetas": d["thetas"], "smpl_betas": d["betas"], "center_confs": d["center_confs"],
             "cam_trans": d["cam_trans"], "center_preds": d["center_preds"], "pred_batch_ids": d["batch_ids"]}
        if self.calc_smpl:
            f.update(joints=d["joints"], pj2d_org=d["pj2d_org"], verts=d["verts"])
        return f, d["count"]

    def collect(self, to_numpy=True, slot=None, stream=None):
        """The single host sync of a batch: person count, then D2H of the N valid rows (utils.py:32-41) into pinned
        host mirrors.  The returned numpy arrays are views of those mirrors: valid until the slot is reused, i.e.
        until the second-next batch (copy them if they must live longer)."""
        slot = self.slots[self._slot] if slot is None else slot
        stream = self.stream if stream is None else stream
        with torch.cuda.stream(stream):
            slot["count_host"].copy_(slot["dev"]["count"], non_blocking=True)
        stream.synchronize()
        n = int(slot["count_host"].item())
        if n == 0:
            return None
        if not to_numpy:
            return self._views(slot["dev"], n)
        host = self._host(slot)
        with torch.cuda.stream(stream):
            for k, v in slot["dev"].items():
                if k != "count":
                    host[k][:n].copy_(v[:n], non_blocking=True)
        stream.synchronize()
        return {k: v.numpy() for k, v in self._views(host, n).items()}

    # ------------------------------------------------------------------------------------------
    def _after_producers(self, *tensors):
        """Device-resident inputs were produced on the caller's current stream; our kernels run on self.stream.  Order
        them (no host sync) and keep the allocator from recycling the inputs while self.stream still reads them."""
        cur = torch.cuda.current_stream(self.tdevice)
        waited = False
        for t in tensors:
            if isinstance(t, torch.Tensor) and t.is_cuda:
                if not waited and cur != self.stream:
                    self.stream.wait_stream(cur)
                    waited = True
                t.record_stream(self.stream)

    def _staging(self, slot, dtype, B):
        """persistent device frame buffer of a slot: a stable pointer keeps the conv graph's CUDA-graph cache at one entry"""
        key = (dtype, B)
        if key not in slot["frames"]:
            slot["frames"][key] = torch.empty((B, 512, 512, 3), dtype=dtype, device=self.tdevice)
        return slot["frames"][key]

    @torch.no_grad()
    def forward_batch(self, frames, offsets=None, to_numpy=True, center_override=None, own=True):
        """frames: [B,512,512,3] RGB uint8/float32 (torch tensor, pinned host or device, or numpy), already
        padded+resized like img_preprocess.  Returns the reference's dict plus ``pred_batch_ids`` or None.
        Device-resident ``frames`` / ``center_override`` may come straight from a producer on the caller's current
        stream.  ``own=True`` (default) returns arrays that own their memory; ``own=False`` returns views of the pinned
        read-back mirrors, valid until the second-next batch."""
        if isinstance(frames, np.ndarray):
            frames = torch.from_numpy(frames)
        B = frames.shape[0]
        self._slot ^= 1
        slot = self.slots[self._slot]
        self._after_producers(frames, center_override)
        with torch.cuda.stream(self.stream):
            fd = self._staging(slot, frames.dtype, B)
            fd.copy_(frames, non_blocking=True)
            self.run_maps(fd)
            self.run_post(B, offsets if offsets is not None else [0, 512, 0, 512, 512, 512], center_override)
            slot["done"].record(self.stream)
        out = self.collect(to_numpy)
        if out is not None and to_numpy and own:
            out = {k: np.array(v) for k, v in out.items()}
            if getattr(self.settings, "cam_trans", "lsq") == "pnp" and self.calc_smpl:
                out["cam_trans"] = estimate_translation_pnp(out["joints"], out["cam"])
        return out

    @torch.no_grad()
    def forward_batches(self, batches, offsets=None, center_override=None, to_numpy=True, gather=None, frame_offset=0):
        """Pipelined streaming over an iterable of host frame batches (video): yields one result dict (or None) per
        batch, in order.  H2D of batch i+1 (copy stream) and D2H of batch i-1 (read-back stream) overlap the
        kernels of batch i; the host waits only for the H2D copy of the batch it just handed over (so a caller may
        refill one pinned buffer in a decode loop) and for the person count of an already finished batch.  With
        ``to_numpy=True`` the yielded arrays are views of pinned mirrors, valid until the second-next batch is yielded.
        Sharded (multi-GPU) use: pass ``gather`` (a shard.ShardGather built on ``record_layout()``) and this rank's
        ``frame_offset``; each batch's records are then packed and all-gathered on the gather's side stream right after
        its kernels, and the generator yields ``(own_result, gather_handle)`` pairs (own shard read back to this rank's
        host; ``gather.result(handle)`` gives every rank's persons on the device)."""
        off = offsets if offsets is not None else [0, 512, 0, 512, 512, 512]
        pending = None
        self._after_producers(center_override)
        for frames in batches:
            if isinstance(frames, np.ndarray):
                frames = torch.from_numpy(frames)
            B = frames.shape[0]
            self._slot ^= 1
            slot = self.slots[self._slot]
            fd = self._staging(slot, frames.dtype, B)
            if frames.is_cuda:
                self.copy_stream.wait_stream(torch.cuda.current_stream(self.tdevice))
                frames.record_stream(self.copy_stream)
            with torch.cuda.stream(self.copy_stream):
                self.copy_stream.wait_event(slot["done"])          # the slot's previous kernels no longer read fd
                fd.copy_(frames, non_blocking=True)
                slot["h2d"].record(self.copy_stream)
            with torch.cuda.stream(self.stream):
                self.stream.wait_event(slot["h2d"])
                if slot.get("packed") is not None:
                    self.stream.wait_event(slot["packed"])          # the gather's pack kernel has read the slot's previous results
                self.run_maps(fd)
                self.run_post(B, off, center_override, slot)
                slot["done"].record(self.stream)
                if gather is not None:
                    fields, count = self.record_fields(slot)
                    slot["gather"] = gather.submit(fields, count, frame_offset)
                    slot["packed"] = slot["gather"]["packed"]
            if not frames.is_cuda:
                slot["h2d"].synchronize()     # the caller may refill its (single) host buffer as soon as we yield / pull the next batch
            if pending is not None:
                res = self._read_back(pending, to_numpy)
                yield (res, pending["gather"]) if gather is not None else res
            pending = slot
        if pending is not None:
            res = self._read_back(pending, to_numpy)
            yield (res, pending["gather"]) if gather is not None else res

    def _read_back(self, slot, to_numpy=True):
        self.d2h_stream.wait_event(slot["done"])
        return self.collect(to_numpy, slot, self.d2h_stream)   # device views stay valid until the slot's next batch

    @torch.no_grad()
    def preprocess(self, image, out=None):
        """img_preprocess (utils.py:26-30) on the GPU: raw HxWx3 uint8 BGR image (numpy / host tensor / device tensor) ->
        ``out`` [512,512,3] uint8 RGB on the device (allocated when None) + pad info [top,bottom,left,right,h,w].
        One kernel (b200romp_preprocess_bgr) on self.stream; the only host work is the H2D copy of the raw image."""
        img = torch.from_numpy(np.ascontiguousarray(image)) if isinstance(image, np.ndarray) else image
        assert img.dtype == torch.uint8 and img.dim() == 3 and img.shape[2] == 3, "image must be HxWx3 uint8 (BGR)"
        h, w = int(img.shape[0]), int(img.shape[1])
        if out is None:
            out = torch.empty((512, 512, 3), dtype=torch.uint8, device=self.tdevice)
        if img.is_cuda:
            self._after_producers(img)
            raw = img.contiguous()
        else:
            n = img.numel()
            if self._raw_host is None or self._raw_host.numel() < n:
                self._raw_host = torch.empty(max(n, 1 << 22), dtype=torch.uint8).pin_memory()
                self._raw_dev = torch.empty(self._raw_host.numel(), dtype=torch.uint8, device=self.tdevice)
            self.stream.synchronize()                       # the previous image's H2D has left the pinned staging buffer
            self._raw_host[:n].copy_(img.reshape(-1))
            raw = self._raw_dev[:n]
            with torch.cuda.stream(self.stream):
                raw.copy_(self._raw_host[:n], non_blocking=True)
        pad = (C.c_float * 6)()
        _lib.check(self.lib.b200romp_preprocess_bgr(_ptr(raw), h, w, 3 * w, 512, _ptr(out), pad,
                                                    C.c_void_p(self.stream.cuda_stream)), "preprocess_bgr")
        return out, np.array(list(pad), dtype=np.float32)

    @torch.no_grad()
    def forward(self, image, signal_ID=0, **kwargs):
        """image: HxWx3 uint8 BGR (cv2.imread).  main.py:160-176; preprocessing, model, parse, SMPL and projection all run
        on the GPU - OpenCV is not involved."""
        self._slot ^= 1
        slot = self.slots[self._slot]
        fd = self._staging(slot, torch.uint8, 1)
        _, pad_info = self.preprocess(image, out=fd[0])
        if self.temporal is not None:
            return self._forward_temporal(fd, pad_info, slot, signal_ID)
        with torch.cuda.stream(self.stream):
            self.run_maps(fd)
            self.run_post(1, pad_info)
            slot["done"].record(self.stream)
        out = self.collect(True)
        if out is None:
            print("None person detected")                                       # post_parser.py:139
            return None
        out.pop("pred_batch_ids")
        out = {k: np.array(v) for k, v in out.items()}                           # arrays own their memory like the reference's
        if getattr(self.settings, "cam_trans", "lsq") == "pnp" and self.calc_smpl:
            out["cam_trans"] = estimate_translation_pnp(out["joints"], out["cam"])
        return out


    @torch.no_grad()
    def _forward_temporal(self, fd, pad_info, slot, signal_ID):
        """forward() with --temporal_optimize (main.py:164-165 -> temporal_optimization :127-157): parse, associate the
        detections with tracks on the host (needs the cams: one small D2H, like the reference), One-Euro smoothing of
        thetas / betas / cam on the device, then SMPL + projection on the smoothed parameters."""
        b, lib, sp = slot["dev"], self.lib, C.c_void_p(self.stream.cuda_stream)
        with torch.cuda.stream(self.stream):
            self.run_maps(fd)
            self.run_parse(1, None, slot)
            slot["count_host"].copy_(b["count"], non_blocking=True)
        self.stream.synchronize()
        n = int(slot["count_host"].item())
        if n == 0:
            print("None person detected")
            return None
        with torch.cuda.stream(self.stream):
            cams = b["cam"][:n].cpu().numpy()
            raw_thetas = b["thetas"][:n].clone()          # global_orient / body_pose keep the UNsmoothed values (main.py:148-153)
        slots, track_ids, reset = self.temporal.assign(cams, signal_ID)
        for sl in reset:
            _lib.check(lib.b200romp_tracks_reset(self._tracks, int(sl), sp), "tracks_reset")
        self._slot_host[:n].copy_(torch.from_numpy(slots))
        n_smpl = n
        with torch.cuda.stream(self.stream):
            self._slot_dev[:n].copy_(self._slot_host[:n], non_blocking=True)
            _lib.check(lib.b200romp_one_euro_smooth(self._tracks, _ptr(self._slot_dev), n, None, _ptr(b["thetas"]), _ptr(b["betas"]),
                                                    10, 10, _ptr(b["cam"]), float(self.settings.smooth_coeff), 30.0, sp), "one_euro")
            if self.temporal.show_largest:                 # only the largest person goes on (main.py:129-134)
                k = int(np.argmax(cams[:, 0]))
                for key in ("thetas", "betas", "cam"):
                    b[key][0].copy_(b[key][k].clone())
                n_smpl = 1
            self.run_smpl_project(n_smpl, pad_info, slot, count_on_device=False)
            slot["done"].record(self.stream)
            host = self._host(slot)
            for key, v in b.items():
                if key != "count":
                    host[key][:n].copy_(v[:n], non_blocking=True)
            raw_host = raw_thetas.cpu()
        self.stream.synchronize()
        v = {k: np.array(t.numpy()) for k, t in self._views(host, n).items()}
        v.pop("pred_batch_ids")
        v["global_orient"], v["body_pose"] = raw_host.numpy()[:, :3].copy(), raw_host.numpy()[:, 3:].copy()
        for key in ("smpl_thetas", "smpl_betas", "cam", "cam_trans", "verts", "joints", "pj2d_org"):
            if key in v:
                v[key] = v[key][:n_smpl]
        if track_ids is not None:
            v["track_ids"] = track_ids                      # main.py:156
        return v

    def __del__(self):
        try:
            if getattr(self, "_tracks", None):
                self.lib.b200romp_tracks_destroy(self._tracks)
        except Exception:
            pass


default_settings = None   # the reference evaluates romp_settings([]) at import (main.py:62); we do not


def main():
    import cv2
    args = romp_settings()
    romp = ROMP(args)
    if args.mode != "image":
        raise NotImplementedError("video/webcam loops are outside the hot path; call ROMP.forward per frame")
    outputs = romp(cv2.imread(args.input))
    if outputs is not None:
        os.makedirs(args.save_path, exist_ok=True)
        np.savez(osp.join(args.save_path, osp.splitext(osp.basename(args.input))[0] + ".npz"), results=outputs)


if __name__ == "__main__":
    main()
