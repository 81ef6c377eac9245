"""Drop-in call surface of ``simple_romp``'s BEV (bev/main.py) on the B200 hot path.

``bev_settings`` mirrors bev/main.py:27-87 (same flags and defaults, including the quirk that ``crowd`` defaults to
True and therefore overrides the thresholds with ``long_conf_dict``), ``BEV(settings)(image_bgr) -> dict | None``
mirrors bev/main.py:91-181 for normal (non-panoramic) images; ``forward_batch`` is the batched entry point.
Per-frame semantics of the two post filters (bev/post_parser.py:167-222) are preserved by applying them per
``pred_batch_ids`` group on the device.  Python only allocates buffers and sequences library calls.
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
from .main import MAX_PERSON, SMPLParser, _ptr, img_preprocess

conf_dict = {1: [0.25, 20, 2], 2: [0.1, 20, 1.6]}                    # bev/main.py:24-25
long_conf_dict = {1: [0.12, 20, 1.5, 0.46], 2: [0.08, 20, 1.6, 0.8]}
model_dict = {1: "BEV_ft_agora.pth", 2: "BEV.pth"}
N_PARAMS = 146                                                       # 3 cam + 22*6 + 11 betas, bev/model.py:116


def bev_settings(input_args=sys.argv[1:]):
    """Same flags/defaults as bev/main.py:27-87; no downloads, no prints."""
    model_id = 2
    home = osp.join(osp.expanduser("~"), ".romp")
    p = argparse.ArgumentParser(description="BEV (B200-native hot path)")
    p.add_argument("-m", "--mode", type=str, default="image")
    p.add_argument("--model_id", type=int, default=2)
    p.add_argument("-i", "--input", type=str, default=None)
    p.add_argument("-o", "--save_path", type=str, default=osp.join(osp.expanduser("~"), "BEV_results"))
    p.add_argument("--crowd", action="store_false")
    p.add_argument("--GPU", type=int, default=0)
    p.add_argument("--overlap_ratio", type=float, default=long_conf_dict[model_id][3])
    p.add_argument("--center_thresh", type=float, default=conf_dict[model_id][0])
    p.add_argument("--nms_thresh", type=float, default=conf_dict[model_id][1])
    p.add_argument("--relative_scale_thresh", type=float, default=conf_dict[model_id][2])
    p.add_argument("--show_largest", action="store_true")
    p.add_argument("--show_patch_results", action="store_true")
    p.add_argument("--calc_smpl", action="store_false")
    p.add_argument("--renderer", type=str, default="sim3dr")
    p.add_argument("--render_mesh", action="store_false")
    p.add_argument("--show", action="store_true")
    p.add_argument("--show_items", type=str, default="mesh,mesh_bird_view")
    p.add_argument("--save_video", action="store_true")
    p.add_argument("--frame_rate", type=int, default=24)
    p.add_argument("--smpl_path", type=str, default=osp.join(home, "SMPLA_NEUTRAL.pth"))
    p.add_argument("--smil_path", type=str, default=osp.join(home, "smil_packed_info.pth"))
    p.add_argument("--model_path", type=str, default=osp.join(home, model_dict[model_id]))
    p.add_argument("-t", "--temporal_optimize", action="store_true")
    p.add_argument("-sc", "--smooth_coeff", type=float, default=3.0)
    p.add_argument("--webcam_id", type=int, default=0)
    p.add_argument("--precision", type=str, default="bf16", choices=["bf16", "tf32", "fp32"])
    p.add_argument("--max_batch", type=int, default=32)
    args = p.parse_args(input_args)
    if args.model_id != 2:                                            # bev/main.py:59-63
        args.model_path = osp.join(home, model_dict[args.model_id])
        args.center_thresh, args.nms_thresh = conf_dict[args.model_id][0], conf_dict[args.model_id][1]
        args.relative_scale_thresh = conf_dict[model_id][2]
    if args.crowd:                                                    # :81-85 (crowd defaults to True)
        args.center_thresh, args.nms_thresh = long_conf_dict[args.model_id][0], long_conf_dict[args.model_id][1]
        args.relative_scale_thresh, args.overlap_ratio = long_conf_dict[model_id][2], long_conf_dict[args.model_id][3]
    return args


class BEV(torch.nn.Module):
    """``BEV(settings)(image_bgr)`` - the contract of simple_romp/bev/main.py:91-181 for normal images."""

    result_keys = ["smpl_thetas", "smpl_betas", "cam", "cam_trans", "params_pred", "center_confs", "pred_batch_ids"]   # :115

    def __init__(self, settings, state_dict=None, smpla_pack=None, smil_pack=None):
        super().__init__()
        self.settings = s = settings
        if not torch.cuda.is_available() or s.GPU < 0:
            raise RuntimeError("romp_b200.BEV needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        if getattr(s, "temporal_optimize", False) or getattr(s, "show", False) or getattr(s, "show_largest", False):
            raise NotImplementedError("temporal smoothing / display are outside the B200 hot path (SURVEY.md section 2)")
        # NB the reference renders by default (render_mesh is store_false, bev/main.py:44); rendering is out of scope and
        # simply not performed here.
        self.lib = _lib.load()
        self.device_index = int(s.GPU)
        self.tdevice = torch.device("cuda", self.device_index)
        torch.cuda.set_device(self.tdevice)
        self.precision = getattr(s, "precision", "bf16")
        self._staging = {}
        self.max_batch = B = int(getattr(s, "max_batch", 32))
        if state_dict is None:
            state_dict = torch.load(s.model_path, map_location="cpu")          # bev/main.py:101 (strict=False)
        self._sd = state_dict
        self._nets = {}
        self.stream = torch.cuda.Stream(device=self.tdevice)
        w = graph.bev_weights(state_dict)
        self._w_keep = w
        fp = C.POINTER(C.c_float)
        bw = _lib.BevWeights(*[w[k].ctypes.data_as(fp) for k in ("center_ref", "cam_ref", "coordmap", "anchors", "embed",
                                                                   "w0", "b0", "w1", "b1", "w2", "b2")])
        self.h = self.lib.b200romp_bev_create(self.device_index, C.byref(bw))
        if not self.h:
            raise RuntimeError("b200romp_bev_create: " + self.lib.b200romp_last_error().decode())
        self.calc_smpl = bool(s.calc_smpl)
        if self.calc_smpl:
            if smpla_pack is None:
                smpla_pack = torch.load(s.smpl_path, map_location="cpu")
            if smil_pack is None:
                smil_pack = torch.load(s.smil_path, map_location="cpu")
            self.smpla = SMPLParser(smpla_pack, self.device_index, n_betas=11, shape_key="smpla_shapedirs")   # post_parser.py:259
            self.smil = SMPLParser(smil_pack, self.device_index, n_betas=10)                                     # :258
        self._alloc(B)

    def _net(self, in_dtype):
        if in_dtype not in self._nets:
            self._nets[in_dtype] = graph.build_bev(self._sd, self.device_index, self.precision, in_dtype, self.max_batch)
        return self._nets[in_dtype]

    def _alloc(self, B):
        dev, cap = self.tdevice, B * MAX_PERSON
        self.cap = cap
        act = torch.bfloat16 if self.precision == "bf16" else torch.float32
        self.act_code = BF16 if self.precision == "bf16" else F32
        z = lambda *shape, dtype=torch.float32: torch.zeros(*shape, dtype=dtype, device=dev)
        i64, i32 = torch.int64, torch.int32
        self.buf = dict(
            maps_fv=z(B, 4, 128, 128), fv_feats=z(B, 128, 128, 128, dtype=act), img_feats=z(B, 128, 128, graph.bev_feats_channels(self.precision), dtype=act),
            bv_in=z(B, 1, 128, 2560, dtype=act), bv_out=z(B, 1, 128, 128, dtype=act),
            c3d_tmp=z(B, 64, 128, 128), center3d=z(B, 64, 128, 128),
            parse_ws=torch.zeros(int(self.lib.b200romp_bev_parse_workspace_bytes(B)), dtype=torch.uint8, device=dev),
            count=z(1, dtype=i32), batch_ids=z(cap, dtype=i64), czyx=z(cap, 3, dtype=i64), conf=z(cap),
            params_pred=z(cap, N_PARAMS), cam_czyx=z(cap, 3, dtype=i64), cam=z(cap, 3), thetas=z(cap, 72), betas=z(cap, 11),
            cam_trans=z(cap, 3), pj2d_org=z(cap, 71, 2), keep=z(cap, dtype=i32), sel=z(cap, dtype=i32), count2=z(1, dtype=i32),
        )
        if self.calc_smpl:
            self.buf.update(verts=z(cap, 6890, 3), joints=z(cap, 71, 3), verts_smil=z(cap, 6890, 3), joints_smil=z(cap, 71, 3),
                            smpl_ws=z(cap, self.smpla.ws_floats))
        # compacted outputs (after the per-frame post filters)
        self.out = {k: torch.zeros_like(self.buf[k]) for k in ("batch_ids", "conf", "params_pred", "cam", "thetas", "betas",
                                                               "cam_trans", "pj2d_org")}
        if self.calc_smpl:
            self.out.update(verts=torch.zeros_like(self.buf["verts"]), joints=torch.zeros_like(self.buf["joints"]))
        self.count_host = torch.zeros(2, dtype=torch.int32).pin_memory()

    # ------------------------------------------------------------------------------------------
    @torch.no_grad()
    def run_model(self, frames_dev, center3d_override=None):
        """BEVv1.forward (bev/model.py:232-250) + pack_params_dict / cam_trans (bev/main.py:128-129); no host sync."""
        B = frames_dev.shape[0]
        assert frames_dev.is_cuda and frames_dev.is_contiguous() and tuple(frames_dev.shape[1:]) == (512, 512, 3)
        assert B <= self.max_batch
        in_dtype = {torch.uint8: U8, torch.float32: F32}[frames_dev.dtype]
        g1, io1, g2, io2 = self._net(in_dtype)
        lib, b, sp, ac = self.lib, self.buf, C.c_void_p(self.stream.cuda_stream), self.act_code
        ck = _lib.check
        ck(lib.b200romp_net_bind(g1.net, io1["frames"], _ptr(frames_dev)))
        ck(lib.b200romp_net_bind(g1.net, io1["maps_fv"], _ptr(b["maps_fv"])))
        ck(lib.b200romp_net_bind(g1.net, io1["fv_feats"], _ptr(b["fv_feats"])))
        ck(lib.b200romp_net_bind(g1.net, io1["img_feats"], _ptr(b["img_feats"])))
        ck(lib.b200romp_net_run(g1.net, B, sp), "net_run(g1)")
        ck(lib.b200romp_bev_bv_input(_ptr(b["maps_fv"]), _ptr(b["img_feats"]), ac, b["img_feats"].shape[-1], B, _ptr(b["bv_in"]), ac, sp),
           "bv_input")
        ck(lib.b200romp_net_bind(g2.net, io2["bv_in"], _ptr(b["bv_in"])))
        ck(lib.b200romp_net_bind(g2.net, io2["bv_out"], _ptr(b["bv_out"])))
        ck(lib.b200romp_net_run(g2.net, B, sp), "net_run(g2)")
        ck(lib.b200romp_bev_center3d(self.h, _ptr(b["maps_fv"]), _ptr(b["bv_out"]), ac, B, _ptr(b["c3d_tmp"]), _ptr(b["center3d"]), sp),
           "center3d")
        c3d = b["center3d"] if center3d_override is None else center3d_override
        ck(lib.b200romp_bev_parse3d(_ptr(c3d), B, float(self.settings.center_thresh), self.cap, _ptr(b["count"]), _ptr(b["batch_ids"]),
                                    _ptr(b["czyx"]), _ptr(b["conf"]), _ptr(b["parse_ws"]), sp), "parse3d")
        ck(lib.b200romp_bev_regress(self.h, _ptr(b["maps_fv"]), _ptr(b["bv_out"]), ac, _ptr(b["fv_feats"]), ac, B * MAX_PERSON,
                                    _ptr(b["count"]), _ptr(b["batch_ids"]), _ptr(b["czyx"]), _ptr(b["params_pred"]),
                                    _ptr(b["cam_czyx"]), _ptr(b["cam"]), _ptr(b["thetas"]), _ptr(b["betas"]), _ptr(b["cam_trans"]), sp),
           "regress")

    @torch.no_grad()
    def run_post(self, B, offsets, img_max_side=512.0):
        """SMPLA_parser + projection + the two per-frame filters (bev/main.py:172-180), then row compaction."""
        lib, b, o, sp = self.lib, self.buf, self.out, C.c_void_p(self.stream.cuda_stream)
        cap = B * MAX_PERSON
        off = (C.c_float * 6)(*[float(v) for v in offsets])
        if not self.calc_smpl:
            return
        st = self.stream.cuda_stream
        self.smpla.forward(b["betas"], b["thetas"], cap, b["count"], True, b["smpl_ws"], b["verts"], b["joints"], st)
        self.smil.forward(b["betas"], b["thetas"], cap, b["count"], True, b["smpl_ws"], b["verts_smil"], b["joints_smil"], st)
        _lib.check(lib.b200romp_bev_post(_ptr(b["betas"]), _ptr(b["verts_smil"]), _ptr(b["joints_smil"]), _ptr(b["verts"]),
                                         _ptr(b["joints"]), _ptr(b["cam"]), _ptr(b["cam_trans"]), _ptr(b["batch_ids"]), B, cap,
                                         _ptr(b["count"]), off, float(self.settings.nms_thresh),
                                         float(self.settings.relative_scale_thresh), float(img_max_side), _ptr(b["pj2d_org"]),
                                         _ptr(b["keep"]), _ptr(b["sel"]), _ptr(b["count2"]), sp), "bev_post")
        for k, dst in o.items():
            src = b["conf"] if k == "conf" else b[k]
            row = src[0].numel() * src.element_size()
            _lib.check(lib.b200romp_gather_rows(_ptr(src), row, _ptr(b["sel"]), _ptr(b["count2"]), cap, _ptr(dst), sp), "gather_rows")

    def collect(self, to_numpy=True):
        with torch.cuda.stream(self.stream):
            self.count_host[0:1].copy_(self.buf["count"], non_blocking=True)
            self.count_host[1:2].copy_(self.buf["count2"], non_blocking=True)
        self.stream.synchronize()
        n_det, n = int(self.count_host[0]), int(self.count_host[1])
        if n_det == 0:
            return None
        if not self.calc_smpl:
            b = self.buf
            out = {"smpl_thetas": b["thetas"][:n_det], "smpl_betas": b["betas"][:n_det], "cam": b["cam"][:n_det],
                   "cam_trans": b["cam_trans"][:n_det], "params_pred": b["params_pred"][:n_det],
                   "center_confs": b["conf"][:n_det], "pred_batch_ids": b["batch_ids"][:n_det]}
        else:
            o = self.out
            out = {"smpl_thetas": o["thetas"][:n], "smpl_betas": o["betas"][:n], "cam": o["cam"][:n], "cam_trans": o["cam_trans"][:n],
                   "params_pred": o["params_pred"][:n], "center_confs": o["conf"][:n], "pred_batch_ids": o["batch_ids"][:n],
                   "verts": o["verts"][:n], "joints": o["joints"][:n], "pj2d_org": o["pj2d_org"][:n]}
        if to_numpy:
            with torch.cuda.stream(self.stream):
                out = {k: v.contiguous().cpu().numpy() for k, v in out.items()}
        return out

    @torch.no_grad()
    def forward_batch(self, frames, offsets=None, to_numpy=True, center3d_override=None, img_max_side=512.0):
        if isinstance(frames, np.ndarray):
            frames = torch.from_numpy(frames)
        B = frames.shape[0]
        # device-resident inputs come from the caller's current stream: order self.stream after it (no host sync)
        cur = torch.cuda.current_stream(self.tdevice)
        for t in (frames, center3d_override):
            if isinstance(t, torch.Tensor) and t.is_cuda:
                if cur != self.stream:
                    self.stream.wait_stream(cur)
                t.record_stream(self.stream)
        key = (frames.dtype, B)
        if key not in self._staging:                  # stable pointer -> one cached CUDA graph per (dtype, B)
            self._staging[key] = torch.empty((B, 512, 512, 3), dtype=frames.dtype, device=self.tdevice)
        fd = self._staging[key]
        with torch.cuda.stream(self.stream):
            fd.copy_(frames, non_blocking=True)
            self.run_model(fd, center3d_override)
            self.run_post(B, offsets if offsets is not None else [0, 512, 0, 512, 512, 512], img_max_side)
        return self.collect(to_numpy)

    @torch.no_grad()
    def forward(self, image, signal_ID=0, **kwargs):
        """image: HxWx3 uint8 BGR.  bev/main.py:139-156 (normal images; the >=2:1 panoramic tiling is out of scope)."""
        if image.shape[1] / image.shape[0] >= 2 and self.settings.crowd:
            raise NotImplementedError("long-image sliding-window processing (bev/split2process.py) is out of scope")
        inp, pad = img_preprocess(image)
        out = self.forward_batch(torch.from_numpy(inp), offsets=pad, img_max_side=float(max(image.shape[:2])))
        if out is None:
            print("No person detected!")                                       # bev/model.py:239
            return None
        return out

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.lib.b200romp_bev_destroy(self.h)
        except Exception:
            pass
