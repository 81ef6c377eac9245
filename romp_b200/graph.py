"""Host-side graph builder for seam S1: turns the reference's state dicts (ROMPv1 / BEVv1) into the fused conv ops
executed by libb200romp (b200romp_net_*).

It mirrors the *structure* of simple_romp/romp/model.py (HigherResolutionNet :246-417, HighResolutionModule
:129-244, Bottleneck :85-123, BasicBlock :54-83, ROMPv1 head :420-481) and of simple_romp/bev/model.py
(BEVv1 heads :142-186) and consumes exactly the reference's state-dict keys, so released checkpoints load
unchanged.  Work done here is weight preprocessing only (BatchNorm folding, constant folding of the coord-map
channels); every per-frame FLOP runs in the CUDA library.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch
import torch.nn.functional as F

from . import _lib
from ._lib import BF16, F32, U8, ConvDesc, SumDesc

BN_EPS = 1e-5


def fold_bn(sd, conv, bn):
    """Conv(+bias) followed by eval BatchNorm -> (weight, bias), folded in float64 (model.py:49-52,70-72)."""
    w = np.asarray(sd[conv + ".weight"], dtype=np.float64)
    b = np.asarray(sd[conv + ".bias"], dtype=np.float64) if (conv + ".bias") in sd else np.zeros(w.shape[0])
    if bn is None:
        return w.astype(np.float32), b.astype(np.float32)
    g = np.asarray(sd[bn + ".weight"], np.float64)
    beta = np.asarray(sd[bn + ".bias"], np.float64)
    mean = np.asarray(sd[bn + ".running_mean"], np.float64)
    var = np.asarray(sd[bn + ".running_var"], np.float64)
    scale = g / np.sqrt(var + BN_EPS)
    return (w * scale.reshape((-1,) + (1,) * (w.ndim - 1))).astype(np.float32), ((b - mean) * scale + beta).astype(np.float32)


def round_bf16(a):
    return torch.from_numpy(np.ascontiguousarray(a)).bfloat16().float().numpy()


def to_numpy_sd(sd):
    return {k: (v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)) for k, v in sd.items()}


class NetBuilder:
    """Thin typed wrapper over b200romp_net_add_tensor / add_conv that tracks tensor shapes."""

    def __init__(self, device: int, precision: str, engine: int = _lib.ENGINE_AUTO):
        """precision: "bf16" = bf16 tensors, tcgen05 kind::f16 (fast); "tf32" = fp32 tensors, tcgen05 kind::tf32 with
        round-to-nearest operand conversion = the arithmetic of the reference's default GPU path (cudnn.allow_tf32);
        "fp32" = fp32 tensors, SIMT fp32 FMA (strict parity engine)."""
        assert precision in ("fp32", "bf16", "tf32")
        self.lib = _lib.load()
        self.precision = precision
        self.act = BF16 if precision == "bf16" else F32
        if precision == "tf32" and engine == _lib.ENGINE_AUTO:
            engine = _lib.ENGINE_TF32
        self.engine = engine
        self.net = self.lib.b200romp_net_create(device)
        if not self.net:
            raise RuntimeError("b200romp_net_create: " + self.lib.b200romp_last_error().decode())
        self.shape = {}
        self.names = {}
        self.flops_per_frame = 0
        self.lane = 0

    def on_lane(self, lane):
        """with nb.on_lane(b): ...  - ops added inside run on concurrency lane b of the captured CUDA graph."""
        nb = self

        class _Lane:
            def __enter__(self_):
                self_.prev, nb.lane = nb.lane, int(lane) % 4

            def __exit__(self_, *exc):
                nb.lane = self_.prev
        return _Lane()

    def _added(self, op_id):
        if self.lane:
            _lib.check(self.lib.b200romp_net_set_lane(self.net, op_id, self.lane), "set_lane")

    def tensor(self, H, W, Cc, dtype=None, nchw=0, external=0, name=None):
        dtype = self.act if dtype is None else dtype
        t = _lib.check(self.lib.b200romp_net_add_tensor(self.net, H, W, Cc, dtype, nchw, external), "add_tensor")
        self.shape[t] = (H, W, Cc, dtype)
        if name:
            self.names[name] = t
        return t

    def const_tensor(self, hwc: np.ndarray, name=None):
        a = np.ascontiguousarray(hwc, dtype=np.float32)
        t = _lib.check(self.lib.b200romp_net_add_const_tensor(self.net, a.shape[0], a.shape[1], a.shape[2], F32,
                                                              a.ctypes.data_as(C.c_void_p)), "add_const_tensor")
        self.shape[t] = (a.shape[0], a.shape[1], a.shape[2], F32)
        if name:
            self.names[name] = t
        return t

    def conv(self, x, w, b, *, stride=1, relu=False, res=None, res_c_off=0, res_broadcast=0, up=1, out=None,
             out_c_off=0, out_dtype=None, in_c_off=0, input_norm=0, pow_channel=-1, engine=None, name=None):
        """w: OIHW (Conv2d) or [O,I,3] (Conv1d along W, ksize code 13)."""
        conv1d = w.ndim == 3
        cout, cin = w.shape[0], w.shape[1]
        k = 13 if conv1d else w.shape[2]
        kh, kw = (1, 3) if conv1d else (k, k)
        H, W, _, _ = self.shape[x]
        Ho, Wo = ((H + 2 * (kh // 2) - kh) // stride + 1) * up, ((W + 2 * (kw // 2) - kw) // stride + 1) * up
        if out is None:
            out = self.tensor(Ho, Wo, cout, out_dtype, name=name)
        d = ConvDesc(x, in_c_off, out, out_c_off, -1 if res is None else res, res_c_off, res_broadcast, cin, cout, k,
                     stride, int(relu), up, input_norm, pow_channel, self.engine if engine is None else engine)
        w = np.ascontiguousarray(w, dtype=np.float32)
        if self.precision == "bf16":
            w = round_bf16(w)      # both engines then see identical bf16-representable weights
        bp = None if b is None else np.ascontiguousarray(b, dtype=np.float32)
        self.flops_per_frame += 2 * cout * cin * kh * kw * (Ho // up) * (Wo // up)
        ksplit = self.precision == "tf32" or (self.precision == "bf16" and os.environ.get("B200ROMP_NO_S2_KSPLIT") != "1")
        if ksplit and k == 3 and stride == 2 and cin == 256 and res is None and up == 1:
            # the weights of a 256-channel 3x3 slab (fp32: 295 KB; bf16: 147 KB at N = 32) leave no room for a pipeline in
            # shared memory: split K into 128-channel convs, each accumulating onto the previous one (fp32
            # intermediates, no rounding in between)
            pc = 128 if self.precision == "tf32" else int(os.environ.get("B200ROMP_S2_KSPLIT_C", "128"))   # channels per part
            n_parts, prev = cin // pc, -1
            for i in range(n_parts):
                last = i == n_parts - 1
                dst = out if last else self.tensor(Ho, Wo, cout, F32)
                dd = ConvDesc(x, in_c_off + i * pc, dst, out_c_off if last else 0, prev, 0, 0, pc, cout, k, stride,
                              int(relu) if last else 0, 1, input_norm, pow_channel if last else -1, d.engine)
                ws = np.ascontiguousarray(w[:, i * pc:(i + 1) * pc])
                bb = bp if i == 0 else None
                self._added(_lib.check(self.lib.b200romp_net_add_conv(
                    self.net, C.byref(dd), ws.ctypes.data_as(C.POINTER(C.c_float)),
                    None if bb is None else bb.ctypes.data_as(C.POINTER(C.c_float))), "add_conv"))
                prev = dst
            return out
        self._added(_lib.check(self.lib.b200romp_net_add_conv(
            self.net, C.byref(d), w.ctypes.data_as(C.POINTER(C.c_float)),
            None if bp is None else bp.ctypes.data_as(C.POINTER(C.c_float))), "add_conv"))
        return out

    def sum(self, base, terms, ups, relu=True, out_dtype=None, name=None, c_offs=None):
        """out = act(base + sum_k nearest_up(terms[k][..., c_offs[k]:c_offs[k]+C], ups[k])) - the HRNet fuse-layer summation
        (model.py:226-244); a term may be a channel slice of a wider tensor (merged 1x1 convs)."""
        H, W, Cc, _ = self.shape[base]
        out = self.tensor(H, W, Cc, out_dtype, name=name)
        c_offs = [0] * len(terms) if c_offs is None else list(c_offs)
        d = SumDesc(out, base, len(terms), (C.c_int * 4)(*(list(terms) + [0] * (4 - len(terms)))),
                    (C.c_int * 4)(*(list(ups) + [1] * (4 - len(ups)))), int(relu),
                    (C.c_int * 4)(*(c_offs + [0] * (4 - len(c_offs)))))
        self._added(_lib.check(self.lib.b200romp_net_add_sum(self.net, C.byref(d)), "add_sum"))
        return out

    def maxpool(self, x, name=None):
        """MaxPool2d(3, 2, 1) (ResNet-50 stem, romp/lib/models/resnet_50.py:42)."""
        H, W, Cc, dt = self.shape[x]
        out = self.tensor((H - 1) // 2 + 1, (W - 1) // 2 + 1, Cc, dt, name=name)
        self._added(_lib.check(self.lib.b200romp_net_add_maxpool(self.net, x, out), "add_maxpool"))
        return out

    def deconv(self, x, w, b, relu=True, name=None):
        """ConvTranspose2d(4, 2, 1) + folded BN (+ReLU); w: PyTorch layout [cin, cout, 4, 4] (resnet_50.py:93-120)."""
        cin, cout = w.shape[0], w.shape[1]
        H, W, _, _ = self.shape[x]
        out = self.tensor(2 * H, 2 * W, cout, None, name=name)
        d = ConvDesc(x, 0, out, 0, -1, 0, 0, cin, cout, 42, 2, int(relu), 1, 0, -1, _lib.ENGINE_SIMT)
        w = np.ascontiguousarray(w, dtype=np.float32)
        bp = np.ascontiguousarray(b, dtype=np.float32)
        self._added(_lib.check(self.lib.b200romp_net_add_conv(self.net, C.byref(d), w.ctypes.data_as(C.POINTER(C.c_float)),
                                                              bp.ctypes.data_as(C.POINTER(C.c_float))), "add_conv(deconv)"))
        self.flops_per_frame += 2 * cin * cout * 4 * (2 * H) * (2 * W)
        return out

    def finalize(self, max_batch):
        _lib.check(self.lib.b200romp_net_finalize(self.net, max_batch), "net_finalize")

    def describe(self):
        buf = C.create_string_buffer(1 << 18)
        self.lib.b200romp_net_describe(self.net, buf, len(buf))
        return buf.value.decode()

    def profile(self, batch, iters=5, stream=None):
        """Mean device time (us) of every op of one run, ops launched one by one (b200romp_net_profile)."""
        n = self.lib.b200romp_net_num_launches(self.net)
        us = (C.c_float * n)()
        st = torch.cuda.current_stream().cuda_stream if stream is None else stream
        _lib.check(self.lib.b200romp_net_profile(self.net, batch, iters, us, C.c_void_p(st)), "net_profile")
        return list(us)


def coord_maps(size=128):
    """get_coord_maps, model.py:8-37 -> [1,2,size,size]; ch0 varies along W, ch1 along H."""
    r = torch.arange(size, dtype=torch.float32) / (size - 1) * 2 - 1
    return torch.stack([r.view(1, size).expand(size, size), r.view(size, 1).expand(size, size)])[None].contiguous()


def build_backbone(nb: NetBuilder, sd, in_dtype):
    """HigherResolutionNet (model.py:336-417) -> (frames tensor id, [B,128,128,32] feature tensor id)."""
    act = nb.act

    def cb(x, conv, bn, **kw):
        w, b = fold_bn(sd, conv, bn)
        return nb.conv(x, w, b, **kw)

    frames = nb.tensor(512, 512, 3, in_dtype, external=1, name="frames")
    p = "backbone."
    x = cb(frames, p + "conv1", p + "bn1", stride=2, relu=True, input_norm=1)   # model.py:385-387
    # layer1.0 (Bottleneck with a 1x1 downsample on the skip, model.py:93-120,391): relu(bn3(conv3(t)) + bnd(convd(x))) is ONE
    # 1x1 conv over the channel concatenation [t | x] with weights [W3 | Wd] and bias b3 + bd.  conv2 of the stem and
    # conv2 of the block write the two halves of that 128-channel tensor, so the 256-channel downsample output (0.5 GB
    # per 64 frames, written once and read once as a residual) never exists.
    q = p + "layer1.0."
    fuse_skip = (q + "downsample.0.weight") in sd and os.environ.get("B200ROMP_NO_SKIP_CONCAT") != "1"
    if fuse_skip:
        cat = nb.tensor(nb.shape[x][0] // 2, nb.shape[x][1] // 2, 128)
        cb(x, p + "conv2", p + "bn2", stride=2, relu=True, out=cat, out_c_off=64)                          # :388-390
        y = cb(cat, q + "conv1", q + "bn1", relu=True, in_c_off=64)
        cb(y, q + "conv2", q + "bn2", relu=True, out=cat, out_c_off=0)
        (w3, b3), (wd, bd) = fold_bn(sd, q + "conv3", q + "bn3"), fold_bn(sd, q + "downsample.0", q + "downsample.1")
        x = nb.conv(cat, np.concatenate([w3, wd], axis=1), b3 + bd, relu=True)
    else:
        x = cb(x, p + "conv2", p + "bn2", stride=2, relu=True)
    for i in range(1 if fuse_skip else 0, 4):                                                              # layer1, :391
        q = f"{p}layer1.{i}."
        y = cb(x, q + "conv1", q + "bn1", relu=True)
        y = cb(y, q + "conv2", q + "bn2", relu=True)
        res = cb(x, q + "downsample.0", q + "downsample.1") if (q + "downsample.0.weight") in sd else x
        x = cb(y, q + "conv3", q + "bn3", relu=True, res=res)

    def basic_block(x, q):
        t = cb(x, q + "conv1", q + "bn1", relu=True)
        return cb(t, q + "conv2", q + "bn2", relu=True, res=x)

    def hr_module(xs, q, nbr, multi=True):
        """HighResolutionModule.forward model.py:226-244.  Every fuse term is produced at its own resolution (1x1 conv
        of a lower-resolution branch / chain of stride-2 3x3 convs of a higher-resolution one) and ONE sum op per
        output branch adds them in fp32 to the identity term, upsampling on the fly, then ReLU (model.py:243)."""
        xs = list(xs)
        # The branches are independent (model.py:226-233): branch b is tagged with concurrency lane b.  Lanes are OFF by
        # default (b200romp_net: B200ROMP_LANES=1 turns them on): measured on B200, capturing the branches on separate
        # streams LOSES 5 % (6044 vs 6373 frames/s, profiles/r02_bench_b_*lanes.json) - the persistent conv kernels each
        # fill every SM, so concurrency only interleaves their CTAs, which thrashes L2 (four streamed working sets
        # instead of one producer->consumer pair) and drops the programmatic-dependent-launch edges at lane crossings.
        # Ops stay in branch-major order: a conv's output is consumed while still L2-resident.
        for b in range(nbr):
            with nb.on_lane(b):
                for k in range(4):
                    xs[b] = basic_block(xs[b], f"{q}branches.{b}.{k}.")
        outs = []
        n_out = nbr if multi else 1
        # The 1x1 fuse convs that read the same branch j (one per output i < j, model.py:188-197) are ONE conv with the
        # output channels concatenated (zero-padded to a multiple of 64): the sums read channel slices of its output.
        merged = {}
        if os.environ.get("B200ROMP_NO_FUSE1X1_MERGE") != "1":
            for j in range(1, nbr):
                tgt = [i for i in range(n_out) if i < j]
                if len(tgt) < 2:
                    continue
                folded = [fold_bn(sd, f"{q}fuse_layers.{i}.{j}.0", f"{q}fuse_layers.{i}.{j}.1") for i in tgt]
                ctot = sum(w.shape[0] for w, _ in folded)
                pad = (-ctot) % 64
                wcat = np.concatenate([w for w, _ in folded] + ([np.zeros((pad,) + folded[0][0].shape[1:], np.float32)] if pad else []), 0)
                bcat = np.concatenate([b for _, b in folded] + ([np.zeros(pad, np.float32)] if pad else []), 0)
                with nb.on_lane(j):
                    t = nb.conv(xs[j], wcat, bcat)
                off = 0
                for i, (w, _) in zip(tgt, folded):
                    merged[(i, j)] = (t, off)
                    off += w.shape[0]
        for i in range(n_out):
            terms, ups, offs = [], [], []
            for j in range(nbr):
                if j == i:
                    continue                  # identity term (model.py:236-239) is the base of the sum
                r = f"{q}fuse_layers.{i}.{j}."
                with nb.on_lane(j):           # a fuse term is computed from branch j alone
                    if (i, j) in merged:
                        terms.append(merged[(i, j)][0])
                        offs.append(merged[(i, j)][1])
                        ups.append(2 ** (j - i))
                    elif j > i:               # 1x1 conv + BN, nearest upsample folded into the sum (model.py:188-197)
                        terms.append(cb(xs[j], r + "0", r + "1"))
                        offs.append(0)
                        ups.append(2 ** (j - i))
                    else:                     # chain of stride-2 3x3 convs (model.py:200-218)
                        t = xs[j]
                        for k in range(i - j - 1):
                            t = cb(t, f"{r}{k}.0", f"{r}{k}.1", stride=2, relu=True)
                        k = i - j - 1
                        terms.append(cb(t, f"{r}{k}.0", f"{r}{k}.1", stride=2))
                        offs.append(0)
                        ups.append(1)
            with nb.on_lane(i):
                acc = nb.sum(xs[i], terms, ups, relu=True, c_offs=offs)
            outs.append(acc)
        return outs

    xs = [cb(x, p + "transition1.0.0", p + "transition1.0.1", relu=True),                     # model.py:393-398
          cb(x, p + "transition1.1.0.0", p + "transition1.1.0.1", stride=2, relu=True)]
    ys = hr_module(xs, p + "stage2.0.", 2)
    xs = ys + [cb(ys[-1], p + "transition2.2.0.0", p + "transition2.2.0.1", stride=2, relu=True)]   # :401-406
    for m in range(4):
        xs = hr_module(xs, f"{p}stage3.{m}.", 3)
    xs = xs + [cb(xs[-1], p + "transition3.3.0.0", p + "transition3.3.0.1", stride=2, relu=True)]    # :409-414
    for m in range(3):
        xs = hr_module(xs, f"{p}stage4.{m}.", 4, multi=(m != 2))
    nb.names["backbone_out"] = xs[0]
    return frames, xs[0]


def build_romp_head(nb: NetBuilder, sd, feat, feat_c):
    """ROMPv1 heads (model.py:445-481) on a [B,128,128,feat_c] feature tensor -> (center_maps, params_maps) external tensor ids.
    The three head-in convs (feat_c+2)->64 (3x3, s2, bias, BN, ReLU) are fused into one feat_c->192 conv; the two constant
    coord channels (model.py:473) become a per-pixel bias map."""
    def cb(x, conv, bn, **kw):
        w, b = fold_bn(sd, conv, bn)
        return nb.conv(x, w, b, **kw)

    def basic_block(x, q, in_c_off=0, res_c_off=0):
        t = cb(x, q + "conv1", q + "bn1", relu=True, in_c_off=in_c_off)
        return cb(t, q + "conv2", q + "bn2", relu=True, res=x, res_c_off=res_c_off)

    order = (3, 1, 2)                 # cam, params, center  -> channel slices 0, 64, 128 of the fused tensor
    ws, bs = zip(*[fold_bn(sd, f"final_layers.{h}.0.0", f"final_layers.{h}.0.1") for h in order])
    w_all, b_all = np.concatenate(ws, 0), np.concatenate(bs, 0)
    with torch.no_grad():
        cm = F.conv2d(coord_maps(128), torch.from_numpy(w_all[:, feat_c:feat_c + 2].copy()), None, stride=2, padding=1)[0]
    bias_map = (cm + torch.from_numpy(b_all)[:, None, None]).permute(1, 2, 0).contiguous().numpy()   # [64,64,192]
    bias_t = nb.const_tensor(bias_map, name="head_bias_map")
    hin = nb.conv(feat, np.ascontiguousarray(w_all[:, :feat_c]), None, stride=2, relu=True, res=bias_t, res_broadcast=1,
                  name="head_in")
    center_maps = nb.tensor(64, 64, 1, F32, nchw=1, external=1, name="center_maps")
    params_maps = nb.tensor(64, 64, 145, F32, nchw=1, external=1, name="params_maps")
    for s, h in enumerate(order):                          # the three heads are independent: one (optional) lane each
        q = f"final_layers.{h}."
        with nb.on_lane(s):
            y = basic_block(hin, q + "1.0.0.", in_c_off=64 * s, res_c_off=64 * s)
            y = basic_block(y, q + "1.1.0.")
            w, b = fold_bn(sd, q + "2", None)
            if h == 3:      # cam maps -> params_maps[:, 0:3], cam scale 1.1**x (model.py:480, main.py:113)
                nb.conv(y, w, b, out=params_maps, out_c_off=0, pow_channel=0)
            elif h == 1:    # params maps -> params_maps[:, 3:145]
                nb.conv(y, w, b, out=params_maps, out_c_off=3)
            else:
                nb.conv(y, w, b, out=center_maps)
    return center_maps, params_maps


def build_romp(sd, device=0, precision="bf16", in_dtype=U8, max_batch=64, engine=_lib.ENGINE_AUTO):
    """ROMPv1 (HRNet-32 + 3 heads) as a libb200romp conv graph.

    Returns (builder, io) with io = dict(frames=, center_maps=, params_maps=) external tensor ids.
    """
    sd = to_numpy_sd(sd)
    nb = NetBuilder(device, precision, engine)
    frames, feat = build_backbone(nb, sd, in_dtype)
    center_maps, params_maps = build_romp_head(nb, sd, feat, 32)
    nb.finalize(max_batch)
    return nb, dict(frames=frames, center_maps=center_maps, params_maps=params_maps)


IMAGENET_MEAN = np.array([0.485, 0.456, 0.406], np.float32)
IMAGENET_STD = np.array([0.229, 0.224, 0.225], np.float32)


def build_romp_resnet50(sd, device=0, precision="fp32", in_dtype=F32, max_batch=1, engine=_lib.ENGINE_AUTO):
    """ROMP with the ResNet-50 backbone (romp/lib/models/resnet_50.py:19-120; BASELINE.json configs[0]) as a libb200romp
    conv graph: 7x7 s2 stem + MaxPool(3,2,1) + [3,4,6,3] Bottlenecks (stride on the 3x3 conv) + three
    ConvTranspose2d(4,2,1)+BN+ReLU 2048->256->128->64, then ROMPv1's heads on the 64+2 channels.
    The input normalisation (x/255 - mean)/std (:32-38) is folded into the stem: weights / (255*std) and - because the
    reference zero-pads the NORMALISED image - a per-pixel bias map = conv(constant image -mean/std, zero padded)."""
    sd = to_numpy_sd(sd)
    nb = NetBuilder(device, precision, engine)

    def cb(x, conv, bn, **kw):
        w, b = fold_bn(sd, conv, bn)
        return nb.conv(x, w, b, **kw)

    frames = nb.tensor(512, 512, 3, in_dtype, external=1, name="frames")
    p = "backbone."
    w, b = fold_bn(sd, p + "conv1", p + "bn1")                                             # [64,3,7,7], BN folded
    with torch.no_grad():
        const_img = torch.from_numpy(-IMAGENET_MEAN / IMAGENET_STD).view(1, 3, 1, 1).expand(1, 3, 512, 512).contiguous()
        bm = F.conv2d(const_img, torch.from_numpy(w), None, stride=2, padding=3)[0] + torch.from_numpy(b)[:, None, None]
    bias_t = nb.const_tensor(bm.permute(1, 2, 0).contiguous().numpy(), name="stem_bias_map")   # [256,256,64]
    w_eff = (w / (255.0 * IMAGENET_STD)[None, :, None, None]).astype(np.float32)
    x = nb.conv(frames, w_eff, None, stride=2, relu=True, res=bias_t, res_broadcast=1)     # resnet_50.py:40-41,57-59
    x = nb.maxpool(x)                                                                       # :42,60
    for li, blocks in enumerate((3, 4, 6, 3), start=1):                                     # :43-46, Bottleneck
        for bi in range(blocks):
            q = f"{p}layer{li}.{bi}."
            stride = 2 if (li > 1 and bi == 0) else 1
            y = cb(x, q + "conv1", q + "bn1", relu=True)
            y = cb(y, q + "conv2", q + "bn2", stride=stride, relu=True)
            res = cb(x, q + "downsample.0", q + "downsample.1", stride=stride) if (q + "downsample.0.weight") in sd else x
            x = cb(y, q + "conv3", q + "bn3", relu=True, res=res)
    for i in range(3):                                                                      # deconv_layers, :93-120
        wd = np.asarray(sd[f"{p}deconv_layers.{3 * i}.weight"], np.float64)                # [cin, cout, 4, 4]
        bn = f"{p}deconv_layers.{3 * i + 1}"
        scale = np.asarray(sd[bn + ".weight"], np.float64) / np.sqrt(np.asarray(sd[bn + ".running_var"], np.float64) + BN_EPS)
        shift = np.asarray(sd[bn + ".bias"], np.float64) - np.asarray(sd[bn + ".running_mean"], np.float64) * scale
        x = nb.deconv(x, (wd * scale[None, :, None, None]).astype(np.float32), shift.astype(np.float32), relu=True)
    nb.names["backbone_out"] = x
    center_maps, params_maps = build_romp_head(nb, sd, x, 64)
    nb.finalize(max_batch)
    return nb, dict(frames=frames, center_maps=center_maps, params_maps=params_maps)


def bev_feats_channels(precision):
    """channels of the img_feats tensor between bv_pre_layers and b200romp_bev_bv_input: 16 (bev/model.py:166-175), stored
    zero-padded to 32 in bf16 mode so that the stack runs on the tcgen05 engine"""
    return 32 if precision == "bf16" else 16


def build_bev(sd, device=0, precision="bf16", in_dtype=U8, max_batch=32, engine=_lib.ENGINE_AUTO):
    """BEVv1's convolutional part as two conv graphs (bev/model.py:142-186,232-250).

    G1: backbone -> det_head (maps_fv [B,4,128,128] fp32 NCHW), param_head (front-view features [B,128,128,128]),
        bv_pre_layers (img_feats [B,128,128,16]).
    G2: bv_out_layers = 3 x BasicBlock_1D (Conv1d k3 as ksize code 13) on the [B,1,128,2560] bird's-eye input that
        b200romp_bev_bv_input assembles from G1's outputs -> [B,1,128,128].
    Returns (g1, io1, g2, io2).
    """
    sd = to_numpy_sd(sd)
    g1 = NetBuilder(device, precision, engine)
    frames, feat = build_backbone(g1, sd, in_dtype)

    def head_block(x, q, out=None):
        """BasicBlock(32->128) whose residual is a biased 1x1 conv without BN (bev/model.py:154-156)."""
        w, b = fold_bn(sd, q + "conv1", q + "bn1")
        t = g1.conv(x, w, b, relu=True)
        wd, bd = fold_bn(sd, q + "downsample", None)
        res = g1.conv(x, wd, bd)
        w, b = fold_bn(sd, q + "conv2", q + "bn2")
        return g1.conv(t, w, b, relu=True, res=res, out=out)

    maps_fv = g1.tensor(128, 128, 4, F32, nchw=1, external=1, name="maps_fv")
    y = head_block(feat, "det_head.0.0.")
    w, b = fold_bn(sd, "det_head.1", None)
    g1.conv(y, w, b, out=maps_fv)                                                     # center_fv | cam_offset(3)
    fv = g1.tensor(128, 128, 128, None, external=1, name="fv_feats")
    head_block(feat, "param_head.0.0.", out=fv)
    # bv_pre_layers (:166-175): 1x1 32->16, 3x3 16->16, 1x1 16->16.  16 channels do not tile onto the tensor-core engine, so in
    # bf16 mode the stack runs zero-padded to 32 channels (zero weight rows / columns and zero bias: the padded channels are
    # relu(0) = 0 and contribute nothing downstream) - exact, and 3 tcgen05 launches instead of 0.84 ms of SIMT per step.
    cf = bev_feats_channels(precision)

    def padded(i, cin_to):
        w, b = fold_bn(sd, f"bv_pre_layers.{i}", f"bv_pre_layers.{i + 1}")
        if cf == 16:
            return w, b
        wp = np.zeros((cf, cin_to) + w.shape[2:], np.float32)
        wp[:w.shape[0], :w.shape[1]] = w
        bp = np.zeros(cf, np.float32)
        bp[:b.shape[0]] = b
        return wp, bp

    eng = _lib.ENGINE_SIMT if cf == 16 else None
    x = feat
    w, b = padded(0, 32)
    x = g1.conv(x, w, b, relu=True, engine=eng)
    w, b = padded(3, cf)
    x = g1.conv(x, w, b, relu=True, engine=eng)
    img_feats = g1.tensor(128, 128, cf, None, external=1, name="img_feats")
    w, b = padded(6, cf)
    g1.conv(x, w, b, relu=True, out=img_feats, engine=eng)
    g1.finalize(max_batch)

    g2 = NetBuilder(device, precision, _lib.ENGINE_AUTO if precision == "bf16" else _lib.ENGINE_SIMT)   # bf16: tcgen05 Conv1d engine (conv1d_tc.cu)
    bv_in = g2.tensor(1, 128, 2560, None, external=1, name="bv_in")
    y = bv_in
    bv_out = g2.tensor(1, 128, 128, None, external=1, name="bv_out")
    for i in range(3):                                                                 # BasicBlock_1D x3, :179-182,24-45
        q = f"bv_out_layers.{i}."
        w, b = fold_bn(sd, q + "conv1", q + "bn1")
        y = g2.conv(y, w, b, relu=True)
        w, b = fold_bn(sd, q + "conv2", q + "bn2")
        y = g2.conv(y, w, b, relu=True, out=bv_out if i == 2 else None)
    g2.finalize(max_batch)
    return (g1, dict(frames=frames, maps_fv=maps_fv, fv_feats=fv, img_feats=img_feats),
            g2, dict(bv_in=bv_in, bv_out=bv_out))


def bev_weights(sd):
    """Host arrays for b200romp_bev_create: BatchNorm3d-folded refiners, coord map, anchors, embedding, MLP."""
    from .synth import bev_cam3dmap_anchor
    sd = to_numpy_sd(sd)

    def ref(name, c):
        w1, b1 = fold_bn(sd, f"{name}.0.conv1", f"{name}.0.bn1")
        w2, b2 = fold_bn(sd, f"{name}.0.conv2", f"{name}.0.bn2")
        return np.concatenate([w1.reshape(-1), b1.reshape(-1), w2.reshape(-1), b2.reshape(-1)]).astype(np.float32)

    out = {"center_ref": ref("center_map_refiner", 1), "cam_ref": ref("cam_map_refiner", 3),
           "coordmap": np.ascontiguousarray(sd["coordmap_3d"], np.float32).reshape(-1),
           "anchors": bev_cam3dmap_anchor(60, 128), "embed": np.ascontiguousarray(sd["position_embeddings.weight"], np.float32)}
    for i, k in zip((0, 3, 6), ("0", "1", "2")):
        out["w" + k] = np.ascontiguousarray(sd[f"transformer.{i}.weight"], np.float32)
        out["b" + k] = np.ascontiguousarray(sd[f"transformer.{i}.bias"], np.float32)
    assert out["center_ref"].size == 56 and out["cam_ref"].size == 492
    return out
