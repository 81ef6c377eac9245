"""Seams S2-S4 through the C ABI against the oracle and the reference-generated golden fixtures."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from oracle import romp_oracle as O
from romp_b200 import _lib, synth
from romp_b200.main import SMPLParser

pytestmark = pytest.mark.gpu
P = lambda t: C.c_void_p(t.data_ptr())


def run_parse(center, params, thresh=0.25, cap=None):
    lib = _lib.load()
    B = center.shape[0]
    cap = cap or B * 64
    dev = "cuda"
    c, p = torch.from_numpy(center).to(dev), torch.from_numpy(params).to(dev)
    z = lambda *s, dt=torch.float32: torch.full(s, -7, dtype=dt, device=dev)
    o = dict(count=z(1, dt=torch.int32), bi=z(cap, dt=torch.int64), fi=z(cap, dt=torch.int64), conf=z(cap),
             pp=z(cap, 145), cam=z(cap, 3), th=z(cap, 72), be=z(cap, 10), cp=z(cap, 2, dt=torch.int64))
    ws = torch.zeros(int(lib.b200romp_parse_workspace_bytes(B)), dtype=torch.uint8, device=dev)
    rc = lib.b200romp_parse(P(c), P(p), B, 64, 10, thresh, cap, P(o["count"]), P(o["bi"]), P(o["fi"]), P(o["conf"]),
                            P(o["pp"]), P(o["cam"]), P(o["th"]), P(o["be"]), P(o["cp"]), P(ws),
                            C.c_void_p(torch.cuda.current_stream().cuda_stream))
    _lib.check(rc, "parse")
    torch.cuda.synchronize()
    n = int(o["count"].item())
    return n, {k: v[:n].cpu().numpy() for k, v in o.items() if k != "count"}


def test_parse_golden_bit_exact(golden_dir):
    z = np.load(os.path.join(golden_dir, "parse_seed3.npz"))
    rs = np.random.RandomState(2)
    cm = z["center_maps"]
    _ = rs.uniform(-0.05, 0.05, size=cm.shape)
    pm = rs.normal(0, 1, size=(6, 145, 64, 64)).astype(np.float32)
    pm[:, 0] = np.power(np.float32(1.1), pm[:, 0])
    n, o = run_parse(cm, pm)
    assert n == len(z["batch_ids"])
    assert np.array_equal(o["bi"], z["batch_ids"]) and np.array_equal(o["fi"], z["flat_inds"])
    assert np.array_equal(o["cp"], z["center_preds"])
    assert np.array_equal(o["conf"][:, None], z["center_confs"])
    assert np.array_equal(o["cam"], z["cam"]) and np.array_equal(o["be"], z["smpl_betas"])
    assert np.abs(o["th"] - z["smpl_thetas"]).max() < 1e-5


def test_parse_random_maps_vs_oracle():
    rs = np.random.RandomState(11)
    cm = rs.normal(0.0, 0.3, size=(5, 1, 64, 64)).astype(np.float32)       # dense random: many local maxima
    cm[3] = -1.0                                                             # nobody
    pm = rs.normal(0, 1, size=(5, 145, 64, 64)).astype(np.float32)
    n, o = run_parse(cm, pm, thresh=0.25)
    ref = O.parsing_outputs(cm, pm, 0.25)
    assert n == len(ref["pred_batch_ids"]) and n > 64
    assert np.array_equal(o["bi"], ref["pred_batch_ids"].numpy()) and np.array_equal(o["fi"], ref["flat_inds"].numpy())
    assert np.array_equal(o["pp"], ref["params_pred"].numpy())
    assert np.abs(o["th"] - ref["smpl_thetas"].numpy()).max() < 1e-5
    assert (np.bincount(o["bi"], minlength=5) <= 64).all()


def test_parse_edge_cases():
    cm = np.zeros((2, 1, 64, 64), np.float32)
    pm = np.zeros((2, 145, 64, 64), np.float32)
    n, _ = run_parse(cm, pm)
    assert n == 0                                                            # nobody anywhere -> count 0 (API returns None)
    cm[1, 0, 10, 10] = 0.5; cm[1, 0, 10, 11] = 0.5                           # plateau: both survive, index asc
    n, o = run_parse(cm, pm)
    assert n == 2 and o["fi"].tolist() == [650, 651] and o["bi"].tolist() == [1, 1]
    cm[:] = 0.6                                                              # constant map: everything is a maximum
    n, o = run_parse(cm, pm)
    assert n == 128 and o["fi"][:64].tolist() == list(range(64))
    n, o = run_parse(cm, pm, cap=70)                                         # capacity clamp
    assert n == 70


def test_rot6d_golden(golden_dir):
    z = np.load(os.path.join(golden_dir, "rot6d.npz"))
    x6 = z["x6"]                                                             # [64,132]
    cm = np.zeros((1, 1, 64, 64), np.float32)
    pm = np.zeros((1, 145, 64, 64), np.float32)
    for i in range(64):                                                      # plant 64 persons, one per row
        y, x = (i // 8) * 8, (i % 8) * 8
        cm[0, 0, y, x] = 0.99 - 0.01 * i
        pm[0, 3:135, y, x] = x6[i]
    n, o = run_parse(cm, pm)
    assert n == 64
    assert np.abs(o["th"][:, :66] - z["aa"]).max() < 1e-5
    assert np.abs(o["th"][:, 66:]).max() == 0.0


@pytest.mark.parametrize("tag,dense", [("sparse", False), ("dense", True)])
def test_smpl_golden_and_oracle(golden_dir, tag, dense):
    z = np.load(os.path.join(golden_dir, f"smpl_{tag}.npz"))
    pack = synth.smpl_pack(0, dense_weights=dense)
    sm = SMPLParser(pack, 0)
    rs = np.random.RandomState(5)
    n = 70                                                                   # > 2 person tiles, ragged tail
    betas = np.concatenate([z["betas"], rs.normal(0, 1, (n - 5, 10)).astype(np.float32)])
    thetas = np.concatenate([z["thetas"], rs.normal(0, 0.4, (n - 5, 72)).astype(np.float32)])
    for ra in (False, True):
        b, t = torch.from_numpy(betas).cuda(), torch.from_numpy(thetas).cuda()
        verts = torch.zeros(n, 6890, 3, device="cuda"); joints = torch.zeros(n, 71, 3, device="cuda")
        ws = torch.zeros(n, sm.ws_floats, device="cuda")
        sm.forward(b, t, n, None, ra, ws, verts, joints, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        v, j = verts.cpu().numpy(), joints.cpu().numpy()
        ov, oj = O.smpl_forward(pack, betas, thetas, root_align=ra)
        assert np.abs(v - ov.numpy()).max() < 1e-4 and np.abs(j - oj.numpy()).max() < 1e-4   # north-star tolerance
        key_v, key_j = ("verts_sel_ra", "joints_ra") if ra else ("verts_sel", "joints")
        assert np.abs(v[:5][:, z["vsel"]] - z[key_v]).max() < 1e-4
        assert np.abs(j[:5] - z[key_j]).max() < 1e-4
    # rest-pose known answer
    assert np.abs(v[0] + 0 - (ov[0].numpy())).max() < 1e-4


def test_smpl_device_count_and_zero_pose():
    pack = synth.smpl_pack(0)
    sm = SMPLParser(pack, 0)
    n = 8
    b = torch.zeros(n, 10, device="cuda"); t = torch.zeros(n, 72, device="cuda")
    verts = torch.full((n, 6890, 3), 123.0, device="cuda"); joints = torch.full((n, 71, 3), 123.0, device="cuda")
    ws = torch.zeros(n, sm.ws_floats, device="cuda")
    cnt = torch.tensor([3], dtype=torch.int32, device="cuda")
    sm.forward(b, t, n, cnt, False, ws, verts, joints, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert np.abs(verts[:3].cpu().numpy() - pack["v_template"][None]).max() < 1e-6    # zero pose/shape -> template
    assert (verts[3:] == 123.0).all() and (joints[3:] == 123.0).all()                  # rows >= *d_count untouched


def test_project_golden(golden_dir):
    z = np.load(os.path.join(golden_dir, "project.npz"))
    lib = _lib.load()
    n = 5
    joints = torch.from_numpy(z["joints"]).cuda(); cam = torch.from_numpy(z["cam"]).cuda()
    pack = synth.smpl_pack(0)
    pj = torch.zeros(n, 71, 2, device="cuda"); weak = torch.zeros(n, 3, device="cuda"); lsq = torch.zeros(n, 3, device="cuda")
    vs = torch.zeros(n, 6890, 3, device="cuda"); vs[:, :512] = torch.from_numpy(z["verts_sel"]).cuda()
    vco = torch.zeros(n, 6890, 3, device="cuda")
    off = (C.c_float * 6)(*[float(v) for v in z["offsets"]])
    rc = lib.b200romp_project(P(joints), P(vs), P(cam), n, None, off, P(pj), P(vco), P(weak), P(lsq),
                              C.c_void_p(torch.cuda.current_stream().cuda_stream))
    _lib.check(rc, "project")
    torch.cuda.synchronize()
    assert np.abs(pj.cpu().numpy() - z["pj2d_org"]).max() < 1e-3           # pixel units, fp32
    assert np.abs(vco[:, :512].cpu().numpy() - z["verts_camed_org_sel"]).max() < 1e-3
    assert np.abs(weak.cpu().numpy() - z["cam_trans_weak"]).max() < 1e-5
    ref = O.project_outputs(torch.from_numpy(z["joints"]), None, z["cam"], z["offsets"])
    assert np.abs(lsq.cpu().numpy() - ref["cam_trans"].numpy()).max() < 1e-3


def test_cam_trans_lsq_matches_reference_fallback(golden_dir):
    """GPU closed-form cam_trans vs the reference's estimate_translation_np path (forced), incl. masked joints and
    the INVALID_TRANS (-1) rows for people with < 4 visible joints."""
    z = np.load(os.path.join(golden_dir, "cam_trans_lsq.npz"))
    lib = _lib.load()
    n = z["joints"].shape[0]
    joints = torch.from_numpy(z["joints"]).cuda(); cam = torch.from_numpy(z["cam"]).cuda()
    lsq = torch.zeros(n, 3, device="cuda")
    off = (C.c_float * 6)(0, 512, 0, 512, 512, 512)
    _lib.check(lib.b200romp_project(P(joints), None, P(cam), n, None, off, None, None, None, P(lsq),
                                    C.c_void_p(torch.cuda.current_stream().cuda_stream)), "project")
    torch.cuda.synchronize()
    ref = z["cam_trans_np"]
    assert (ref == -1).all(1).sum() == 2
    assert np.abs(lsq.cpu().numpy() - ref).max() < 1e-3
