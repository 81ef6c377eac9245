"""BEV variant on the GPU (romp_b200.bev.BEV, all through the C ABI) against the BEV oracle, stage-wise."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from oracle import bev_oracle as B
from oracle import romp_oracle as O
from romp_b200 import _lib, synth
from romp_b200.bev import BEV, bev_settings

pytestmark = pytest.mark.gpu
P = lambda t: C.c_void_p(t.data_ptr())


@pytest.fixture(scope="module")
def params():
    return synth.bev_state_dict(0), synth.smpl_pack(0, num_betas=11), synth.smpl_pack(1)


def planted_volume(nframes, seed=3):
    rs = np.random.RandomState(seed)
    vol = rs.uniform(0, 0.05, size=(nframes, 64, 128, 128)).astype(np.float32)
    truth = []
    for b in range(nframes):
        cells = []
        for _ in range(rs.randint(1, 7)):
            z, y, x = rs.randint(0, 64), rs.randint(0, 128), rs.randint(0, 128)
            if all(max(abs(z - c[0]), abs(y - c[1]), abs(x - c[2])) >= 3 for c in cells):
                cells.append((z, y, x))
        vals = np.sort(rs.uniform(0.2, 1.0, len(cells)).astype(np.float32))[::-1]
        for (z, y, x), v in zip(cells, vals):
            vol[b, z, y, x] = v
        truth.append(list(zip(cells, vals)))
    return vol, truth


def make(params, precision, B_=2, extra=()):
    s = bev_settings(["--precision", precision, "--max_batch", str(B_), *extra])
    return BEV(s, state_dict=params[0], smpla_pack=params[1], smil_pack=params[2])


def test_settings_defaults_match_reference():
    s = bev_settings([])
    assert s.crowd is True and s.center_thresh == 0.08 and s.nms_thresh == 20 and s.relative_scale_thresh == 1.6
    s = bev_settings(["--crowd"])
    assert s.crowd is False and s.center_thresh == 0.1


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_model_stages_vs_oracle(params, precision):
    nf = 2
    frames = synth.synthetic_frames(nf, seed=11)
    vol, truth = planted_volume(nf)
    m = make(params, precision, nf)
    with torch.cuda.stream(m.stream):
        m.run_model(torch.from_numpy(frames).cuda(), center3d_override=torch.from_numpy(vol).cuda())
    m.stream.synchronize()
    ref = B.bev_model(params[0], frames, m.settings.center_thresh, center3d_override=vol)
    b = m.buf
    n = int(b["count"].item())
    tol = 1.0 if precision == "fp32" else 40.0
    # maps: det_head output, refined 3-D centre map (own, not the planted one), front-view features
    sdt = O.to_torch_sd(params[0])
    feat = O.hrnet32_forward(sdt, torch.from_numpy(frames).float())
    c3d_ref, cam3d_ref, cfv_ref = B.coarse2fine(sdt, feat)
    e_fv = (b["maps_fv"][:, :1].cpu() - cfv_ref).abs().max().item()
    e_c3d = (b["center3d"].cpu() - c3d_ref).abs().max().item()
    print(f"{precision}: center_fv err {e_fv:.2e}, center3d err {e_c3d:.2e} (std {c3d_ref.std():.3f})")
    assert e_fv < 2e-3 * tol and e_c3d < 2e-3 * tol
    # parse on the identical planted volume: integers bit-exact
    assert n == len(ref["pred_batch_ids"]) == sum(len(t) for t in truth)
    assert np.array_equal(b["batch_ids"][:n].cpu().numpy(), ref["pred_batch_ids"].numpy())
    assert np.array_equal(b["czyx"][:n].cpu().numpy(), ref["pred_czyxs"].numpy())
    assert np.array_equal(b["conf"][:n].cpu().numpy(), ref["center_confs"].numpy())
    # lazy cam refiner + anchors + sampling + MLP
    pp, rp = b["params_pred"][:n].cpu().numpy(), ref["params_pred"].numpy()
    print(f"{precision}: cams err {np.abs(pp[:, :3] - rp[:, :3]).max():.2e}  mlp err {np.abs(pp[:, 3:] - rp[:, 3:]).max():.2e}")
    assert np.abs(pp[:, :3] - rp[:, :3]).max() < 2e-3 * tol
    if precision == "fp32":
        assert np.array_equal(b["cam_czyx"][:n].cpu().numpy(), ref["cam_czyx"].numpy())
        assert np.abs(pp[:, 3:] - rp[:, 3:]).max() < 5e-3
    # unpack on the GPU's own params_pred: thetas / betas / cam_trans
    pk = O.pack_params(torch.from_numpy(pp), num_betas=11)
    assert np.abs(b["thetas"][:n].cpu().numpy() - pk["smpl_thetas"].numpy()).max() < 1e-5
    assert np.array_equal(b["betas"][:n].cpu().numpy(), pk["smpl_betas"].numpy())
    assert np.abs(b["cam_trans"][:n].cpu().numpy() - B.cam_to_trans(pk["cam"]).numpy()).max() < 1e-4


def test_post_filters_golden(params, golden_dir):
    """SMPL-A/SMIL merge, perspective projection, projection-NMS and outlier removal vs the reference's own results."""
    z = np.load(os.path.join(golden_dir, "bev_post.npz"))
    m = make(params, "fp32", 1)
    n = len(z["cam"])
    b = m.buf
    b["betas"][:n] = torch.from_numpy(z["betas"]).cuda(); b["thetas"][:n] = torch.from_numpy(z["thetas"]).cuda()
    b["cam"][:n] = torch.from_numpy(z["cam"]).cuda(); b["cam_trans"][:n] = torch.from_numpy(z["cam_trans"]).cuda()
    b["batch_ids"][:n] = 0
    b["count"][0] = n
    with torch.cuda.stream(m.stream):
        m.run_post(1, [0, 512, 0, 512, 512, 512], 512.0)
    m.stream.synchronize()
    assert np.abs(b["verts"][:n].cpu().numpy()[:, z["vsel"]] - z["verts_sel"]).max() < 1e-4
    assert np.abs(b["joints"][:n].cpu().numpy() - z["joints"]).max() < 1e-4
    assert np.abs(b["pj2d_org"][:n].cpu().numpy() - z["pj2d"]).max() < 5e-2          # pixels
    n2 = int(b["count2"].item())
    assert b["sel"][:n2].cpu().numpy().tolist() == z["kept_after_outlier"].tolist()
    assert np.array_equal(m.out["cam"][:n2].cpu().numpy(), z["cam"][z["kept_after_outlier"]])


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_forward_batch_dict(params, precision):
    nf = 2
    frames = synth.synthetic_frames(nf, seed=12)
    vol, truth = planted_volume(nf, seed=5)
    m = make(params, precision, nf)
    out = m.forward_batch(torch.from_numpy(frames), center3d_override=torch.from_numpy(vol).cuda())
    assert set(out) == {"smpl_thetas", "smpl_betas", "cam", "cam_trans", "params_pred", "center_confs", "pred_batch_ids",
                        "verts", "joints", "pj2d_org"}
    n = len(out["cam"])
    assert out["smpl_betas"].shape == (n, 11) and out["params_pred"].shape == (n, 146) and out["verts"].shape == (n, 6890, 3)
    # the same post pipeline in the oracle on the GPU's params_pred (stage-wise): survivors and meshes
    pk = O.pack_params(torch.from_numpy(m.buf["params_pred"][: int(m.buf["count"])].cpu().numpy()), num_betas=11)
    v, j = B.smpla_forward(params[1], params[2], pk["smpl_betas"], pk["smpl_thetas"])
    trans = B.cam_to_trans(pk["cam"])
    pj = O.to_org_image(B.perspective_project(j, trans), [0, 512, 0, 512, 512, 512])
    bids = m.buf["batch_ids"][: int(m.buf["count"])].cpu().tolist()
    keep = []
    for fb in sorted(set(bids)):
        idx = [i for i, q in enumerate(bids) if q == fb]
        k1 = [idx[i] for i in B.suppress_redundant(pj[idx], pk["cam"][idx], (512, 512), m.settings.nms_thresh)]
        keep += [k1[i] for i in B.remove_outlier(trans[k1], pk["cam"][k1], m.settings.relative_scale_thresh)]
    assert n == len(keep)
    assert np.abs(out["verts"] - v[keep].numpy()).max() < 1e-4 and np.abs(out["joints"] - j[keep].numpy()).max() < 1e-4
    assert np.array_equal(out["pred_batch_ids"], np.array(bids)[keep])
    # nobody detected -> None
    assert m.forward_batch(torch.from_numpy(frames), center3d_override=torch.zeros(nf, 64, 128, 128).cuda()) is None
