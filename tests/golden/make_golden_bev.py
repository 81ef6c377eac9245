"""Golden fixtures for the BEV variant, produced by running the REFERENCE's own bev/model.py and bev/post_parser.py
(build container only).    python tests/golden/make_golden_bev.py"""
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)
from make_golden import load_reference  # noqa: E402


def main():
    import importlib
    from romp_b200 import synth
    load_reference()
    M = importlib.import_module("bev.model")
    PP = importlib.import_module("bev.post_parser")
    torch.set_num_threads(os.cpu_count())
    tt = lambda d: {k: torch.from_numpy(np.asarray(v)) for k, v in d.items()}
    sd = synth.bev_state_dict(0)
    model = M.BEVv1(center_thresh=0.08).eval()
    model.load_state_dict(tt(sd), strict=False)
    frames = synth.synthetic_frames(1, seed=0)
    x = torch.from_numpy(frames).float()
    rs = np.random.RandomState(7)
    with torch.no_grad():
        feat = model.backbone(x)
        c3d, cam3d, cfv = model.coarse2fine_localization(feat)
        fv = model.param_head(feat)
    vox = rs.choice(64 * 128 * 128, size=4096, replace=False)
    np.savez_compressed(os.path.join(HERE, "bev_maps_seed0.npz"), vox=vox,
                        center3d_at=c3d.reshape(-1)[vox].numpy(), cam3d_at=cam3d.reshape(3, -1)[:, vox].numpy(),
                        center_fv=cfv[0, 0].numpy(), fv_pix=fv[0, :, 5::17, 3::19].numpy(),
                        center3d_stats=np.array([c3d.mean().item(), c3d.std().item(), c3d.max().item()]))
    # ---- parse of a planted 3-D center map + the rest of BEVv1.forward on it
    planted = np.zeros((1, 64, 128, 128), np.float32) + rs.uniform(0, 0.05, size=(1, 64, 128, 128)).astype(np.float32)
    cells = [(5, 20, 30, 0.9), (5, 22, 31, 0.85), (40, 100, 64, 0.7), (63, 127, 127, 0.6), (0, 0, 0, 0.5), (30, 64, 64, 0.07),
             (20, 50, 50, 0.3), (20, 50, 56, 0.31)]
    for z, y, xx, v in cells:
        planted[0, z, y, xx] = v
    with torch.no_grad():
        bi, czyx, conf = model.centermap_parser.parse_3dcentermap(torch.from_numpy(planted))
        cams = cam3d[bi, :, czyx[:, 0], czyx[:, 1], czyx[:, 2]]
        params, cam_czyx = model.mesh_parameter_regression(fv, cams, bi)
        pk = PP.pack_params_dict(params)
        trans = PP.denormalize_cam_params_to_trans(pk["cam"])
    np.savez_compressed(os.path.join(HERE, "bev_parse.npz"), planted_cells=np.array(cells, np.float32), noise_seed=7,
                        batch_ids=bi.numpy(), czyx=czyx.numpy(), conf=conf.numpy(), cams=cams.numpy(),
                        params_pred=params.numpy(), cam_czyx=cam_czyx.numpy(), smpl_thetas=pk["smpl_thetas"].numpy(),
                        smpl_betas=pk["smpl_betas"].numpy(), cam_trans=trans.numpy())
    # ---- SMPL-A / SMIL split, perspective projection and the two post filters
    pack_a, pack_s = synth.smpl_pack(0, num_betas=11), synth.smpl_pack(1)
    with tempfile.TemporaryDirectory() as d:
        torch.save(tt(pack_a), os.path.join(d, "a.pth")); torch.save(tt(pack_s), os.path.join(d, "s.pth"))
        parser = PP.SMPLA_parser(os.path.join(d, "a.pth"), os.path.join(d, "s.pth"))
    n = 9
    betas = rs.normal(0, 1, (n, 11)).astype(np.float32); betas[:, 10] = rs.uniform(0, 1, n); betas[2, 10] = 0.95; betas[5, 10] = 0.81
    thetas = rs.normal(0, 0.3, (n, 72)).astype(np.float32)
    cam = np.stack([rs.uniform(0.2, 1.2, n), rs.uniform(-0.6, 0.6, n), rs.uniform(-0.6, 0.6, n)], 1).astype(np.float32)
    cam[4] = cam[3] + 0.002; cam[7, 0] = 0.05; cam[7, 1:] = 0.9           # a near-duplicate and a far outlier
    thetas[4] = thetas[3]; betas[4] = betas[3]
    with torch.no_grad():
        v, j, _ = parser(torch.from_numpy(betas), torch.from_numpy(thetas))
        out = {"verts": v, "joints": j, "cam": torch.from_numpy(cam), "smpl_thetas": torch.from_numpy(thetas),
               "center_confs": torch.from_numpy(rs.uniform(0.1, 1, n).astype(np.float32)),
               "params_pred": torch.zeros(n, 146)}
        out["cam_trans"] = PP.denormalize_cam_params_to_trans(out["cam"])
        proj = PP.body_mesh_projection2image(j, out["cam"], vertices=None, input2org_offsets=torch.Tensor([0, 512, 0, 512, 512, 512]))
        out.update(proj)
        pj2d_before = out["pj2d"].clone().numpy()
        o1 = PP.suppressing_redundant_prediction_via_projection(dict(out), (512, 512), thresh=20)
        kept1 = [int(np.argmin(np.abs(cam[:, 1] - c))) for c in o1["cam"][:, 1].numpy()]
        o2 = PP.remove_outlier(dict(o1), relative_scale_thresh=1.6)
        kept2 = [int(np.argmin(np.abs(cam[:, 1] - c))) for c in o2["cam"][:, 1].numpy()]
    vsel = rs.choice(6890, 256, replace=False)
    np.savez_compressed(os.path.join(HERE, "bev_post.npz"), betas=betas, thetas=thetas, cam=cam, vsel=vsel,
                        verts_sel=v[:, vsel].numpy(), joints=j.numpy(), cam_trans=out["cam_trans"].numpy(),
                        pj2d=pj2d_before, kept_after_nms=np.array(kept1), kept_after_outlier=np.array(kept2))
    for f in sorted(os.listdir(HERE)):
        if f.startswith("bev_"):
            print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
