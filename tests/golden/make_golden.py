"""Generate the golden fixtures in tests/golden/ by running the REFERENCE code itself.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

The reference modules are imported through shim packages so that ``romp/__init__.py`` (which imports
the renderer and downloads checkpoints at import time) never executes.  Inputs are
the seeded synthetic parameters of ``romp_b200/synth.py``; outputs are what the reference's own
functions return (ROMPv1, parsing_outputs, SMPL, body_mesh_projection2image, rot6D_to_angular ...).
Nothing from the reference's sources is copied - only its outputs are stored.
"""
import importlib
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference/simple_romp"


def load_reference():
    for name in ("romp", "bev"):
        pkg = types.ModuleType(name)
        pkg.__path__ = [os.path.join(REF, name)]
        sys.modules[name] = pkg
    return {m: importlib.import_module(m) for m in
            ("romp.model", "romp.smpl", "romp.utils", "romp.post_parser")}


def main():
    from romp_b200 import synth
    ref = load_reference()
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count())
    tt = lambda d: {k: torch.from_numpy(np.asarray(v)) for k, v in d.items()}

    # ---- G1: backbone + head maps on one synthetic frame (ROMPv1.forward, model.py:470-481)
    sd = synth.romp_state_dict(0)
    model = ref["romp.model"].ROMPv1().eval()
    model.load_state_dict(tt(sd))
    frames = synth.synthetic_frames(1, seed=0)
    with torch.no_grad():
        center, params = model(torch.from_numpy(frames).float())
        feat = model.backbone(torch.from_numpy(frames).float())
    rs = np.random.RandomState(1)
    pix = rs.choice(4096, size=96, replace=False)
    np.savez_compressed(
        os.path.join(HERE, "maps_seed0.npz"),
        center=center.numpy(),                                    # [1,1,64,64]
        params_pix=pix,
        params_at_pix=params.reshape(1, 145, -1)[0][:, pix].numpy(),   # [145,96]
        params_mean=params.mean((0, 2, 3)).numpy(), params_std=params.std((0, 2, 3)).numpy(),
        feat_ch0=feat[0, 0].numpy(),                               # [128,128]
        feat_mean=feat.mean((0, 2, 3)).numpy(),
    )

    # ---- G2: parse on planted + adversarial center maps (post_parser.py:27-47,135-146)
    cmaps, _ = synth.plant_centers(6, seed=3)
    cmaps[4] = 0.0                                            # a frame with nobody
    cmaps[5, 0, 40, 40] = 0.7; cmaps[5, 0, 41, 41] = 0.6      # suppressed neighbour (SURVEY 8c KAT)
    cmaps[5, 0, 10, 5] = 0.2                                  # below threshold
    cmaps[5, 0, 0, 0] = 0.9; cmaps[5, 0, 63, 63] = 0.8        # corners
    rs = np.random.RandomState(2)
    cmaps = cmaps + rs.uniform(-0.05, 0.05, size=cmaps.shape).astype(np.float32) * (cmaps == 0)
    pmaps = rs.normal(0, 1, size=(6, 145, 64, 64)).astype(np.float32)
    pmaps[:, 0] = np.power(np.float32(1.1), pmaps[:, 0])
    parser = ref["romp.post_parser"].CenterMap(conf_thresh=0.25)
    out = ref["romp.post_parser"].parsing_outputs(torch.from_numpy(cmaps), torch.from_numpy(pmaps), parser)
    bi, fi, yx, sc = parser.parse_centermap(torch.from_numpy(cmaps))
    np.savez_compressed(
        os.path.join(HERE, "parse_seed3.npz"), center_maps=cmaps, params_seed=2,
        batch_ids=bi.numpy(), flat_inds=fi.numpy(), center_yxs=yx.numpy(), scores=sc.numpy(),
        cam=out["cam"].numpy(), global_orient=out["global_orient"].numpy(), body_pose=out["body_pose"].numpy(),
        smpl_betas=out["smpl_betas"].numpy(), smpl_thetas=out["smpl_thetas"].numpy(),
        center_preds=out["center_preds"].numpy(), center_confs=out["center_confs"].numpy(),
    )

    # ---- G3: 6D -> axis-angle incl. every quaternion branch (utils.py:471-682)
    x6 = rs.normal(0, 1, size=(64, 22 * 6)).astype(np.float32)
    x6[0] = np.tile([1, 0, 0, 1, 0, 0], 22)                  # identity
    x6[1] = np.tile([-1, 0, 0, 1, 0, 0], 22)                 # 180 deg turns
    x6[2] = np.tile([1, 0, 0, -1, 0, 0], 22)
    x6[3] = np.tile([0, 1, 1, 0, 0, 0], 22)
    aa = ref["romp.utils"].rot6D_to_angular(torch.from_numpy(x6))
    np.savez_compressed(os.path.join(HERE, "rot6d.npz"), x6=x6, aa=aa.numpy())

    # ---- G4: SMPL forward (smpl.py:62-108) on the synthetic pack, both weight styles
    import tempfile
    for tag, dense in (("sparse", False), ("dense", True)):
        pack = synth.smpl_pack(0, dense_weights=dense)
        with tempfile.NamedTemporaryFile(suffix=".pth") as f:
            torch.save(tt(pack), f.name)
            smpl = ref["romp.smpl"].SMPL(f.name)
        betas = rs.normal(0, 1, size=(5, 10)).astype(np.float32)
        thetas = rs.normal(0, 0.3, size=(5, 72)).astype(np.float32)
        betas[0] = 0; thetas[0] = 0                            # rest pose KAT
        thetas[1, 3:] = 0                                      # global rotation only
        thetas[:, 66:] = 0                                     # hands are exact zeros on the ROMP path
        res = {}
        for ra in (False, True):
            v, j, _ = smpl(torch.from_numpy(betas), torch.from_numpy(thetas), root_align=ra)
            res[ra] = (v.numpy(), j.numpy())
        vsel = rs.choice(6890, size=512, replace=False)
        np.savez_compressed(
            os.path.join(HERE, f"smpl_{tag}.npz"), betas=betas, thetas=thetas, vsel=vsel,
            verts_sel=res[False][0][:, vsel], verts_sum=res[False][0].astype(np.float64).sum(1),
            joints=res[False][1], verts_sel_ra=res[True][0][:, vsel], joints_ra=res[True][1],
        )
        if not dense:
            # ---- G5: projection (post_parser.py:104-114) incl. the reference's cv2 PnP cam_trans
            cam = np.stack([rs.uniform(0.3, 1.2, 5), rs.uniform(-0.5, 0.5, 5), rs.uniform(-0.5, 0.5, 5)], 1).astype(np.float32)
            offsets = torch.Tensor([40, 472, 0, 512, 432, 512])
            v, j, _ = smpl(torch.from_numpy(betas), torch.from_numpy(thetas))
            proj = ref["romp.post_parser"].body_mesh_projection2image(j, torch.from_numpy(cam), vertices=v, input2org_offsets=offsets)
            np.savez_compressed(
                os.path.join(HERE, "project.npz"), cam=cam, offsets=offsets.numpy(), joints=j.numpy(),
                pj2d_org=proj["pj2d_org"].numpy(), cam_trans_pnp=proj["cam_trans"].numpy(),
                verts_camed_org_sel=proj["verts_camed_org"][:, vsel].numpy(), vsel=vsel,
                verts_sel=v[:, vsel].numpy(),
                cam_trans_weak=ref["romp.utils"].convert_cam_to_3d_trans(torch.from_numpy(cam)).numpy(),
            )
            # ---- G5b: the reference's own closed-form fallback (utils.py:429-434): force the cv2 path to raise so
            # estimate_translation() runs estimate_translation_np, incl. its validity mask, on partly off-screen people
            U = ref["romp.utils"]
            cam2 = cam.copy(); cam2[:, 2] -= np.array([0.0, 1.2, 1.6, 2.5, 0.4], np.float32)   # push people off the top
            pj = U.batch_orth_proj(j, torch.from_numpy(cam2), mode='2d')
            j3 = j[:, :24].contiguous().numpy(); p2 = (pj[:, :24, :2].numpy() + 1) * 256
            orig = U.estimate_translation_cv2
            def boom(*a, **k): raise RuntimeError("forced")
            U.estimate_translation_cv2 = boom
            lsq = U.estimate_translation(j3, p2, focal_length=443.4, img_size=np.array([512, 512])).numpy()
            U.estimate_translation_cv2 = orig
            np.savez_compressed(os.path.join(HERE, "cam_trans_lsq.npz"), joints=j.numpy(), cam=cam2, cam_trans_np=lsq)
    print("golden fixtures written to", HERE)
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
