"""diagnostic: where does the tensor-core SMPL blend deviate?  per-person / per-column-tile error of v_posed and verts"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from romp_b200 import synth
from romp_b200.main import SMPLParser
from oracle import romp_oracle as O

for n in (70, 300):
    pack = synth.smpl_pack(0)
    sm = SMPLParser(pack, 0)
    rs = np.random.RandomState(5)
    betas = rs.normal(0, 1, (n, 10)).astype(np.float32); thetas = rs.normal(0, 0.4, (n, 72)).astype(np.float32)
    b, t = torch.from_numpy(betas).cuda(), torch.from_numpy(thetas).cuda()
    verts = torch.zeros(n, 6890, 3, device="cuda"); joints = torch.zeros(n, 71, 3, device="cuda")
    ws = torch.full((n, sm.ws_floats), float("nan"), device="cuda")
    sm.forward(b, t, n, None, False, ws, verts, joints, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    ov, oj = O.smpl_forward(pack, betas, thetas)
    err = (verts.cpu() - ov).abs().reshape(n, -1)
    e = torch.nan_to_num(err, nan=9.0)
    print(f"n={n}: verts max err {e.max():.3e}; bad persons: {[int(i) for i in torch.nonzero(e.max(1).values > 1e-4).flatten()[:40]]}")
    ct = e.max(0).values.reshape(-1)                                  # per coordinate
    bad_cols = torch.nonzero(ct > 1e-4).flatten()
    print("  bad coordinate range:", (int(bad_cols.min()), int(bad_cols.max())) if len(bad_cols) else None, "count", len(bad_cols))
    # v_posed directly
    S = np.asarray(pack["shapedirs"], np.float64).reshape(20670, 10); Pd = np.asarray(pack["posedirs"], np.float64)
    R = O.batch_rodrigues(torch.from_numpy(thetas).reshape(-1, 3)).reshape(n, 24, 3, 3).numpy().astype(np.float64)
    feat = np.concatenate([betas.astype(np.float64), (R[:, 1:] - np.eye(3)).reshape(n, 207)], 1)
    vp_ref = np.asarray(pack["v_template"], np.float64).reshape(1, -1) + feat @ np.concatenate([S.T, Pd], 0)
    # workspace layout (smpl.cu): n records of 1424 floats, then v_posed [81 coordinate tiles][n][256]
    vp = ws.flatten()[n * 1424:n * 1424 + 81 * n * 256].reshape(81, n, 256).permute(1, 0, 2).reshape(n, 20736)[:, :20670].cpu().double().numpy()
    ev = np.nan_to_num(np.abs(vp - vp_ref), nan=9.0)
    print(f"  v_posed max err {ev.max():.3e}; per-person max (first 12): {np.round(ev.max(1)[:12], 6)}; bad persons {np.nonzero(ev.max(1) > 1e-5)[0][:40]}")
    bc = np.nonzero(ev.max(0) > 1e-5)[0]
    print("  v_posed bad columns:", (bc.min(), bc.max(), len(bc)) if len(bc) else None, " distinct col%256:", sorted(set((bc % 256).tolist()))[:20] if len(bc) else None)
