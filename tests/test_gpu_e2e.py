"""Whole hot path on the GPU (ROMP call surface) against the CPU oracle, stage-wise as SURVEY 8c prescribes:
P1 maps, P2 parse on identical maps (bit-exact indices), P3 thetas, P4 SMPL (1e-4), P5 MPJPE."""
import numpy as np
import pytest
import torch

from oracle import romp_oracle as O
from romp_b200 import ROMP, romp_settings, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def params():
    return synth.romp_state_dict(0), synth.smpl_pack(0)


@pytest.fixture(scope="module")
def oracle_maps(params):
    frames = synth.synthetic_frames(2, seed=5)
    return frames, O.romp_maps(params[0], frames)


def make(params, precision, max_batch=2, extra=()):
    s = romp_settings(["--precision", precision, "--max_batch", str(max_batch), *extra])
    return ROMP(s, state_dict=params[0], smpl_pack=params[1])


def test_p1_maps_fp32_parity(params, oracle_maps):
    frames, (oc, op) = oracle_maps
    m = make(params, "fp32")
    for dt in (torch.uint8, torch.float32):                      # both input dtypes of the stem
        with torch.cuda.stream(m.stream):
            c, p = m.run_maps(torch.from_numpy(frames).to(dt).cuda())
        m.stream.synchronize()
        ec, ep = (c.cpu() - oc).abs().max().item(), (p.cpu() - op).abs().max().item()
        print(f"fp32 engine vs oracle: center {ec:.3e} params {ep:.3e}")
        assert ec < 2e-5 and ep < 3e-5       # ~100 fp32 layers, different summation order; measured 3.4e-6 / 5.0e-6 on B200
    # the backbone feature map itself
    nb, _ = m._net(0)                           # F32 input net
    feat = torch.zeros(2, 128, 128, 32, device="cuda")
    m.lib.b200romp_net_read_tensor(nb.net, nb.names["backbone_out"], 2, feat.data_ptr(), None)
    torch.cuda.synchronize()
    of = O.hrnet32_forward(O.to_torch_sd(params[0]), torch.from_numpy(frames).float())
    assert (feat.cpu().permute(0, 3, 1, 2) - of).abs().max().item() < 5e-5


def test_p1_maps_bf16_tolerance(params, oracle_maps):
    frames, (oc, op) = oracle_maps
    m = make(params, "bf16")
    with torch.cuda.stream(m.stream):
        c, p = m.run_maps(torch.from_numpy(frames).cuda())
    m.stream.synchronize()
    ec, ep = (c.cpu() - oc).abs().max().item(), (p.cpu() - op).abs().max().item()
    print(f"bf16 engine vs fp32 oracle: center max|err| {ec:.3e} (std {oc.std():.3f}) params {ep:.3e} (std {op.std():.3f})")
    # bf16 storage of every activation through ~100 layers; SURVEY 8c measured 0.072/0.088 for a bf16-autocast
    # run of the reference itself - we must be in that regime, not better than fp32 and not broken.
    assert ec < 0.08 and ep < 0.11          # 2x the values measured on B200 (0.039 / 0.054); maps have std 0.12 / 1.18
    assert np.corrcoef(c.cpu().numpy().ravel(), oc.numpy().ravel())[0, 1] > 0.995
    # the 64 BasicBlock convs of the 32-channel branch run pixel-pair folded (net.cu fold_pixel_pairs) - this parity covers them
    nb, _ = m._net(1)
    import os
    if os.environ.get("B200ROMP_TC_NO_FOLD") != "1":
        assert nb.describe().count("pixel-pairs") == 64


@pytest.mark.parametrize("precision", ["fp32", "tf32", "bf16"])
def test_p2_to_p5_planted_batch(params, precision):
    B = 4
    frames = synth.synthetic_frames(B, seed=9)
    planted, truth = synth.plant_centers(B, seed=9)
    m = make(params, precision, max_batch=B)
    out = m.forward_batch(torch.from_numpy(frames), center_override=torch.from_numpy(planted).cuda())
    gpu_params = m.buf["params_maps"][:B].cpu()
    ref = O.parsing_outputs(planted, gpu_params, 0.25)           # identical maps -> indices must be bit-exact
    n = sum(len(t) for t in truth)
    assert len(out["cam"]) == n == len(ref["pred_batch_ids"])
    assert np.array_equal(out["pred_batch_ids"], ref["pred_batch_ids"].numpy())
    assert np.array_equal(out["center_preds"], ref["center_preds"].numpy())
    flat = np.concatenate([[f for f, _ in t] for t in truth])
    assert np.array_equal(out["center_preds"][:, 0] // 8 + out["center_preds"][:, 1] // 8 * 64, flat)
    assert np.array_equal(out["cam"], ref["cam"].numpy()) and np.array_equal(out["smpl_betas"], ref["smpl_betas"].numpy())
    assert np.abs(out["smpl_thetas"] - ref["smpl_thetas"].numpy()).max() < 1e-5
    v, j = O.smpl_forward(params[1], out["smpl_betas"], out["smpl_thetas"])
    assert np.abs(out["verts"] - v.numpy()).max() < 1e-4 and np.abs(out["joints"] - j.numpy()).max() < 1e-4
    pr = O.project_outputs(j, None, ref["cam"], [0, 512, 0, 512, 512, 512])
    assert np.abs(out["pj2d_org"] - pr["pj2d_org"].numpy()).max() < 2e-3
    assert np.abs(out["cam_trans"] - pr["cam_trans"].numpy()).max() < 2e-3
    assert set(out.keys()) == {"cam", "global_orient", "body_pose", "smpl_betas", "smpl_thetas", "center_preds",
                               "center_confs", "cam_trans", "verts", "joints", "pj2d_org", "pred_batch_ids"}
    assert out["center_confs"].shape == (n, 1) and out["body_pose"].shape == (n, 69) and out["center_preds"].dtype == np.int64
    # P5: end to end against the fp32 oracle (its own maps for the params, same planted detections)
    full = O.romp_forward(params[0], params[1], frames, center_override=planted)
    mp = O.mpjpe_mm(out["joints"], full["joints"])
    print(f"{precision}: persons {n}  MPJPE vs fp32 oracle {mp:.3f} mm")
    # measured on B200: fp32 0.001 mm, tf32 0.77 mm (TF32 rounding of the conv operands, like the reference's GPU path),
    # bf16 9.1-10.7 mm (bf16 storage of every activation); bounds = 2x measured
    assert mp < {"fp32": 0.01, "tf32": 1.6, "bf16": 22.0}[precision]


def test_nobody_returns_none_and_single_image_forward(params):
    m = make(params, "fp32", max_batch=1)
    frames = synth.synthetic_frames(1, seed=5)
    assert m.forward_batch(torch.from_numpy(frames)) is None      # synthetic weights: no natural detections
    # single BGR image through the reference-style __call__, non-square -> pad info is used
    rs = np.random.RandomState(0)
    img = rs.randint(0, 256, size=(300, 400, 3)).astype(np.uint8)
    assert m(img) is None
    import cv2
    from romp_b200.main import img_preprocess
    inp, pad = img_preprocess(img)
    assert inp.shape == (1, 512, 512, 3) and pad.tolist() == [50, 350, 0, 400, 300, 400]
    planted, _ = synth.plant_centers(1, seed=1)
    out = m.forward_batch(torch.from_numpy(inp), offsets=pad, center_override=torch.from_numpy(planted).cuda())
    full = O.romp_maps(params[0], inp)
    ref = O.parsing_outputs(planted, full[1], 0.25)
    assert np.array_equal(out["center_preds"], ref["center_preds"].numpy())
    v, j = O.smpl_forward(params[1], ref["smpl_betas"], ref["smpl_thetas"])
    pr = O.project_outputs(j, None, ref["cam"], pad)
    err_px = np.abs(out["pj2d_org"] - pr["pj2d_org"].numpy()).max()
    print(f"pj2d_org max err {err_px:.2e} px")
    assert err_px < 0.02                                                   # pixels in the 400x300 original image (fp32 engine)


def test_forward_batches_pipeline_equals_forward_batch(params):
    """The streaming API (H2D / kernels / D2H of neighbouring batches overlapped) returns, batch by batch, exactly
    what the synchronous forward_batch returns."""
    B = 3
    m = make(params, "bf16", max_batch=B)
    batches = [synth.synthetic_frames(B, seed=20 + i) for i in range(4)]
    planted, _ = synth.plant_centers(B, seed=4)
    planted = torch.from_numpy(planted).cuda()
    ref = []
    for fr in batches:
        o = m.forward_batch(torch.from_numpy(fr), center_override=planted)
        ref.append({k: np.array(v) for k, v in o.items()})
    got = []
    for o in m.forward_batches((torch.from_numpy(fr).pin_memory() for fr in batches), center_override=planted):
        got.append({k: np.array(v) for k, v in o.items()})
    assert len(got) == len(ref)
    for a, b in zip(got, ref):
        assert set(a) == set(b)
        for k in a:
            assert np.array_equal(a[k], b[k]), k
    # batches really differ from each other (the pipeline did not return a stale slot)
    assert not np.array_equal(got[0]["smpl_thetas"], got[1]["smpl_thetas"])


# ------------------------------------------------------------------------------------------------
# Natural (un-planted) detections: the engines' OWN center maps decide who is detected.
# ------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def firing(params):
    """Weights whose center head fires on its own (synth.calibrate_center_head), the fp32 oracle's result on 4 frames, and
    the same oracle under TF32-equivalent operand rounding = the reference's default GPU arithmetic."""
    frames = synth.synthetic_frames(4, seed=11)
    c, _ = O.romp_maps(params[0], frames)
    sd, shift, margin = synth.calibrate_center_head(params[0], c.numpy())
    ref = O.romp_forward(sd, params[1], frames)
    ref_maps = O.romp_maps(sd, frames)
    with O.conv_rounding("tf32"):
        ref_tf32 = O.romp_forward(sd, params[1], frames)
        tf32_maps = O.romp_maps(sd, frames)
    assert ref is not None and 4 <= len(ref["cam"]) <= 40
    return dict(sd=sd, frames=frames, margin=margin, ref=ref, ref_maps=ref_maps, ref_tf32=ref_tf32, tf32_maps=tf32_maps)


def _det_keys(out):
    cp = np.asarray(out["center_preds"])
    return [(int(b), int(x), int(y)) for b, (x, y) in zip(np.asarray(out["pred_batch_ids"]), cp)]


def test_oracle_tf32_yardstick(firing):
    """The yardstick itself: the reference's TF32 GPU arithmetic keeps the detection set of its fp32 CPU arithmetic on
    these inputs (the calibration margin is several times the TF32 map error) and moves the joints by well under 1 mm."""
    ref, t = firing["ref"], firing["ref_tf32"]
    ec = (firing["tf32_maps"][0] - firing["ref_maps"][0]).abs().max().item()
    print(f"oracle TF32 vs fp32: center max|err| {ec:.2e}, decision margin {firing['margin']:.2e}, persons {len(ref['cam'])}")
    assert firing["margin"] > 3 * ec
    assert _det_keys(t) == _det_keys(ref)
    print(f"oracle TF32 vs fp32 MPJPE {O.mpjpe_mm(t['joints'], ref['joints']):.4f} mm")


@pytest.mark.parametrize("precision", ["fp32", "tf32", "bf16"])
def test_natural_detections_vs_oracle(params, firing, precision):
    """P5 without planting: person count, (frame, x, y) of every detection and its order must equal the fp32 oracle's for
    the fp32 and TF32 engines; MPJPE of the TF32 engine must be in the regime of the oracle's own TF32 run."""
    f = firing
    m = make((f["sd"], params[1]), precision, max_batch=4)
    out = m.forward_batch(torch.from_numpy(f["frames"]))
    assert out is not None
    c = m.buf["center_maps"][:4].cpu()
    ec = (c - f["ref_maps"][0]).abs()
    e_tf = (f["tf32_maps"][0] - f["ref_maps"][0]).abs()
    ep = (m.buf["params_maps"][:4].cpu() - f["ref_maps"][1]).abs()
    ep_tf = (f["tf32_maps"][1] - f["ref_maps"][1]).abs()
    got, want = _det_keys(out), _det_keys(f["ref"])
    common = [k for k in got if k in set(want)]
    print(f"{precision}: center map max|err| {ec.max():.2e} mean {ec.mean():.2e} (oracle-TF32: {e_tf.max():.2e} / {e_tf.mean():.2e}); "
          f"params map {ep.max():.2e} / {ep.mean():.2e} (oracle-TF32: {ep_tf.max():.2e} / {ep_tf.mean():.2e}); "
          f"detections {len(got)} vs oracle {len(want)}, common {len(common)}")
    if precision in ("fp32", "tf32"):
        assert got == want, "detection set / order differs from the fp32 oracle"
        assert np.array_equal(out["center_preds"], f["ref"]["center_preds"].numpy())
        mp = O.mpjpe_mm(out["joints"], f["ref"]["joints"])
        mp_tf = O.mpjpe_mm(f["ref_tf32"]["joints"], f["ref"]["joints"])
        print(f"{precision}: MPJPE vs fp32 oracle {mp:.4f} mm (oracle-TF32 vs oracle-fp32: {mp_tf:.4f} mm)")
        if precision == "fp32":
            assert mp < 0.01 and ec.max() < 2e-5
        else:
            # Same arithmetic as the oracle's TF32 run up to summation order and BN folding before/after rounding: the
            # two are independent realisations of one error distribution.  Over the 2 x 10^6 map values the means
            # agree within a few % (measured on B200: center 4.82e-4 vs 4.85e-4); the MPJPE is an average over only
            # 8 persons x 24 joints and scatters more (measured 0.91 vs 0.65 mm) - bounded by 2x the yardstick.
            assert ec.mean() <= 1.25 * e_tf.mean() + 1e-5 and ec.max() <= 1.5 * e_tf.max()
            assert ep.mean() <= 1.25 * ep_tf.mean() + 1e-5 and ep.max() <= 2.0 * ep_tf.max()
            assert mp <= 2.0 * mp_tf + 0.01
    else:
        # bf16 engine: 8-bit mantissa storage of every activation - detections may flip near the threshold; what it
        # keeps must be the oracle's, and most of the oracle's must be found
        assert len(common) >= 0.75 * len(want) and len(got) <= 1.25 * len(want) + 1
        idx_g = [got.index(k) for k in common]
        idx_w = [want.index(k) for k in common]
        mp = O.mpjpe_mm(out["joints"][idx_g], f["ref"]["joints"].numpy()[idx_w])
        print(f"bf16: MPJPE on the {len(common)} common detections {mp:.2f} mm")
        assert mp < 25.0
