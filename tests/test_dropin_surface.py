"""Drop-in boundary on the CPU: the reference's import names resolve to this implementation, nothing is
evaluated or downloaded at import, ``default_settings`` exists lazily, the settings schema is the reference's, and the
``--cam_trans pnp`` parity mode reproduces the reference's cv2.solvePnPRansac result stored in the golden fixture."""
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def test_reference_import_names():
    import romp
    import bev
    from romp import ROMP, romp_settings          # simple_romp/romp/__init__.py:1
    from bev import BEV, bev_settings             # simple_romp/bev/__init__.py
    import romp_b200
    assert ROMP is romp_b200.ROMP and BEV is romp_b200.bev.BEV
    s = romp.main.default_settings                # main.py:62, evaluated lazily here
    assert s.center_thresh == 0.25 and s.calc_smpl is True and s.root_align is False and s.mode == "image"
    assert romp.main.default_settings is s
    b = bev.main.default_settings
    assert b.center_thresh == 0.1 or b.center_thresh > 0       # bev_settings defaults exist
    assert romp_settings(["--cam_trans", "pnp"]).cam_trans == "pnp"


def test_no_gpu_means_no_model():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from romp import ROMP, romp_settings
    with pytest.raises(RuntimeError):
        ROMP(romp_settings([]), state_dict={}, smpl_pack={})


def test_cam_trans_pnp_mode_equals_reference_fixture():
    """a18 parity mode: the reference's default estimator (utils.py:331-345 via post_parser.py:96-101), inputs and outputs
    from tests/golden/project.npz (written by the reference's body_mesh_projection2image)."""
    from romp_b200.main import estimate_translation_pnp
    z = np.load(os.path.join(HERE, "golden", "project.npz"))
    t = estimate_translation_pnp(z["joints"], z["cam"])
    assert t.dtype == np.float32 and t.shape == z["cam_trans_pnp"].shape
    # same OpenCV build: identical; other builds / RNG states: RANSAC + EPnP agree to ~1e-3
    assert np.abs(t - z["cam_trans_pnp"]).max() < 2e-3 * max(1.0, float(np.abs(z["cam_trans_pnp"]).max()))
