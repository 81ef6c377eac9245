"""cfg1 on the GPU path (BASELINE.json configs[0]; verdict r1 row R50): ROMP with the ResNet-50 backbone
(romp/lib/models/resnet_50.py:19-120 - 7x7 s2 stem, MaxPool, [3,4,6,3] Bottlenecks, three ConvTranspose2d(4,2,1)) as a
libb200romp conv graph, against the fixture written from the reference's own module (tests/golden/resnet50_seed0.npz) and
against the oracle, then the whole hot path for one planted person."""
import os

import numpy as np
import pytest
import torch

from oracle import romp_oracle as O
from romp_b200 import ROMP, romp_settings, synth

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("dtype", [torch.float32, torch.uint8])
def test_resnet50_graph_fp32_vs_reference_fixture_and_oracle(dtype):
    z = np.load(os.path.join(HERE, "golden", "resnet50_seed0.npz"))
    sd, pack = synth.resnet50_state_dict(0), synth.smpl_pack(0)
    frames = synth.synthetic_frames(1, seed=0)
    m = ROMP(romp_settings(["--backbone", "resnet50", "--precision", "fp32", "--max_batch", "1"]), state_dict=sd, smpl_pack=pack)
    with torch.cuda.stream(m.stream):
        c, p = m.run_maps(torch.from_numpy(frames).to(dtype).cuda())
    m.stream.synchronize()
    nb, _ = m._net({torch.uint8: 2, torch.float32: 0}[dtype])
    assert "maxpool" in nb.describe() and " k42 " in nb.describe() and " k7 " in nb.describe()
    feat = torch.zeros(1, 128, 128, 64, device="cuda")
    m.lib.b200romp_net_read_tensor(nb.net, nb.names["backbone_out"], 1, feat.data_ptr(), None)
    torch.cuda.synchronize()
    f = feat.cpu().permute(0, 3, 1, 2)
    e0 = np.abs(f[0, 0].numpy() - z["feat_ch0"]).max()
    e1 = np.abs(f.mean((0, 2, 3)).numpy() - z["feat_mean"]).max()
    oc, op = O.romp_resnet50_maps(sd, frames)
    ec, ep = (c.cpu() - oc).abs().max().item(), (p.cpu() - op).abs().max().item()
    print(f"resnet50 fp32 ({dtype}): feature ch0 max|err| vs reference fixture {e0:.2e}, channel means {e1:.2e}; maps vs oracle {ec:.2e} / {ep:.2e}")
    assert e0 < 1e-4 and e1 < 1e-5 and ec < 1e-4 and ep < 2e-4


def test_cfg1_whole_path_one_planted_person():
    sd, pack = synth.resnet50_state_dict(0), synth.smpl_pack(0)
    frames = synth.synthetic_frames(1, seed=0)
    m = ROMP(romp_settings(["--backbone", "resnet50", "--precision", "fp32", "--max_batch", "1"]), state_dict=sd, smpl_pack=pack)
    planted = np.zeros((1, 1, 64, 64), np.float32)
    planted[0, 0, 30, 20] = 0.8
    out = m.forward_batch(torch.from_numpy(frames), center_override=torch.from_numpy(planted).cuda())
    _, op = O.romp_resnet50_maps(sd, frames)
    ref = O.parsing_outputs(planted, op, 0.25)
    assert len(out["cam"]) == 1 and out["center_preds"].tolist() == [[160, 240]]
    assert np.abs(out["smpl_thetas"] - ref["smpl_thetas"].numpy()).max() < 2e-3
    v, j = O.smpl_forward(pack, out["smpl_betas"], out["smpl_thetas"])
    assert np.abs(out["verts"] - v.numpy()).max() < 1e-4
