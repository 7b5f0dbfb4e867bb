"""Checkpoint compatibility with the reference (README.md:154-164): the state_dict of the REAL
reference modules (the key set and shapes of the official pose_hrnet_* / pose_resnet_* checkpoints, which are plain
state_dicts of these modules) is accepted by the plan's strict packer, in both checkpoint formats the reference loads
(SimpleHRNet.py:117-121), and every tensor lands where the plan says.  Runs wherever the reference is importable
(oracle/_ref or /root/reference); CPU only."""
import numpy as np
import pytest
import torch

from oracle.install_ref import import_reference
from simple_hrnet_b200.engine import Plan, expected_state_dict_keys, pack_state_dict


@pytest.fixture(scope="module")
def ref():
    r = import_reference()
    if r is None:
        pytest.skip("reference not importable here")
    return r


@pytest.mark.parametrize("arch,c,res", [("hrnet", 32, (256, 192)), ("hrnet", 48, (384, 288)), ("poseresnet", 50, (256, 192)),
                                         ("poseresnet", 101, (256, 192))])
def test_real_module_state_dict_is_accepted(ref, arch, c, res):
    _, HRNet, PoseResNet = ref
    torch.manual_seed(0)
    module = HRNet(c, 17) if arch == "hrnet" else PoseResNet(c, 17)
    sd = module.state_dict()
    plan = Plan(arch, c, 17, res, 2)
    req, opt = expected_state_dict_keys(plan.infos)
    assert set(req) | opt == set(sd.keys())                       # exactly the reference's key set (1,754 keys for HRNet)
    buf = pack_state_dict(sd, plan.infos, plan.weight_bytes)
    buf2 = pack_state_dict({"model": sd, "epoch": 3}, plan.infos, plan.weight_bytes)   # training checkpoints (SimpleHRNet.py:118-119)
    assert np.array_equal(buf, buf2)
    # spot-check the layout: first conv (fp32 [cout][kh][kw][cin]) and the head bias
    pi = plan.infos[0]
    w = sd[pi.conv_key.decode() + ".weight"].permute(0, 2, 3, 1).contiguous().numpy().astype(np.float32 if pi.w_f32 else np.float16)
    assert np.array_equal(np.frombuffer(buf[pi.w_offset:pi.w_offset + w.nbytes].tobytes(), dtype=w.dtype), w.ravel())
    head = [p for p in plan.infos if p.conv_key.decode() == "final_layer"][0]
    b = sd["final_layer.bias"].numpy()
    assert np.array_equal(np.frombuffer(buf[head.bias_offset:head.bias_offset + b.nbytes].tobytes(), dtype=np.float32), b)
    # a checkpoint with a missing / an unexpected key is rejected like load_state_dict(strict=True)
    bad = dict(sd); bad.pop("final_layer.bias")
    with pytest.raises(RuntimeError, match="Missing key"):
        pack_state_dict(bad, plan.infos, plan.weight_bytes)
    bad = dict(sd); bad["module.conv1.weight"] = sd["conv1.weight"]
    with pytest.raises(RuntimeError, match="Unexpected key"):
        pack_state_dict(bad, plan.infos, plan.weight_bytes)
