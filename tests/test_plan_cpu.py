"""CPU tests of the host logic: C-ABI exports, plan construction (no CUDA needed), the parameter
manifest vs the reference's state_dict names, the weight packer, and a torch emulation of the
plan's op list against the oracle (graph wiring / fusion / residual correctness)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import simple_hrnet_b200 as S
from simple_hrnet_b200 import _lib
from simple_hrnet_b200.engine import Plan, pack_state_dict, expected_state_dict_keys
from oracle import hrnet_oracle as O
from tests.plan_emulator import emulate

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    """include/hrnet_b200.h is the boundary: every function it declares must be exported."""
    hdr = open(os.path.join(ROOT, "include", "hrnet_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(hrnet_[a-z_0-9]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    lib = ctypes.CDLL(S.library_path())
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/hrnet_b200.h but not exported"
    assert declared == set(_lib.SYMBOLS), "ctypes binding table out of sync with the header"


def test_plan_create_rejects_bad_configs():
    for kw in (dict(resolution=(250, 192)), dict(max_batch=0), dict(nof_joints=33), dict(c=20)):
        args = dict(arch="hrnet", c=32, nof_joints=17, resolution=(256, 192), max_batch=2)
        args.update(kw)
        with pytest.raises(S.HrnetError):
            Plan(**args)
    with pytest.raises(ValueError, match="Wrong model name"):
        Plan("vgg", 32, 17, (256, 192), 1)
    with pytest.raises(S.HrnetError, match="50, 101 or 152"):
        Plan("poseresnet", 18, 17, (256, 192), 1)


@pytest.mark.parametrize("c", [32, 48])
def test_manifest_matches_reference_state_dict_names(c):
    p = Plan("hrnet", c, 17, (256, 192), 1)
    req, opt = expected_state_dict_keys(p.infos)
    spec = O.hrnet_param_spec(c, 17)     # verified == HRNet(c,17).state_dict() by make_golden.py
    names = [k for k, _, _ in spec]
    assert set(req) | opt == set(names)
    assert len(p.infos) == 293           # 293 convs
    shapes = {k: s for k, s, _ in spec}
    for pi in p.infos:
        assert shapes[pi.conv_key.decode() + ".weight"] == (pi.cout, pi.cin, pi.kh, pi.kw)


def test_poseresnet_manifest_matches_reference_state_dict_names():
    p = Plan("poseresnet", 50, 17, (256, 192), 1)
    req, opt = expected_state_dict_keys(p.infos)
    names = [k for k, _, _ in O.poseresnet_param_spec(50, 17)]
    assert set(req) | opt == set(names)


def test_packer_is_strict_like_load_state_dict():
    p = Plan("hrnet", 32, 17, (64, 64), 1)
    sd = O.make_state_dict(O.hrnet_param_spec(32, 17), seed=0)
    bad = dict(sd); bad.pop("conv1.weight")
    with pytest.raises(RuntimeError, match="Missing key"):
        pack_state_dict(bad, p.infos, p.weight_bytes)
    bad = dict(sd); bad["module.extra"] = torch.zeros(1)
    with pytest.raises(RuntimeError, match="Unexpected key"):
        pack_state_dict(bad, p.infos, p.weight_bytes)
    # training checkpoints wrap the weights in {'model': ...} (SimpleHRNet.py:118-119)
    a = pack_state_dict({"model": sd, "epoch": 3}, p.infos, p.weight_bytes)
    b = pack_state_dict(sd, p.infos, p.weight_bytes)
    assert np.array_equal(a, b)


def test_plan_graph_emulation_matches_oracle_hrnet(golden_dir):
    """Op list + packed weights, executed with torch on CPU, reproduce the reference heat-maps:
    tight in `exact` mode (only the fp16 weight rounding remains), < 1e-3-scale with fp16 activations."""
    g = np.load(os.path.join(golden_dir, "w32_64x64_n2_rand.npz"))
    sd = O.make_state_dict(O.hrnet_param_spec(32, 17), seed=int(g["wseed"]), bn=str(g["bn"]))
    x = torch.randn(2, 3, 64, 64, generator=torch.Generator().manual_seed(int(g["xseed"])))
    p = Plan("hrnet", 32, 17, (64, 64), 2)
    packed = pack_state_dict(sd, p.infos, p.weight_bytes)
    desc = p.describe()
    ref = torch.from_numpy(g["heatmaps"])
    hm_exact, _ = emulate(desc, p.infos, packed, x, exact=True)
    assert (hm_exact - ref).abs().max() < 3e-4 * max(1.0, ref.abs().max().item())
    hm, _ = emulate(desc, p.infos, packed, x)
    err = (hm - ref).abs().max().item()
    assert err < 2e-3 * max(1.0, ref.abs().max().item()), err
    # every tensor written once before it is read; every op's inputs are produced by its deps or earlier same-stream ops
    seen = {desc["input"]}
    for i, op in enumerate(desc["ops"]):
        ins = [t for t in [op["in"], op["res"]] + op["src"][: op["nsrc"]] if t >= 0]
        assert all(t in seen for t in ins), (i, op["name"])
        assert all(d < i for d in op["deps"])
        if op["out"] >= 0:
            seen.add(op["out"])


def test_plan_graph_emulation_matches_oracle_poseresnet():
    sd = O.make_state_dict(O.poseresnet_param_spec(50, 17), seed=2, bn="random")
    x = torch.randn(1, 3, 64, 64, generator=torch.Generator().manual_seed(9))
    ref = O.poseresnet_forward(sd, x, 50)
    p = Plan("poseresnet", 50, 17, (64, 64), 1)
    packed = pack_state_dict(sd, p.infos, p.weight_bytes)
    hm, _ = emulate(p.describe(), p.infos, packed, x, exact=True)
    assert (hm - ref).abs().max() < 5e-4 * max(1.0, ref.abs().max().item())


def test_workspace_tensors_do_not_overlap_within_a_module():
    """Tensors that are live at the same time must not alias: check all tensors of one arena generation."""
    p = Plan("hrnet", 48, 17, (384, 288), 4)
    d = p.describe()
    assert d["act_bytes"] == p.act_bytes
    last_write = {}
    for i, op in enumerate(d["ops"]):
        if op["out"] < 0 or d["tensors"][op["out"]]["offset"] < 0:
            continue
        last_write[op["out"]] = i
    # readers of each tensor
    readers = {}
    for i, op in enumerate(d["ops"]):
        for t in [op["in"], op["res"]] + op["src"][: op["nsrc"]]:
            if t >= 0:
                readers.setdefault(t, []).append(i)
    def span(t):
        tt = d["tensors"][t]
        nbytes = 4 * tt["C"] * tt["H"] * tt["W"] * (4 if tt["f32"] else 2)
        return tt["offset"], tt["offset"] + nbytes
    firsts = {}
    for i, op in enumerate(d["ops"]):
        if op["out"] >= 0 and op["out"] not in firsts:
            firsts[op["out"]] = i
    ids = [t for t in firsts if d["tensors"][t]["offset"] >= 0]
    for a in ids:
        for b in ids:
            if a >= b:
                continue
            (a0, a1), (b0, b1) = span(a), span(b)
            if a0 < b1 and b0 < a1:   # aliasing in memory: lifetimes must be disjoint in op order
                a_live = (firsts[a], max(readers.get(a, [last_write[a]]) + [last_write[a]]))
                b_live = (firsts[b], max(readers.get(b, [last_write[b]]) + [last_write[b]]))
                assert a_live[1] < b_live[0] or b_live[1] < a_live[0], (a, b, a_live, b_live)


@pytest.mark.parametrize("arch,c,res", [("hrnet", 48, (384, 288)), ("hrnet", 32, (256, 192)), ("poseresnet", 50, (256, 192))])
def test_two_issuer_pipelines_keep_their_ring_depth(arch, c, res):
    """mma_warps == 2 splits the slot / stage ring between two independent issuer pipelines: every ring must keep two
    buffers, CTA-pair kernels keep one issuer, and wide / deep-K im2col tiles (where the halved ring costs more than the
    overlap gains, profiles/r01_exp_mma2_issuers.log) stay on one issuer."""
    d = Plan(arch, c, 17, res, 8).describe()
    n2 = 0
    for op in d["ops"]:
        if not op.get("use_tc"):
            continue
        t = op["tc"]
        if t["mma_warps"] == 2:
            n2 += 1
            if op["use_patch"]:
                assert t["patch_slots"] >= 4 and t["patch_cs"] == 1, op["name"]
            else:
                assert t["stages"] >= 4 and t["n_tile"] <= 96 and t["cs"] == 1, op["name"]
        else:
            assert t["mma_warps"] == 1, op["name"]
    assert n2 > 0


def test_chain_and_exchange_unit_planning():
    """Host-side planning of round 2 (no GPU): 26 branch chains of 8 convs, their SM shares (per mille, each module's sum
    to 1000, calibrated 375 / 235 / 195 / 195 in a 4-branch module), the three-slot layout of the 64 + 32-channel chains,
    exchange units per source branch, launch counts of the variants, CTA pairs only on the im2col chains."""
    from simple_hrnet_b200 import _lib
    from simple_hrnet_b200.engine import Plan
    p = Plan("hrnet", 48, 17, (384, 288), 64)
    d = p.describe()
    chains = d["chains"]
    assert len(chains) == 26 and all(len(c["ops"]) == 8 for c in chains)
    by_mod = {}
    for c in chains:
        by_mod.setdefault(c["module"], []).append(c)
    for m, cs in by_mod.items():
        assert abs(sum(c["share_permille"] for c in cs) - 1000) <= len(cs), (m, [c["share_permille"] for c in cs])
    four = [c["share_permille"] for c in sorted(by_mod[max(by_mod)], key=lambda c: c["branch"])]
    assert four == [375, 235, 195, 195]
    assert [c["tail64"] for c in sorted(by_mod[max(by_mod)], key=lambda c: c["branch"])] == [0, 1, 0, 0]      # C = 96 only
    assert all(c["patch"] == (1 if c["branch"] < 2 else 0) for c in chains) and all(c["pair"] == 0 for c in chains)
    assert all(c["smem"] <= 227 * 1024 for c in chains)
    assert sorted(len(x["ops"]) for x in d["xunits"]) == sorted([3, 2, 2] * 4 + [6, 4, 3, 3] * 2)
    assert p.launch_count == 317 - 26 * 7 - (60 - 20)
    assert Plan("hrnet", 48, 17, (384, 288), 64, tune={_lib.TUNE_XUNIT: 3}).launch_count == 317 - 26 * 7
    assert Plan("hrnet", 48, 17, (384, 288), 64, flags=_lib.FLAG_NO_CHAIN, tune={_lib.TUNE_XUNIT: 3}).launch_count == 317
    pp = Plan("hrnet", 48, 17, (384, 288), 64, tune={_lib.TUNE_CHAIN_PAIR: 2}).describe()["chains"]
    assert all(c["pair"] == (0 if c["patch"] else 1) for c in pp) and all(c["stages"] == 8 for c in pp if c["pair"])
    two = Plan("hrnet", 48, 17, (384, 288), 64, tune={_lib.TUNE_CHAIN_SKIP: 8}).describe()["chains"]
    assert all(c["tail64"] == 0 for c in two)
    w32 = Plan("hrnet", 32, 17, (256, 192), 32).describe()["chains"]
    assert len(w32) == 26 and all(c["tail64"] == 0 for c in w32)              # C = 32 / 64: one chunk
