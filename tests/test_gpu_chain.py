"""`-m gpu`: branch chains (conv_chain.cu) and exchange units (conv_xunit.cu) against the per-conv launches of the same plan.

A chain runs the eight 3x3 convs of a StageModule branch (reference models_/modules.py:56-72 x 4, models_/hrnet.py:15-20)
as ONE persistent kernel with tile-level dependencies.  Inside a tile the MMA order and the epilogue arithmetic are
those of the per-conv kernels, so the two paths must agree BIT FOR BIT (HRNET_FLAG_NO_CHAIN selects the per-conv path,
which tests/test_gpu_forward.py pins against the oracle and the reference fixtures).  A dependency bug (a tile started
before its halo was stored, a stale L1 line, a missing fence) shows up here as a mismatch or a run-to-run difference."""
import numpy as np
import pytest
import torch

from oracle import hrnet_oracle as O
from simple_hrnet_b200 import B200Engine, HostPipeline, _lib

pytestmark = pytest.mark.gpu


def _engine(c, res, maxb, sd, flags=0, tune=None):
    e = B200Engine("hrnet", c, 17, res, maxb, torch.device("cuda:0"), flags=flags, tune=tune)
    e.load_state_dict(sd)
    return e


PER_CONV = _lib.FLAG_NO_CHAIN      # with NO_XUNIT: every conv its own launch (round 1's path)
NO_XUNIT = {_lib.TUNE_XUNIT: 3}
XUNIT_CONVS, XUNIT_ALL = {_lib.TUNE_XUNIT: 1}, {_lib.TUNE_XUNIT: 2}   # exchange units per source branch (default) / per module + sums
PAIR = {_lib.TUNE_CHAIN_PAIR: 2}       # im2col chains on CTA pairs (cta_group::2)
TWO_SLOTS = {_lib.TUNE_CHAIN_SKIP: 8}  # C = 96 halo-patch chains with two equal patch slots instead of the 2 + 1 narrow layout


def test_chain_plan_is_active():
    e = _engine(32, (64, 64), 2, O.make_state_dict(O.hrnet_param_spec(32, 17), seed=3, bn="random"))
    d = e.describe()
    assert len(d["chains"]) == 26 and all(len(ch["ops"]) == 8 for ch in d["chains"])
    # exchange units per source branch: stage 3 modules 3 + 2 + 2 convs, stage 4 modules 6 + 4 + 3 + 3 (single-conv groups stay launches)
    assert sorted(len(x["ops"]) for x in d["xunits"]) == sorted([3, 2, 2] * 4 + [6, 4, 3, 3] * 2)
    assert e.launch_count == 317 - 26 * 7 - (60 - 20)
    sd = O.make_state_dict(O.hrnet_param_spec(32, 17), seed=3, bn="random")
    assert _engine(32, (64, 64), 2, sd, PER_CONV, NO_XUNIT).launch_count == 317
    assert _engine(32, (64, 64), 2, sd, 0, NO_XUNIT).launch_count == 317 - 26 * 7
    xa = _engine(32, (64, 64), 2, sd, 0, XUNIT_ALL)
    assert sorted(len(x["ops"]) for x in xa.describe()["xunits"]) == [2, 3, 7, 7, 7, 7, 16, 16]
    assert xa.launch_count == 317 - 26 * 7 - (65 - 8) - 23    # sums as tickets


@pytest.mark.parametrize("c,res,n", [(32, (64, 64), 2), (32, (128, 96), 5), (32, (256, 192), 32), (48, (384, 288), 16), (48, (256, 192), 9)])
def test_chain_equals_per_conv_launches(c, res, n):
    sd = O.make_state_dict(O.hrnet_param_spec(c, 17), seed=5, bn="random")
    x = torch.randn(n, 3, *res, generator=torch.Generator().manual_seed(2)).cuda()
    ref = _engine(c, res, n, sd, PER_CONV, NO_XUNIT)
    jr, ir, hr = ref.forward_decode(x, return_heatmaps=True)
    for flags, tune in ((0, None), (_lib.FLAG_NO_GRAPH, None), (_lib.FLAG_SERIAL, None), (0, PAIR), (_lib.FLAG_SERIAL, PAIR),
                        (0, NO_XUNIT), (_lib.FLAG_NO_CHAIN, None), (0, XUNIT_ALL), (_lib.FLAG_NO_CHAIN, XUNIT_ALL)):
        e = _engine(c, res, n, sd, flags, tune)
        for rep in range(3):                          # graph replays / epoch stamps of consecutive launches
            j, i, h = e.forward_decode(x, return_heatmaps=True)
            assert torch.equal(h, hr), f"flags={flags} rep={rep}: max diff {(h - hr).abs().max().item():.3e}"
            assert torch.equal(j, jr) and torch.equal(i, ir)
        for k in (1, n // 2, n):                      # any batch size may follow any other (flags are epoch stamps)
            if k >= 1:
                assert torch.equal(e.forward_decode(x[:k], return_heatmaps=True)[2], hr[:k]), f"flags={flags} n={k}"


def test_chain_full_size_headline_and_grid_splits():
    """W48 384x288, 64 crops (BASELINE config 3): chains == per-conv launches bit for bit, for the default grid split
    and for skewed ones (few CTAs per chain: long dependency waits; one CTA: strictly sequential tickets)."""
    sd = O.make_state_dict(O.hrnet_param_spec(48, 17), seed=0, bn="default")
    x = torch.randn(64, 3, 384, 288, generator=torch.Generator().manual_seed(3)).cuda()
    hr = _engine(48, (384, 288), 64, sd, PER_CONV, NO_XUNIT)(x)
    for tune in (None, {_lib.TUNE_CHAIN_SHARE0: 100, 1: 100, 2: 400, 3: 400}, {_lib.TUNE_CHAIN_GRID_CAP: 3}, {_lib.TUNE_CHAIN_M2: 2},
                 PAIR, {**PAIR, _lib.TUNE_CHAIN_GRID_CAP: 5}, {**PAIR, _lib.TUNE_CHAIN_SHARE0: 100, 1: 100, 2: 400, 3: 400},
                 NO_XUNIT, XUNIT_ALL, TWO_SLOTS):
        e = _engine(48, (384, 288), 64, sd, 0, tune)
        for rep in range(2):
            h = e(x)
            assert torch.equal(h, hr), f"tune={tune} rep={rep}: max diff {(h - hr).abs().max().item():.3e}"
        assert torch.equal(e(x[:33]), hr[:33])


def test_two_forwards_share_the_gpu():
    """HostPipeline keeps two forwards in flight on two streams: the chain kernels of both compete for the SMs (partial
    residency of every grid) -- dynamic tickets keep that deadlock-free and the results identical."""
    sd = O.make_state_dict(O.hrnet_param_spec(32, 17), seed=4, bn="random")
    e = _engine(32, (256, 192), 16, sd)
    g = torch.Generator().manual_seed(11)
    batches = [torch.randint(0, 256, (16, 256, 192, 3), generator=g, dtype=torch.uint8).pin_memory() for _ in range(6)]
    want = [e.forward_host_u8(b.numpy())[0] for b in batches]
    got = list(HostPipeline(e, depth=2).run(batches))
    assert all(np.array_equal(a, b) for a, b in zip(got, want))
