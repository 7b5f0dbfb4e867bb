"""World-size-2 gloo tests (CPU) of the batch-sharding + joints all-gather logic that replaces the reference's
nn.DataParallel (SimpleHRNet.py:123-135).  The per-rank engine is substituted by the CPU oracle decode so the
collective plumbing (shard ranges, in-place gather buffer, ragged shards) is exercised without a GPU."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import hrnet_oracle as O
from simple_hrnet_b200.dist import ShardedPredictor, shard_range


def test_shard_range_partitions_exactly():
    for total in (0, 1, 7, 64, 512, 513):
        for world in (1, 2, 3, 8):
            spans = [shard_range(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, B, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(0)
        hm = rng.standard_normal((B, 17, 16, 12)).astype(np.float32)    # "heat-maps" of the global batch
        boxes = np.repeat(np.asarray([[0, 0, 48, 64]], np.float32), B, 0)
        lo, hi = shard_range(B, world, rank)

        def local_fn(images_local, joints_out):
            pts, _ = O.decode_joints(images_local.numpy(), boxes[lo:hi])
            joints_out.copy_(torch.from_numpy(pts))

        pred = ShardedPredictor(local_fn, 17, torch.device("cpu"))
        joints = pred.predict_joints(B, torch.from_numpy(hm[lo:hi]))
        ref, _ = O.decode_joints(hm, boxes)
        assert np.array_equal(joints.numpy(), ref), f"rank {rank}: gathered joints differ"
        if B % world == 0:   # pipelined form: two batches in flight, results in submission order
            hm2 = rng.standard_normal((B, 17, 16, 12)).astype(np.float32)
            ref2, _ = O.decode_joints(hm2, boxes)
            h1 = pred.submit(B, torch.from_numpy(hm[lo:hi]))
            h2 = pred.submit(B, torch.from_numpy(hm2[lo:hi]))
            assert np.array_equal(pred.result(h1).numpy(), ref) and np.array_equal(pred.result(h2).numpy(), ref2)
            h3 = pred.submit(B, torch.from_numpy(hm[lo:hi]))           # reuses the first buffer
            assert np.array_equal(pred.result(h3).numpy(), ref)
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("B", [8, 7])   # equal shards (single in-place all_gather_into_tensor) and ragged shards
def test_sharded_predictor_gloo_world2(tmp_path, B):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, B, str(tmp_path)), nprocs=2, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(2))
