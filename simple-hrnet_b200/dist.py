"""Batch-sharded multi-GPU inference: one process per GPU (torchrun), weights resident per GPU,
crops split contiguously by batch, ONE all-gather of the decoded joints [B,J,3] over NCCL/NVLink.
Replaces the reference's single-process nn.DataParallel scatter/replicate/gather
(SimpleHRNet.py:123-135).  The decode kernel writes straight into this rank's slice of the gather
buffer, so the collective follows the kernel with no staging copy."""
import torch
import torch.distributed as dist


def shard_range(total, world_size, rank):
    """Contiguous, balanced split of `total` items: the first (total % world) ranks get one extra."""
    base, extra = divmod(total, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


class ShardedPredictor:
    """predict_joints(images) -> joints [B,J,3] for the GLOBAL batch on every rank.

    `local_fn(images_local, joints_out)` runs the engine on this rank's shard and writes
    [n_local,J,3] into `joints_out` (a view into the gather buffer).  With equal shards the
    collective is a single in-place all_gather_into_tensor; ragged shards gather max-size padded slots
    (still one collective) and compact."""

    def __init__(self, local_fn, nof_joints, device, group=None):
        self.local_fn = local_fn
        self.J = nof_joints
        self.device = torch.device(device)
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0

    def predict_joints(self, images_global_count, images_local):
        B = int(images_global_count)
        lo, hi = shard_range(B, self.world, self.rank)
        assert images_local.shape[0] == hi - lo, "local shard does not match shard_range"
        buf = torch.empty(B, self.J, 3, dtype=torch.float32, device=self.device)
        if hi > lo:
            self.local_fn(images_local, buf[lo:hi])
        if self.world == 1:
            return buf
        if B % self.world == 0:
            dist.all_gather_into_tensor(buf, buf[lo:hi], group=self.group)
        else:
            # ragged shards: one all-gather of max-size padded slots, then compact
            mx = -(-B // self.world)
            slots = torch.zeros(self.world * mx, self.J, 3, dtype=torch.float32, device=self.device)
            mine = torch.zeros(mx, self.J, 3, dtype=torch.float32, device=self.device)
            mine[: hi - lo] = buf[lo:hi]
            dist.all_gather_into_tensor(slots, mine, group=self.group)
            for r in range(self.world):
                a, b = shard_range(B, self.world, r)
                buf[a:b] = slots[r * mx: r * mx + (b - a)]
        return buf
