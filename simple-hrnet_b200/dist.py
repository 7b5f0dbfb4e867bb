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

    # ---- pipelined form: the collective of batch i overlaps the forward of batch i+1 -----------------------------------
    def submit(self, images_global_count, images_local):
        """Enqueue one global batch and return a handle; `result(handle)` gives joints [B,J,3] on every rank.

        The forward + decode of the local shard runs on the caller's current stream; the all-gather is issued on a side
        stream that only waits for THAT forward (one CUDA event), into one of two rotating gather buffers.  The next
        `submit` can therefore start its forward immediately: the per-step NCCL launch latency and the rank skew a
        collective exposes (the ranks never finish a forward at the same instant) hide behind compute instead of sitting
        between two forwards.  At most two batches may be in flight; equal shards only (B % world == 0)."""
        B = int(images_global_count)
        if self.world > 1 and B % self.world:
            raise ValueError("pipelined gather needs equal shards (B % world_size == 0); use predict_joints")
        lo, hi = shard_range(B, self.world, self.rank)
        assert images_local.shape[0] == hi - lo, "local shard does not match shard_range"
        if not hasattr(self, "_slots"):
            self._slots, self._turn = [None, None], 0
            self._comm = torch.cuda.Stream(self.device) if self.device.type == "cuda" else None
        slot = self._turn
        self._turn ^= 1
        st = self._slots[slot]
        if st is None or st["buf"].shape[0] != B:
            st = {"buf": torch.empty(B, self.J, 3, dtype=torch.float32, device=self.device), "done": None}
            self._slots[slot] = st
        if self._comm is not None and st["done"] is not None:
            torch.cuda.current_stream(self.device).wait_event(st["done"])   # the buffer's previous gather has finished
        buf = st["buf"]
        if hi > lo:
            self.local_fn(images_local, buf[lo:hi])
        if self.world > 1:
            if self._comm is not None:
                ready = torch.cuda.Event()
                ready.record(torch.cuda.current_stream(self.device))
                with torch.cuda.stream(self._comm):
                    self._comm.wait_event(ready)
                    dist.all_gather_into_tensor(buf, buf[lo:hi], group=self.group)
                    st["done"] = torch.cuda.Event()
                    st["done"].record(self._comm)
            else:
                dist.all_gather_into_tensor(buf, buf[lo:hi], group=self.group)
        return slot

    def result(self, handle):
        """Joints [B,J,3] of the batch submitted under `handle`; the caller's stream waits for its gather."""
        st = self._slots[handle]
        if self._comm is not None and st["done"] is not None:
            torch.cuda.current_stream(self.device).wait_event(st["done"])
        return st["buf"]
