#!/usr/bin/env python
"""bench.py -- persons/s of the simple-HRNet hot path (HRNet forward + argmax decode) on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  (N > 1: launched by torchrun, one rank per GPU; crops sharded by batch, one NCCL all-gather of joints)

One "step" = one pass of the hot path over one batch of synthetic crops: HRNet-W48 @ 384x288, 17 joints,
64 persons per GPU (BASELINE.json configs[2], the config the metric is quoted on), random-init weights
(oracle.make_state_dict, seed 0), inputs torch.randn at the model boundary.
  value : whole-job persons/s, inputs resident in HBM, CUDA-event timed, max over ranks
  e2e   : same metric from pinned HOST uint8 crops to HOST joints through HostPipeline (double-buffered
          hrnet_forward_host_u8_async): every step's H2D and D2H copies are inside the timed region; the blocking
          one-call-per-step figure is reported next to it (sync_call_value)
  roofline     : the dominant kernel (stage-4 3x3 implicit-GEMM branch convs, tcgen05) timed alone
                 with CUDA events, algorithmic FLOPs / time vs the measured dense bf16 peak
  cpu_baseline : the oracle (CPU port of the reference's PyTorch path) on this box's host cores
`--impl reference` times that CPU implementation as the reference arm (same metric / config).
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

C, J, H, W = 48, 17, 384, 288
PER_GPU_BATCH = 64
GFLOP_PER_PERSON = 70.613          # 2*MACs of all 293 convs, W48 @ 384x288
STAGE4_BRANCH = [                   # (C, H, W) of the four stage-4 branches; every 3x3 conv = 0.1433 GMAC/person
    (48, 96, 72), (96, 48, 36), (192, 24, 18), (384, 12, 9)]
METRIC = "persons/sec HRNet-W48 384x288 (forward + argmax decode)"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(tflops=d.get("bf16_tflops", 1590.0), tflops_sustained=d.get("bf16_tflops_sustained", 1400.0),
                    hbm=d.get("hbm_gbs", 6650.0), source="measured (MEASURED_PEAKS.json)")
    return dict(tflops=1590.0, tflops_sustained=1400.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


def load_traffic():
    """ncu dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernels, extracted from the committed
    `ncu --set full` capture (profiles/r01_traffic.json, written by tools/ncu_traffic.py); None if no capture exists."""
    p = os.path.join(ROOT, "profiles", "r01_traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p))
        except Exception:
            return None
    return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = float(r[1])
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def make_weights():
    from oracle import hrnet_oracle as O
    return O.make_state_dict(O.hrnet_param_spec(C, J), seed=0, bn="default")


def best_cpu_threads(sd):
    """The CPU arm gets its best thread count: oneDNN is often slower with every hardware thread of a big host
    (measured 0.45-1.3 persons/s at 128 threads vs several at 16-32).  Quick calibration on a 4-person forward."""
    from oracle import hrnet_oracle as O
    ncpu = os.cpu_count() or 1
    # capped at 64: a 4-person forward on all 128 threads of the round-1 GPU host took 123 s (oversubscribed oneDNN)
    cands = sorted({min(ncpu, 64), min(ncpu, 32), min(ncpu, 16)}, reverse=True)
    if len(cands) == 1:
        return cands[0], {}
    x = torch.randn(4, 3, H, W, generator=torch.Generator().manual_seed(0))
    res = {}
    for t in cands:
        torch.set_num_threads(t)
        O.hrnet_forward(sd, x[:1])
        t0 = time.perf_counter()
        O.hrnet_forward(sd, x)
        res[t] = time.perf_counter() - t0
    return min(res, key=res.get), {k: round(v, 2) for k, v in res.items()}


def cpu_reference_persons_per_s(sd, n_persons, reps, threads):
    """The reference's CPU path (oracle port: same aten ops as models_/hrnet.py + the numpy decode)."""
    from oracle import hrnet_oracle as O
    torch.set_num_threads(threads)
    x = torch.randn(n_persons, 3, H, W, generator=torch.Generator().manual_seed(0))
    boxes = np.repeat(np.asarray([[0, 0, W, H]], dtype=np.float32), n_persons, axis=0)
    O.decode_joints(O.hrnet_forward(sd, x[:2]).numpy(), boxes[:2])          # warm-up
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        hm = O.hrnet_forward(sd, x).numpy()
        O.decode_joints(hm, boxes)
        ts.append(time.perf_counter() - t0)
    return n_persons / float(np.median(ts)), ts


def run_reference(args, rank):
    if rank != 0:
        return
    sd = make_weights()
    threads, calib = best_cpu_threads(sd)
    n = 16
    cpu_reference_persons_per_s(sd, 2, 1, threads)
    ts = []
    torch.set_num_threads(threads)
    from oracle import hrnet_oracle as O
    x = torch.randn(n, 3, H, W, generator=torch.Generator().manual_seed(0))
    boxes = np.repeat(np.asarray([[0, 0, W, H]], dtype=np.float32), n, axis=0)
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        O.decode_joints(O.hrnet_forward(sd, x).numpy(), boxes)
        if i >= args.warmup:
            ts.append(time.perf_counter() - t0)
    total = float(np.sum(ts))
    v = n * len(ts) / total
    out = {"impl": "reference", "metric": METRIC, "value": v, "unit": "persons/s", "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(ts),
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": "HRNet-W48 384x288, 17 joints, forward + argmax decode, CPU fp32 (oracle port of "
                                  "the reference PyTorch path), bounded sample of 16 persons per step",
                      "per_gpu_batch": PER_GPU_BATCH},
           "cpu_baseline": {"value": v, "unit": "persons/s", "cores": threads, "kind": "port",
                            "sample": f"{n} persons x {len(ts)} steps, torch {torch.__version__} CPU, {threads} of {os.cpu_count()} "
                                      f"threads (fastest of a 4-person calibration {calib})"},
           "e2e": {"value": v, "unit": "persons/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-kernel-roofline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch.distributed as dist
    from simple_hrnet_b200 import B200Engine, HostPipeline, ShardedPredictor, shard_range, _lib
    from tests import gpu_util as G

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    peaks = load_peaks()

    sd = make_weights()
    eng = B200Engine("hrnet", C, J, (H, W), PER_GPU_BATCH, dev)
    eng.load_state_dict(sd)
    B = PER_GPU_BATCH * world
    lo, hi = shard_range(B, world, rank)
    # three rotating input buffers (85 MB each at 64 crops); the ~2.7 GB of activations a step streams through
    # HBM evicts the 126 MB L2 many times over between two reads of the same input
    gen = torch.Generator().manual_seed(1000 + rank)
    xs = [torch.randn(hi - lo, 3, H, W, generator=gen).to(dev) for _ in range(3)]
    pred = ShardedPredictor(lambda imgs, out: eng.forward_decode(imgs, joints_out=out), J, dev)

    def step(i):
        return pred.predict_joints(B, xs[i % 3])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step(i)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        joints = step(i)
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    clocks = sampler.stop() if rank == 0 else None
    value = B * args.steps / (ms_total / 1e3)

    # ---- e2e: host (pinned) buffers through hrnet_forward_host, H2D + D2H inside the timed region
    # the reference-facing call: uint8 BGR crops at network resolution in (pinned) host memory -> joints in host memory
    xh = [torch.randint(0, 256, (hi - lo, H, W, 3), generator=gen, dtype=torch.uint8).pin_memory() for _ in range(2)]
    xh_np = [t.numpy() for t in xh]
    for i in range(3):
        eng.forward_host_u8(xh_np[i % 2])
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(3, args.steps // 2)
    for i in range(e2e_steps):
        jh, ih, _ = eng.forward_host_u8(xh_np[i % 2])
        if world > 1:   # the gathered result is what a multi-GPU caller receives
            buf = torch.empty(B, J, 3, device=dev)
            buf[lo:hi].copy_(torch.from_numpy(jh), non_blocking=True)
            dist.all_gather_into_tensor(buf, buf[lo:hi])
            buf.cpu()
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_sync_value = B * e2e_steps / float(e2e_s.item())
    # the serving loop a throughput user runs: HostPipeline double-buffers two engines that share the device weights, so
    # batch i+1's H2D copy and batch i-1's D2H copy overlap batch i's forward; every step's copies are inside the timed region
    pipe = HostPipeline(eng, depth=2)

    def consume(jh):
        if world > 1:
            buf = torch.empty(B, J, 3, device=dev)
            buf[lo:hi].copy_(torch.from_numpy(jh), non_blocking=True)
            dist.all_gather_into_tensor(buf, buf[lo:hi])
            buf.cpu()
    for jh in pipe.run(xh[i % 2] for i in range(4)):
        consume(jh)
    barrier()
    t0 = time.perf_counter()
    for jh in pipe.run(xh[i % 2] for i in range(e2e_steps)):
        consume(jh)
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = B * e2e_steps / float(e2e_s.item())
    h2d = (hi - lo) * 3 * H * W          # uint8 images; normalisation runs in the stem kernel
    d2h = (hi - lo) * J * 3 * 4 + (hi - lo) * J * 4

    # ---- dominant kernel: the stage-4 3x3 implicit-GEMM branch convs, timed live inside the network with one CUDA
    # event pair per kernel (hrnet_profile_ops: serial pass over the plan's ops, inputs = real activations)
    roofline, breakdown = None, None
    if rank == 0 and not args.no_kernel_roofline:
        ops, desc = eng.profile_ops(xs[0], iters=5)
        n = PER_GPU_BATCH
        cls_t, cls_f, cls_n = {}, {}, {}
        per_branch = {}
        for name, op, us in ops:
            flop = 0.0
            if op["kind"] == 1:
                tin = desc["tensors"][op["in"]]
                flop = 2.0 * n * (tin["H"] // op["stride"]) * (tin["W"] // op["stride"]) * op["k"] ** 2 * op["cin"] * op["cout"]
                if ".branches." in name:
                    cls = name.split(".")[0] + " 3x3 branch convs"
                    if name.startswith("stage4."):
                        key = (op["cin"], f"{tin['H']}x{tin['W']}", "patch" if op["use_patch"] else "im2col")
                        pb = per_branch.setdefault(key, [0.0, 0.0, 0])
                        pb[0] += us; pb[1] += flop; pb[2] += 1
                elif "fuse_layers" in name:
                    cls = "exchange-unit convs (1x1 up, 3x3 s2 down)"
                elif name.startswith("layer1"):
                    cls = "layer1 bottlenecks"
                else:
                    cls = "stem conv2 + transitions"
            else:
                cls = {0: "stem conv1 (hi/lo-split tcgen05)", 2: "exchange-unit sum", 3: "head 1x1", 4: "argmax decode"}[op["kind"]]
            cls_t[cls] = cls_t.get(cls, 0.0) + us
            cls_f[cls] = cls_f.get(cls, 0.0) + flop
            cls_n[cls] = cls_n.get(cls, 0) + 1
        tot = sum(cls_t.values())
        breakdown = {"serial_total_us": round(tot, 1), "classes": {
            k: {"us": round(v, 1), "share": round(v / tot, 4), "launches": cls_n[k],
                "tflops": round(cls_f[k] / v / 1e6, 1) if cls_f[k] else None} for k, v in sorted(cls_t.items(), key=lambda kv: -kv[1])}}
        k4 = "stage4 3x3 branch convs"
        ach = cls_f[k4] / cls_t[k4] / 1e6
        roofline = {"kernel": "stage-4 3x3 s1 branch convs: 96 launches per forward (conv3x3_patch_tc_kernel for C=48/96, "
                              "conv_igemm_tc_kernel for C=192/384), N=64, each timed with a CUDA event pair inside a serial "
                              "pass of the whole network; achieved = algorithmic conv FLOPs / summed kernel time",
                    "bound": "tensor", "achieved": round(ach, 1), "peak": peaks["tflops_sustained"], "unit": "TFLOP/s",
                    "frac": round(ach / peaks["tflops_sustained"], 4),
                    "frac_of_burst_peak": round(ach / peaks["tflops"], 4), "traffic": load_traffic(),
                    "peak_source": peaks["source"] + ", sustained (kernels timed inside a long step)",
                    "launches": cls_n[k4], "us_total": round(cls_t[k4], 1),
                    "per_branch": [{"C": k[0], "map": k[1], "kernel": k[2], "launches": v[2], "us_avg": round(v[0] / v[2], 2),
                                    "tflops": round(v[1] / v[0] / 1e6, 1), "frac_of_burst_peak": round(v[1] / v[0] / 1e6 / peaks["tflops"], 4)}
                                   for k, v in sorted(per_branch.items())]}

    # ---- CPU baseline on this box's host cores (rank 0, N=1 only), bounded sample
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads, calib = best_cpu_threads(sd)
        v, ts = cpu_reference_persons_per_s(sd, 16, 2, threads)
        cpu = {"value": round(v, 3), "unit": "persons/s", "cores": threads, "kind": "port",
               "sample": f"16 persons x 2 reps (after warm-up), oracle CPU fp32 forward + numpy decode, {threads} of "
                         f"{os.cpu_count()} threads (fastest of a 4-person calibration {calib}), rep times "
                         f"{[round(t, 2) for t in ts]} s"}

    if rank == 0:
        step_tflops = GFLOP_PER_PERSON * value / 1e3
        out = {"metric": METRIC, "value": round(value, 2), "unit": "persons/s", "n_gpus": world, "steps": args.steps,
               "warmup": args.warmup, "ms_per_step": round(ms_total / args.steps, 4), "higher_is_better": True,
               "scaling": "weak", "vs_baseline": None, "dtype": "f16", "data": "synthetic",
               "config": {"workload": "HRNet-W48 384x288, 17 joints, 64 persons per GPU per step, fp16 NHWC activations "
                                      "with fp32 accumulate, random-init weights (seed 0), torch.randn inputs",
                          "global_batch": B, "per_gpu_batch": PER_GPU_BATCH, "parallelism": f"dp{world}",
                          "l2": "3 rotating input buffers; a step streams ~2.7 GB of activations through HBM (>> 126 MB L2)",
                          "collective": "one NCCL all_gather_into_tensor of joints [B,17,3] f32 per step" if world > 1 else "none"},
               "clocks": clocks,
               "e2e": {"value": round(e2e_value, 2), "unit": "persons/s", "h2d_bytes_per_step": h2d,
                       "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                       "call": "HostPipeline.run (hrnet_forward_host_u8_async on 2 engines sharing the weights): pinned uint8 BGR crops "
                               "[n,384,288,3] -> H2D -> forward + decode -> D2H joints, batch i+1's copy overlaps batch i's forward",
                       "sync_call_value": round(e2e_sync_value, 2),
                       "sync_call": "hrnet_forward_host_u8, one blocking call per step (no overlap)"},
               "gpu_launches": eng.launch_count * args.steps,
               "step_flops": {"tflops": round(step_tflops, 1), "per_gpu_tflops": round(step_tflops / world, 1),
                              "frac_of_sustained_peak": round(step_tflops / world / peaks["tflops_sustained"], 4),
                              "gflop_per_person": GFLOP_PER_PERSON},
               "roofline": roofline, "layer_breakdown": breakdown, "cpu_baseline": cpu}
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
