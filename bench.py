#!/usr/bin/env python
"""
bench.py — Segmentor.fit throughput on B200 (BASELINE.json metric:
"Segmentor.fit images/sec (512x512, 3-class)"; workload = configs[1]: Unet nb_classes=3,
batch 32 of 512x512 fp32 synthetic stacks per GPU).

  python bench.py --gpus N --steps K --warmup W           (N > 1: launched by torchrun)
  python bench.py --impl reference ...                    (the reference algorithm on host cores)

A "step" is one Segmentor.fit training cycle of the reference (atomai/trainers/trainer.py:233-251):
one train mini-batch (forward + backward + Adam) followed by one test mini-batch forward.
Prints ONE JSON line (rank 0).  Multi-GPU runs are data-parallel with weak scaling: every rank
trains on its own shard of a global batch of 32*N images.
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

H = W = 512
BATCH = 32
NB_CLASSES = 3
FLOP_PER_PIXEL_TRAIN = 196_992        # fwd 65,664 x 3 (fwd + dgrad + wgrad)
FLOP_PER_PIXEL_FWD = 65_664


def host_cores():
    """Threads the CPU arm may really use: scheduler affinity, capped by the cgroup CPU quota
    (os.cpu_count() reports the machine, not the container: oversubscribing it is 10x slower)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        q, per = open("/sys/fs/cgroup/cpu.max").read().split()
        if q != "max":
            n = max(1, min(n, int(float(q) / float(per) + 0.5)))
    except Exception:  # noqa
        pass
    return n


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], bf16=d["bf16_tflops"], bf16_sustained=d["bf16_tflops_sustained"],
                    src="measured")
    return dict(hbm=6650.0, bf16=1590.0, bf16_sustained=1400.0, src="fallback")


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                out = subprocess.check_output(
                    ["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                     "--format=csv,noheader,nounits"], encoding="utf-8", timeout=5)
                self.rows.append([c.strip() for c in out.strip().split(",")])
            except Exception:  # noqa
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows)
        reasons = []
        for i, name in enumerate(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                  "sw_power_cap"]):
            if any(r[3 + i].lower().startswith("active") for r in self.rows):
                reasons.append(name)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]),
                "power_w_max": max(float(r[2]) for r in self.rows), "samples": len(self.rows),
                "reasons": reasons}


def synth(n, seed):
    """Synthetic microscopy-like stack: uniform [0,1) images, labels in {0,1,2} (all present)."""
    rs = np.random.RandomState(seed)
    X = rs.rand(n, 1, H, W).astype(np.float32)
    y = rs.randint(0, NB_CLASSES, (n, H, W)).astype(np.int64)
    y[:, 0, :NB_CLASSES] = np.arange(NB_CLASSES)
    return X, y


# ------------------------------------------------------------------------------ reference arm
def run_reference(args):
    """The reference's algorithm for this path on the host cores: oracle/nets_ref.py (plain torch
    CPU fp32 restatement, pinned against the unmodified reference's goldens).  /root/reference is
    Python and cannot travel to the GPU box, so this port is the reference arm (kind "port").
    Each step is a bounded sample of the workload: a batch of `sample_batch` 512x512 images."""
    from collections import OrderedDict
    from oracle import nets_ref
    from atomai_b200.nets import Unet
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_cores()
    torch.set_num_threads(cores)
    sample_batch = args.ref_batch
    torch.manual_seed(1)
    net = Unet(nb_classes=NB_CLASSES)
    sd = OrderedDict((k, v.detach().clone()) for k, v in net.state_dict().items())
    cfg = dict(nb_classes=NB_CLASSES)
    X, y = synth(2 * sample_batch, 1)
    X, y = torch.from_numpy(X), torch.from_numpy(y)
    state = {}

    def step(i):
        nets_ref.unet_train_step(X[:sample_batch], y[:sample_batch], sd, cfg, state)
        with torch.no_grad():
            lt = nets_ref.unet_forward(X[sample_batch:], sd, cfg, training=False)
            nets_ref.seg_loss(lt, y[sample_batch:], NB_CLASSES)

    for i in range(args.warmup):
        step(i)
    t0 = time.perf_counter()
    for i in range(args.steps):
        step(i)
    dt = time.perf_counter() - t0
    ips = sample_batch * args.steps / dt
    line = {
        "impl": "reference", "metric": "Segmentor.fit images/sec (512x512, 3-class)",
        "value": ips, "unit": "images/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"Unet nb_classes=3, {H}x{W}x1 fp32, train step + test forward per "
                               f"cycle; bounded sample: batch {sample_batch} instead of {BATCH}"},
        "cpu_baseline": {"value": ips, "unit": "images/s", "cores": cores, "kind": "port",
                         "sample": f"{args.steps} cycles of batch {sample_batch} (train step + "
                                   f"test forward), torch CPU fp32, {cores} threads"},
        "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------ our arm
def time_layer(ops, Source, N, hh, ww, cin, cout, reps=10):
    """Live CUDA-event timing of the fused conv kernel on one UNet layer shape (roofline)."""
    dev = "cuda"
    x = torch.rand(N, hh, ww, cin, device=dev)
    sc = torch.rand(cin, device=dev) + 0.5
    sh = torch.rand(cin, device=dev)
    w = torch.randn(cout, cin, 3, 3, device=dev) * 0.05
    b = torch.randn(cout, device=dev) * 0.1
    out = torch.empty(N, hh, ww, cout, device=dev)
    st = torch.zeros(2 * cout, device=dev, dtype=torch.float64)
    d = ops.conv_desc([Source(x, sc, sh)], N, hh, ww, cout, (3, 3), 1, 0.01, ops.MATH_TF32)
    wp = ops.prep_weights(w, ops.WMODE_FWD, ops.MATH_TF32)
    for _ in range(3):
        ops.conv_fwd(d, wp, b, out, st)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        ops.conv_fwd(d, wp, b, out, st)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    flops = 2.0 * N * hh * ww * cin * cout * 9
    byts = 4.0 * N * hh * ww * (cin + cout)
    return ms, flops / ms / 1e9, byts / ms / 1e6


def run_ours(args):
    import atomai_b200 as ab
    from atomai_b200 import _C, ops
    from atomai_b200.models import Segmentor
    from atomai_b200.ops import Source
    from atomai_b200.parallel import init_distributed
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus or world == 1, "launch N > 1 with torchrun --nproc-per-node N"
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU path); use --impl reference for the host arm")
    torch.cuda.set_device(local_rank)
    comm = init_distributed()
    ab.set_math(args.math)
    pk = peaks()

    gb = BATCH * world                       # weak scaling: 32 images per GPU
    n_train_batches = 2
    X, y = synth(gb * n_train_batches, 1)
    Xt, yt = synth(gb, 2)
    m = Segmentor("Unet", nb_classes=NB_CLASSES, seed=1)
    cycles = args.warmup + args.steps
    m.compile_trainer((X, y, Xt, yt), loss="ce", training_cycles=2 * cycles + 8, batch_size=gb,
                      full_epoch=False, memory_alloc=64, plot_training_history=False,
                      sync_host=False, sync_bn=True, filename="/tmp/bench_model")
    dev = torch.device("cuda", local_rank)

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident loop (value)
    e = 0
    for _ in range(args.warmup):
        m.step(e)
        e += 1
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    calls0 = _C.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        m.step(e)
        e += 1
    ev1.record()
    barrier()
    launches = _C.launch_count() - calls0
    t = torch.tensor([ev0.elapsed_time(ev1)], device=dev, dtype=torch.float64)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    ms_total = float(t.item())
    final_losses = [float(v) for v in m.loss_acc["train_loss"][-2:]]

    # train-step-only timing (secondary figure)
    barrier()
    ev0.record()
    for i in range(args.steps):
        f, tg = m.dataloader(m.batch_idx_train[i], mode="train")
        m.train_step(f, tg)
    ev1.record()
    barrier()
    t2 = torch.tensor([ev0.elapsed_time(ev1)], device=dev, dtype=torch.float64)
    if world > 1:
        torch.distributed.all_reduce(t2, op=torch.distributed.ReduceOp.MAX)
    ms_train_only = float(t2.item()) / args.steps

    # ---------------- end-to-end loop: the same public call (Segmentor.fit's cycle, `step`) on a
    # HOST-resident data set (memory_alloc = 0 GB keeps the batches in pinned host memory, as the
    # reference does for data sets that do not fit): every cycle copies its images + labels
    # host->device (on the trainer's copy stream, overlapping the previous cycle's kernels) and
    # reads the losses back (loss.item()), all inside the timed region.
    del m
    torch.cuda.empty_cache()
    m2 = Segmentor("Unet", nb_classes=NB_CLASSES, seed=1)
    n_e2e = max(1, args.warmup // 2) + args.steps
    m2.compile_trainer((X, y, Xt, yt), loss="ce", training_cycles=n_e2e + 2, batch_size=gb,
                       full_epoch=False, memory_alloc=0, plot_training_history=False,
                       sync_host=True, sync_bn=True, filename="/tmp/bench_model_e2e")
    assert not m2.X_train[0].is_cuda and m2.X_train[0].is_pinned(), "e2e data must be pinned host memory"
    lb = BATCH
    h2d = sum(t_[0][:lb].numel() * t_[0].element_size()
              for t_ in (m2.X_train, m2.y_train, m2.X_test, m2.y_test))
    e = 0
    for _ in range(max(1, args.warmup // 2)):
        m2.step(e)
        e += 1
    barrier()
    ev0.record()
    for _ in range(args.steps):
        m2.step(e)
        e += 1
    ev1.record()
    barrier()
    t3 = torch.tensor([ev0.elapsed_time(ev1)], device=dev, dtype=torch.float64)
    if world > 1:
        torch.distributed.all_reduce(t3, op=torch.distributed.ReduceOp.MAX)
    ms_e2e = float(t3.item())
    sampler.stop_flag = True          # clocks are sampled over all three timed regions

    if rank != 0:
        return
    ips = gb * args.steps / (ms_total / 1e3)
    ips_e2e = gb * args.steps / (ms_e2e / 1e3)
    step_tflops = FLOP_PER_PIXEL_TRAIN * H * W * BATCH * world / (ms_train_only / 1e3) / 1e12

    # ---------------- roofline of the dominant kernel (conv_tc_kernel), timed live on this GPU
    tf32_peak = pk["bf16"] / 2.0             # TF32 dense = half the measured bf16 rate
    layers = {}
    for tag, (hh, cin, cout) in {"c5.block.0": (256, 64, 32), "c4.block.0": (128, 128, 64),
                                 "bn.block.3": (64, 128, 128), "c6.block.0": (512, 32, 16)}.items():
        ms, tfl, gbs = time_layer(ops, Source, BATCH, hh, hh, cin, cout)
        layers[tag] = {"ms": round(ms, 4), "tflops": round(tfl, 1), "gbs": round(gbs, 1),
                       "frac_tensor": round(tfl / tf32_peak, 3), "frac_hbm": round(gbs / pk["hbm"], 3)}
    dom = "bn.block.3"
    # DRAM traffic of that kernel (dram__bytes_read.sum + dram__bytes_write.sum of one launch) from
    # the committed `ncu --set full` capture of the same layer shape
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "r01_roofline.json")))["traffic_bytes"]
    except Exception:  # noqa
        pass
    roof = {"bound": "tensor", "kernel": "conv_tc_kernel (fused conv3x3+bias+LeakyReLU+BN stats), "
            f"layer {dom} 128->128 @64x64, batch {BATCH}", "achieved": layers[dom]["tflops"],
            "peak": round(tf32_peak, 1), "unit": "TFLOP/s", "frac": layers[dom]["frac_tensor"],
            "peak_note": f"0.5 x {pk['src']} dense bf16 ({pk['bf16']} TF/s): TF32 operands",
            "traffic": traffic, "traffic_note": "bytes per launch, profiles/r01_roofline.json "
            "(ncu --set full; algorithmic bytes = 4*N*H*W*(Cin+Cout) = 134.2 MB, the 126 MB L2 "
            "holds back part of the output writes)", "layers": layers,
            "step_tflops_algorithmic": round(step_tflops, 1),
            "step_frac": round(step_tflops / tf32_peak, 3)}

    # ---------------- CPU baseline (oracle port) on a bounded sample, rank 0, N = 1 only
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            out = subprocess.check_output(
                [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "4",
                 "--warmup", "1", "--ref-batch", str(args.ref_batch)], encoding="utf-8",
                timeout=600, env={**os.environ, "CUDA_VISIBLE_DEVICES": ""})
            ref = json.loads([l for l in out.splitlines() if l.startswith("{")][-1])
            cpu = ref["cpu_baseline"]
        except Exception as ex:  # noqa
            cpu = {"value": None, "unit": "images/s", "cores": host_cores(), "kind": "port",
                   "sample": f"failed: {type(ex).__name__}: {ex}"[:200]}

    line = {
        "metric": "Segmentor.fit images/sec (512x512, 3-class)",
        "value": ips, "unit": "images/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "tf32" if args.math == "tf32" else "f32",
        "data": "synthetic",
        "config": {"workload": f"Segmentor('Unet', nb_classes=3).fit cycle = 1 train step "
                               f"(fwd+bwd+fused Adam) + 1 test forward; batch {BATCH} x {H}x{W}x1 "
                               f"fp32 per GPU, global batch {gb}",
                   "parallelism": f"dp{world}", "sync_bn": True, "math": args.math,
                   "l2": "inputs larger than L2: ~19 GB of activations per step, no flush needed",
                   "train_step_only_ms": ms_train_only,
                   "train_step_only_images_per_s": gb / (ms_train_only / 1e3),
                   "final_train_losses": final_losses},
        "e2e": {"value": ips_e2e, "unit": "images/s", "h2d_bytes_per_step": h2d * world,
                "d2h_bytes_per_step": 8 * world, "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": launches,
        "clocks": sampler.summary(),
        "roofline": roof,
        "cpu_baseline": cpu,
    }
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--math", default="tf32", choices=["tf32", "fp32"])
    ap.add_argument("--ref-batch", type=int, default=8,
                    help="images per step of the bounded CPU sample (reference arm / cpu_baseline)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    a = ap.parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            torch.distributed.destroy_process_group()
