"""
The BASELINE.json configurations besides the headline one, behind `bench.py --workload ...`:

  rvae    cfg3   rVAE((64,64), latent_dim=2, conv_encoder=True) on 10k synthetic 64x64 patches,
                 batch 100 (a "step" = one mini-batch of viBaseTrainer.train_epoch)
  imspec  cfg4   ImSpec((64,64), (128,), latent_dim=10), global batch 256 (data parallel: 256/N per GPU),
                 loss 'mse' (a "step" = one fit cycle: train mini-batch + test mini-batch)
  gram    cfg5i  dense RBF Gram on 50k x 128 embedded features (row blocks over the ranks),
                 metric = GB/s of K written against the measured HBM copy rate

Each returns the one-line dict bench.py prints; the reference arm (`--impl reference`) for these
workloads runs the unmodified reference's own model classes on the host cores.
"""
import json
import os
import time

import numpy as np
import torch


def _max_ms(ms, dev, world):
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    return float(t.item())


def _barrier(world):
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()


def _math(args):
    return "tf32x3" if args.math == "auto" else args.math


# ----------------------------------------------------------------------------- rVAE (cfg3)
def run_rvae(args, world, rank, dev, pk, metric):
    import atomai_b200 as ab
    from atomai_b200 import _C
    from atomai_b200.models import rVAE
    ab.set_math(_math(args))
    B = 100 * world if args.scaling == "weak" else 100
    K, W = args.steps, max(args.warmup, 1)
    rs = np.random.RandomState(1)
    out = {}
    for tag, alloc in (("device", 64), ("e2e", 0)):
        m = rVAE((64, 64), latent_dim=2, conv_encoder=True, seed=1)
        Xw = rs.rand(W * B, 64, 64).astype(np.float32)
        Xk = rs.rand(K * B, 64, 64).astype(np.float32)
        m.compile_trainer((Xw, None), training_cycles=1, batch_size=B, memory_alloc=alloc,
                          filename="/tmp/bench_rvae")
        m.kdict_["phi_prior"] = 0.1
        m.dx_prior = 0.1
        m.train_epoch()                                    # warm-up: W mini-batches
        m.train_iterator = m._set_data(Xk, None, store_on_cpu=(alloc == 0))
        _barrier(world)
        c0 = _C.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        elbo = m.train_epoch()                             # exactly K mini-batches
        e1.record()
        _barrier(world)
        ms = _max_ms(e0.elapsed_time(e1), dev, world)
        out[tag] = {"ms_per_step": ms / K, "images_per_s": B * K / (ms / 1e3), "elbo": float(elbo),
                    "launches": _C.launch_count() - c0}
        del m
        torch.cuda.empty_cache()
    if rank != 0:
        return None
    # algorithmic work per image: encoder conv 1.217 GFLOP fwd, rDecoder 0.271 GFLOP
    # fwd, x3 for training; the two 524288x5 heads stream 21 MB of weights forward and backward
    flop_img = 3.0 * (1.217e9 + 0.271e9)
    tfl = flop_img * out["device"]["images_per_s"] / 1e12
    tf32_peak = pk["bf16_sustained"] / 2
    return {
        "metric": metric, "value": out["device"]["images_per_s"], "unit": "images/s", "n_gpus": world,
        "steps": K, "warmup": W, "ms_per_step": out["device"]["ms_per_step"], "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": _math(args), "data": "synthetic",
        "config": {"workload": f"rVAE((64,64), latent_dim=2, conv_encoder=True) train mini-batches, "
                               f"batch {B // world} x 64x64 fp32 per GPU (global {B})",
                   "parallelism": f"dp{world}", "l2": "activations of one step (>3 GB) exceed L2",
                   "final_elbo": out["device"]["elbo"]},
        "e2e": {"value": out["e2e"]["images_per_s"], "unit": "images/s",
                "h2d_bytes_per_step": B * 64 * 64 * 4, "d2h_bytes_per_step": 4 * world,
                "ms_per_step": out["e2e"]["ms_per_step"]},
        "gpu_launches": out["device"]["launches"],
        "roofline": {"bound": "tensor", "achieved": round(tfl, 1), "peak": round(tf32_peak, 1),
                     "unit": "TFLOP/s", "frac": round(tfl / tf32_peak, 3), "traffic": None,
                     "kernel": "whole step, algorithmic FLOPs 3 x (1.217 + 0.271) GFLOP per image"},
        "cpu_baseline": None,
    }


# ----------------------------------------------------------------------------- ImSpec (cfg4)
def run_imspec(args, world, rank, dev, pk, metric):
    import atomai_b200 as ab
    from atomai_b200 import _C
    from atomai_b200.models import ImSpec
    ab.set_math(_math(args))
    gb = 256 if args.scaling == "strong" or world > 1 else 256   # BASELINE: global batch 256
    K, W = args.steps, args.warmup
    rs = np.random.RandomState(1)
    X = rs.rand(2 * gb, 1, 64, 64).astype(np.float32)
    y = rs.rand(2 * gb, 1, 128).astype(np.float32)
    Xt = rs.rand(gb, 1, 64, 64).astype(np.float32)
    yt = rs.rand(gb, 1, 128).astype(np.float32)
    out = {}
    for tag, alloc, sync in (("device", 64, False), ("e2e", 0, True)):
        m = ImSpec((64, 64), (128,), latent_dim=10, seed=1)
        m.compile_trainer((X, y, Xt, yt), loss="mse", training_cycles=W + K + 2, batch_size=gb,
                          full_epoch=False, memory_alloc=alloc, plot_training_history=False,
                          sync_host=sync, filename="/tmp/bench_imspec")
        e = 0
        for _ in range(W):
            m.step(e)
            e += 1
        _barrier(world)
        c0 = _C.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(K):
            m.step(e)
            e += 1
        e1.record()
        _barrier(world)
        ms = _max_ms(e0.elapsed_time(e1), dev, world)
        out[tag] = {"ms_per_step": ms / K, "images_per_s": gb * K / (ms / 1e3),
                    "launches": _C.launch_count() - c0,
                    "loss": float(m.loss_acc["train_loss"][-1])}
        del m
        torch.cuda.empty_cache()
    if rank != 0:
        return None
    flop_img = 3.0 * 0.609e9 + 0.609e9       # encoder conv: train step (3x fwd) + test forward
    tfl = flop_img * out["device"]["images_per_s"] / 1e12
    tf32_peak = pk["bf16_sustained"] / 2
    return {
        "metric": metric, "value": out["device"]["images_per_s"], "unit": "images/s", "n_gpus": world,
        "steps": K, "warmup": W, "ms_per_step": out["device"]["ms_per_step"], "higher_is_better": True,
        "scaling": "strong" if world > 1 else args.scaling, "vs_baseline": None, "dtype": _math(args),
        "data": "synthetic",
        "config": {"workload": f"ImSpec((64,64),(128,), latent_dim=10).fit cycle (train mini-batch + "
                               f"test mini-batch), global batch {gb} ({gb // world} per GPU), loss mse",
                   "parallelism": f"dp{world}", "final_train_loss": out["device"]["loss"]},
        "e2e": {"value": out["e2e"]["images_per_s"], "unit": "images/s",
                "h2d_bytes_per_step": 2 * gb * (64 * 64 + 128) * 4, "d2h_bytes_per_step": 8 * world,
                "ms_per_step": out["e2e"]["ms_per_step"]},
        "gpu_launches": out["device"]["launches"],
        "roofline": {"bound": "tensor", "achieved": round(tfl, 1), "peak": round(tf32_peak, 1),
                     "unit": "TFLOP/s", "frac": round(tfl / tf32_peak, 3), "traffic": None,
                     "kernel": "whole cycle, algorithmic FLOPs of the encoder ConvBlock "
                               "(0.609 GFLOP/img fwd)"},
        "cpu_baseline": None,
    }


# ----------------------------------------------------------------------------- Gram (cfg5 i)
def run_gram(args, world, rank, dev, pk, metric):
    import atomai_b200 as ab
    from atomai_b200 import ops
    mode = _math(args)
    ab.set_math(mode)
    n, d = 50_000, 128
    rows = n // world                      # row-block partition: rank r -> K[r*n/G:(r+1)*n/G, :]
    g = torch.Generator(device="cpu").manual_seed(1)
    Z = torch.randn(n, d, generator=g).to(dev)
    inv_ls = torch.full((d,), 1.0 / d ** 0.5, device=dev)
    Zr = Z[rank * rows:(rank + 1) * rows].contiguous()
    K_out = torch.empty((rows, n), device=dev, dtype=torch.float32)
    mm = {"fp32": ops.MATH_FP32, "tf32": ops.MATH_TF32, "tf32x3": ops.MATH_TF32X3}[mode]
    for _ in range(max(args.warmup, 1)):
        ops.gram(Zr, Z, inv_ls, 1.0, 0, K_out, mm)
    _barrier(world)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        ops.gram(Zr, Z, inv_ls, 1.0, 0, K_out, mm)
    e1.record()
    _barrier(world)
    ms = _max_ms(e0.elapsed_time(e1), dev, world) / args.steps
    # end to end: operands from pinned host memory, a (rows x 1024)-column slab of K read back
    Zh = Z.cpu().pin_memory()
    host_slab = torch.empty((rows, 1024), dtype=torch.float32).pin_memory()
    _barrier(world)
    e0.record()
    for _ in range(args.steps):
        Zd = Zh.to(dev, non_blocking=True)
        ops.gram(Zd[rank * rows:(rank + 1) * rows].contiguous(), Zd, inv_ls, 1.0, 0, K_out, mm)
        host_slab.copy_(K_out[:, :1024], non_blocking=True)
        torch.cuda.current_stream().synchronize()
    e1.record()
    _barrier(world)
    ms_e2e = _max_ms(e0.elapsed_time(e1), dev, world) / args.steps
    diag = float(torch.diagonal(K_out[:, rank * rows:(rank + 1) * rows]).mean())
    if rank != 0:
        return None
    byts = 4.0 * n * n                         # algorithmic: every element of K written once
    gbs = byts / (ms / 1e3) / 1e9
    flop = 2.0 * n * n * d
    return {
        "metric": metric, "value": gbs, "unit": "GB/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": mode, "data": "synthetic",
        "config": {"workload": f"dense RBF Gram K(Z, Z), Z = randn({n}, {d}) fp32, lengthscale sqrt(128), "
                               f"row blocks of {rows} per GPU, {n}x{n} fp32 = 10 GB written",
                   "parallelism": f"rows/{world}", "l2": "output (10 GB) exceeds L2",
                   "tflops_algorithmic": flop / (ms / 1e3) / 1e12, "mean_diagonal": diag},
        "e2e": {"value": byts / (ms_e2e / 1e3) / 1e9, "unit": "GB/s",
                "h2d_bytes_per_step": n * d * 4 * world, "d2h_bytes_per_step": rows * 1024 * 4 * world,
                "ms_per_step": ms_e2e},
        "gpu_launches": args.steps * (2 + 2 * ((n + 255) // 256)),
        "roofline": {"bound": "hbm", "achieved": round(gbs / world, 1), "peak": pk["hbm"], "unit": "GB/s",
                     "frac": round(gbs / world / pk["hbm"], 3), "traffic": None,
                     "kernel": "conv_tc_kernel with RBF epilogue (one launch per 256-column block); "
                               "algorithmic bytes = 4*n1*n2 per GPU"},
        "cpu_baseline": None,
    }


def run_other(args, world, rank, dev, pk):
    from bench import METRIC
    fn = {"rvae": run_rvae, "imspec": run_imspec, "gram": run_gram}[args.workload]
    line = fn(args, world, rank, dev, pk, METRIC[args.workload])
    return line


# ----------------------------------------------------------------------------- reference arms
def run_other_reference(args, cores):
    """CPU arm for the non-headline workloads: the unmodified reference's own classes (rVAE,
    ImSpec) or, for the Gram, a torch CPU restatement in row blocks (gpytorch is not installed and
    the reference never forms this matrix)."""
    from bench import METRIC
    torch.set_num_threads(cores)
    K, W = args.steps, args.warmup
    rs = np.random.RandomState(1)
    if args.workload == "gram":
        n, d, blk = 50_000, 128, 2000
        Z = torch.randn(n, d) / d ** 0.5
        nz = (Z * Z).sum(1)
        t0 = time.perf_counter()
        rows = 0
        for _ in range(max(K, 1)):
            a = Z[rows:rows + blk]
            Kb = torch.exp(-0.5 * torch.clamp(nz[rows:rows + blk, None] + nz[None] - 2 * a @ Z.T, min=0))
            rows += blk
        dt = time.perf_counter() - t0
        val = 4.0 * rows * n / dt / 1e9
        sample = f"{rows} of {n} rows (row blocks of {blk}), torch CPU fp32, {cores} threads"
        return {"impl": "reference", "metric": METRIC["gram"], "value": val, "unit": "GB/s",
                "n_gpus": args.gpus, "steps": K, "warmup": W, "ms_per_step": 1e3 * dt / max(K, 1),
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": {"workload": "dense RBF Gram 50k x 128 (bounded sample)"},
                "cpu_baseline": {"value": val, "unit": "GB/s", "cores": cores, "kind": "port",
                                 "sample": sample},
                "e2e": {"value": val, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    from baseline.ref_loader import import_reference
    aoi = import_reference()
    if args.workload == "rvae":
        B = 100
        m = aoi.models.rVAE((64, 64), latent_dim=2, conv_encoder=True, seed=1)
        X = rs.rand(K * B, 64, 64).astype(np.float32)
        m.compile_trainer((X, None), training_cycles=1, batch_size=B, filename="/tmp/bench_ref_rvae")
        m.kdict_["phi_prior"] = 0.1
        m.dx_prior = 0.1
        t0 = time.perf_counter()
        m.train_epoch()
        dt = time.perf_counter() - t0
        n_img, what = K * B, "rVAE((64,64), latent_dim=2, conv_encoder=True) mini-batches of 100"
    else:
        B = 256
        m = aoi.models.ImSpec((64, 64), (128,), latent_dim=10)
        X = rs.rand(2 * B, 1, 64, 64).astype(np.float32)
        y = rs.rand(2 * B, 1, 128).astype(np.float32)
        m.compile_trainer((X, y, X[:B], y[:B]), loss="mse", training_cycles=K + W + 1, batch_size=B,
                          full_epoch=False, plot_training_history=False, filename="/tmp/bench_ref_imspec")
        for e in range(W):
            m.step(e)
        t0 = time.perf_counter()
        for e in range(W, W + K):
            m.step(e)
        dt = time.perf_counter() - t0
        n_img, what = K * B, "ImSpec((64,64),(128,),10) fit cycles, batch 256"
    val = n_img / dt
    return {"impl": "reference", "metric": METRIC[args.workload], "value": val, "unit": "images/s",
            "n_gpus": args.gpus, "steps": K, "warmup": W, "ms_per_step": 1e3 * dt / K,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": {"workload": what, "parallelism": "cpu"},
            "cpu_baseline": {"value": val, "unit": "images/s", "cores": cores, "kind": "reference",
                             "sample": f"{K} steps, unmodified reference v{aoi.__version__}, {cores} threads"},
            "e2e": {"value": val, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
