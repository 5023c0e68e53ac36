"""A/B of the conv_tc tile-plan knobs (ATOMAI_B200_NA4 / _SMEM_KB / _MIN_NR / _TMA_KC8) in ONE
process: the plan is recomputed (and the environment re-read) at every launch, so a variant is
just an os.environ update.  Part 1: forward + dgrad of every tensor-core layer shape of the
default Unet at the bench workload (batch 32 x 512^2, tf32x3), outputs compared against variant
A (max abs difference: a different k-chunk reorders the fp32 accumulation, ~1e-7 relative).
Part 2: the Segmentor.fit cycle under each variant.  usage: plan_ab.py [layers] [step]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from atomai_b200 import ops
from atomai_b200.ops import Source

VARIANTS = [  # name, env
    ("A base", {}),
    ("B na4=1 224", {"ATOMAI_B200_NA4": "1", "ATOMAI_B200_SMEM_KB": "224", "ATOMAI_B200_MIN_NR": "4"}),
    ("C na4=3 224", {"ATOMAI_B200_NA4": "3", "ATOMAI_B200_SMEM_KB": "224", "ATOMAI_B200_MIN_NR": "4"}),
    ("D na4=3 212", {"ATOMAI_B200_NA4": "3", "ATOMAI_B200_SMEM_KB": "212"}),
    ("E C+kc8tma", {"ATOMAI_B200_NA4": "3", "ATOMAI_B200_SMEM_KB": "224", "ATOMAI_B200_MIN_NR": "4",
                    "ATOMAI_B200_TMA_KC8": "1"}),
    ("F na4=2 224", {"ATOMAI_B200_NA4": "2", "ATOMAI_B200_SMEM_KB": "224"}),
]
KEYS = ["ATOMAI_B200_NA4", "ATOMAI_B200_SMEM_KB", "ATOMAI_B200_MIN_NR", "ATOMAI_B200_TMA_KC8",
        "ATOMAI_B200_RES_NA4", "ATOMAI_B200_PLAN"]
if os.environ.get("PLAN_AB_SET") == "res":     # resident n_a = 2 vs streamed n_a = 4 (c3.0, c5.0)
    VARIANTS = [("A default", {}), ("G res_na4", {"ATOMAI_B200_RES_NA4": "1"}),
                ("H streamed s1", {"ATOMAI_B200_PLAN": "2,16"}),
                ("I res kc8 tma", {"ATOMAI_B200_NA4": "3", "ATOMAI_B200_TMA_KC8": "1", "ATOMAI_B200_PLAN": "0,8"})]


def set_variant(env):
    for k in KEYS:
        os.environ.pop(k, None)
    os.environ.update(env)


N = 32
LAYERS = [  # name, H, [cins], cout, ks, pool
    ("c1.1", 512, [16], 16, 3, False),
    ("c2.0", 256, [16], 32, 3, True), ("c2.1", 256, [32], 32, 3, False),
    ("c3.0", 128, [32], 64, 3, True), ("c3.1", 128, [64], 64, 3, False),
    ("bn.0", 64, [64], 128, 3, True), ("bn.1", 64, [128], 128, 3, False),
    ("c4.0", 128, [64, 64], 64, 3, False),
    ("u2", 128, [64], 32, 1, False), ("c5.0", 256, [32, 32], 32, 3, False),
    ("u3", 256, [32], 16, 1, False), ("c6.0", 512, [16, 16], 16, 3, False),
]
if os.environ.get("PLAN_AB_SET") == "res":
    LAYERS = [l for l in LAYERS if l[0] in ("c3.0", "c5.0", "c2.0")]
MATH = ops.MATH_TF32X3
dev = "cuda"


def timeit(fn, reps=4):
    for _ in range(2):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


def part_layers():
    print("layer        " + " | ".join(f"{v[0]:>13s}" for v in VARIANTS) + "   (us; fwd then dgrad)", flush=True)
    tot = [0.0] * len(VARIANTS)
    for name, hh, cins, cout, ks, pool in LAYERS:
        torch.manual_seed(0)
        srcs = []
        for ci in cins:
            s = 2 * hh if pool else hh
            srcs.append(Source(torch.rand(N, s, s, ci, device=dev), torch.rand(ci, device=dev) + 0.5,
                               torch.rand(ci, device=dev) - 0.5, pool))
        cin = sum(cins)
        w = torch.randn(cout, cin, ks, ks, device=dev) * 0.05
        b = torch.randn(cout, device=dev) * 0.1
        out = torch.empty(N, hh, hh, cout, device=dev)
        st = torch.zeros(2 * cout, device=dev, dtype=torch.float64)
        d = ops.conv_desc(srcs, N, hh, hh, cout, (ks, ks), 1, 0.01, MATH)
        wp = ops.prep_weights(w, ops.WMODE_FWD, MATH)
        row, ref = [], None
        for vi, (_, env) in enumerate(VARIANTS):
            set_variant(env)
            try:
                out.zero_()
                us = timeit(lambda: ops.conv_fwd(d, wp, b, out, st))
                if ref is None:
                    ref = out.clone()
                    row.append(f"{us:13.1f}")
                else:
                    dmax = (out - ref).abs().max().item()
                    row.append(f"{us:8.1f}{' ok  ' if dmax == 0 else f'!{dmax:.0e}'}")
                tot[vi] += us
            except Exception as ex:   # noqa: BLE001
                row.append(f"{'ERR':>13s}")
                print("  ", type(ex).__name__, str(ex)[:150])
        print(f"{name:5s} fwd    " + " | ".join(row), flush=True)
        del ref
        if not pool:
            dy = torch.randn(N, hh, hh, cout, device=dev)
            dd = ops.conv_desc([Source(dy)], N, hh, hh, cin, (ks, ks), 1, 1.0, MATH, act=ops.ACT_LRELU)
            wd = ops.prep_weights(w, ops.WMODE_DGRAD, MATH)
            dx = torch.empty(N, hh, hh, cin, device=dev)
            row, ref = [], None
            for vi, (_, env) in enumerate(VARIANTS):
                set_variant(env)
                try:
                    dx.zero_()
                    us = timeit(lambda: ops.conv_fwd(dd, wd, None, dx, None))
                    if ref is None:
                        ref = dx.clone()
                        row.append(f"{us:13.1f}")
                    else:
                        dmax = (dx - ref).abs().max().item()
                        row.append(f"{us:8.1f}{' ok  ' if dmax == 0 else f'!{dmax:.0e}'}")
                    tot[vi] += us
                except Exception as ex:   # noqa: BLE001
                    row.append(f"{'ERR':>13s}")
                    print("  ", type(ex).__name__, str(ex)[:150])
            print(f"{name:5s} dgrad  " + " | ".join(row), flush=True)
            del dy, dx, ref
        del srcs, out
        torch.cuda.empty_cache()
    print("sum          " + " | ".join(f"{t:13.1f}" for t in tot), flush=True)
    set_variant({})


def part_step():
    import atomai_b200 as ab
    from atomai_b200.models import Segmentor
    from bench import synth, BATCH, NB_CLASSES
    ab.set_math("tf32x3")
    X, y = synth(2 * BATCH, 1, 512)
    Xt, yt = synth(BATCH, 2, 512)
    m = Segmentor("Unet", nb_classes=NB_CLASSES, seed=1)
    m.compile_trainer((X, y, Xt, yt), loss="ce", training_cycles=20 * len(VARIANTS) + 8, batch_size=BATCH,
                      full_epoch=False, memory_alloc=64, plot_training_history=False, sync_host=False,
                      filename="/tmp/plan_ab_model")
    e = 0
    for rnd in range(2):          # two rounds: the second shows run-to-run noise
        for name, env in VARIANTS:
            set_variant(env)
            for _ in range(3):
                m.step(e); e += 1
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                m.step(e); e += 1
            e1.record()
            torch.cuda.synchronize()
            print(f"step  {name:14s} {e0.elapsed_time(e1) / 5:7.3f} ms / fit cycle  "
                  f"(loss {float(m.loss_acc['train_loss'][-1]):.5f})", flush=True)
    set_variant({})


if __name__ == "__main__":
    which = sys.argv[1:] or ["layers", "step"]
    if "layers" in which:
        part_layers()
    if "step" in which:
        part_step()
