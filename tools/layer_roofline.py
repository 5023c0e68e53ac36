"""Per-launch roofline table of the tensor-core kernels of one Segmentor.fit cycle (default Unet,
batch 32 x 512^2, tf32x3) from an ncu launch list (`--metrics gpu__time_duration.sum`): layer,
time, algorithmic TFLOP/s and GB/s, the HBM floor (MEASURED_PEAKS.json) and, for conv_tc, the
shared-memory operand floor of DESIGN 3.3 (32 + N/4 clocks per M128 x N x K8 MMA, two MMAs per
k-step in tf32x3).  usage: layer_roofline.py launches.csv out.md"""
import csv, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
HBM, CLK, SMS, N = PK["hbm_gbs"], PK["sm_max_mhz"] * 1e6, 148, 32
FWD = [  # name, H (output), Cin, Cout, taps, pooled input
    ("c2.0", 256, 16, 32, 9, True), ("c2.1", 256, 32, 32, 9, False), ("c3.0", 128, 32, 64, 9, True),
    ("c3.1", 128, 64, 64, 9, False), ("bn.0", 64, 64, 128, 9, True), ("bn.1", 64, 128, 128, 9, False),
    ("bn.2", 64, 128, 128, 9, False), ("u1", 64, 128, 64, 1, False), ("c4.0", 128, 128, 64, 9, False),
    ("c4.1", 128, 64, 64, 9, False), ("u2", 128, 64, 32, 1, False), ("c5.0", 256, 64, 32, 9, False),
    ("c5.1", 256, 32, 32, 9, False), ("u3", 256, 32, 16, 1, False), ("c6.0", 512, 32, 16, 9, False)]


def main(src, dst):
    rows = list(csv.reader(open(src)))
    start = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[start]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    seq = [(r[ki], float(r[vi]) / 1e3) for r in rows[start + 2:] if len(r) > vi]
    conv = [t for n, t in seq if "conv_tc_kernel" in n]
    wg = [t for n, t in seq if "wgrad_tc_kernel" in n]
    assert len(conv) == 45 and len(wg) == 15, (len(conv), len(wg))
    out = ["# round 2 — per-layer roofline of the tensor-core kernels (one fit cycle, batch 32 x 512^2, tf32x3)", "",
           f"`python tools/layer_roofline.py {os.path.relpath(src, ROOT)}`: ncu `gpu__time_duration.sum` per launch "
           "(cold-cache, serialised: a few % above the in-step CUDA-event times).  `HBM floor` = algorithmic bytes / "
           f"{HBM:.0f} GB/s (MEASURED_PEAKS.json); `operand floor` = tiles per SM x k-steps x taps x 2 MMAs x "
           "(32 + N/4) clocks at 1.965 GHz — the shared-memory operand fetch of tcgen05.mma measured in DESIGN 3.3; "
           "`x floor` = time / max(floors).", "",
           "| pass | layer | shape | us | TFLOP/s alg. | GB/s alg. | HBM floor us | operand floor us | x floor |",
           "|---|---|---|---:|---:|---:|---:|---:|---:|"]
    tot = {}

    def row(kind, name, hh, k, n, taps, pooled, t, opfloor=True):
        px = N * hh * hh
        fl = 2.0 * px * taps * k * n
        by = 4.0 * px * (k * (4 if pooled else 1) + n)
        hbm = by / HBM / 1e3
        tiles = (px / 128) / SMS
        op = tiles * (k / 8) * taps * 2 * (32 + n / 4) / CLK * 1e6 if opfloor else 0.0
        fl_ = max(hbm, op)
        out.append(f"| {kind} | {name} | {k}->{n} @{hh}^2{' pooled' if pooled else ''} | {t:.0f} | "
                   f"{fl / t / 1e6:.0f} | {by / t / 1e3:.0f} | {hbm:.0f} | {op:.0f} | {t / fl_:.2f} |")
        a = tot.setdefault(kind, [0.0, 0.0, 0.0])
        a[0] += t; a[1] += fl; a[2] += fl_
    for i, (name, hh, ci, co, taps, pooled) in enumerate(FWD):
        row("fwd train", name, hh, ci, co, taps, pooled, conv[i])
    for i, (name, hh, ci, co, taps, pooled) in enumerate(reversed(FWD)):
        row("dgrad", name, hh, co, ci, taps, False, conv[15 + i])
    for i, (name, hh, ci, co, taps, pooled) in enumerate(FWD):
        row("fwd test", name, hh, ci, co, taps, pooled, conv[30 + i])
    for i, (name, hh, ci, co, taps, pooled) in enumerate(reversed(FWD)):
        row("wgrad", name, hh, ci, co, taps, pooled, wg[i], opfloor=False)
    out += ["", "| pass | total us | TFLOP/s alg. | sum of floors us | x floor |", "|---|---:|---:|---:|---:|"]
    for k, (t, fl, f) in tot.items():
        out.append(f"| {k} | {t:.0f} | {fl / t / 1e6:.0f} | {f:.0f} | {t / f:.2f} |")
    open(dst, "w").write("\n".join(out) + "\n")
    print("\n".join(out[-8:]))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
