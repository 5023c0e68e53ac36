"""Bring-up tool: (1) stress the tcgen05 conv for intermittent errors, (2) wgrad tcgen05 kernel,
(3) per-case parity numbers of the native Unet vs goldens (verbose, no asserts).
Writes gpurun_out/check2.json."""
import json, os, sys, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import numpy as np, torch
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
from atomai_b200 import ops
from atomai_b200.ops import Source
import atomai_b200 as ab
res = {}
dev = "cuda"
g = torch.Generator(device="cpu").manual_seed(0)
def rnd(*s, scale=1.0): return (torch.randn(*s, generator=g) * scale).to(dev)
def rel(a, b): return float((a - b).abs().max() / (b.abs().max() + 1e-20))

def stress():
    out = {}
    shapes = [(2, 32, 32, 64, 32), (2, 32, 32, 32, 64), (4, 64, 64, 32, 16), (2, 32, 32, 128, 128), (8, 64, 64, 16, 32), (1, 48, 40, 64, 64)]
    for (N, H, W, ci, co) in shapes:
        x = rnd(N, H, W, ci); w = rnd(co, ci, 3, 3, scale=0.1); b = rnd(co)
        ds = ops.conv_desc([Source(x)], N, H, W, co, (3, 3), 1, 0.01, ops.MATH_FP32)
        wps = ops.prep_weights(w, ops.WMODE_FWD, ops.MATH_FP32)
        ref = torch.empty(N, H, W, co, device=dev); ops.conv_fwd(ds, wps, b, ref, None)
        dt = ops.conv_desc([Source(x)], N, H, W, co, (3, 3), 1, 0.01, ops.MATH_TF32)
        bad = 0; worst = 0.0
        for it in range(60):
            wpt = ops.prep_weights(w, ops.WMODE_FWD, ops.MATH_TF32)
            o = torch.full((N, H, W, co), float("nan"), device=dev)
            st = torch.zeros(2 * co, device=dev, dtype=torch.float64)
            ops.conv_fwd(dt, wpt, b, o, st)
            r = rel(o, ref)
            if not (r < 2e-3): bad += 1
            worst = max(worst, r if r == r else 9e9)
        out[f"{N}x{H}x{W}_{ci}to{co}"] = {"bad": bad, "worst": worst}
    return out

def wgrad():
    import torch.nn.functional as F
    out = {}
    def case(tag, N, H, W, cins, Cout, ks=3, dil=1, affine=False, pool=False):
        srcs, refs = [], []
        for ci in cins:
            hh, ww = (2 * H, 2 * W) if pool else (H, W)
            x = rnd(N, ci, hh, ww); sc = sh = None; xr = x
            if affine:
                sc = (torch.rand(ci, generator=g) + 0.5).to(dev); sh = rnd(ci, scale=0.3)
                xr = x * sc.view(1, -1, 1, 1) + sh.view(1, -1, 1, 1)
            if pool: xr = F.max_pool2d(xr, 2, 2)
            srcs.append(Source(x.permute(0, 2, 3, 1).contiguous(), sc, sh, pool)); refs.append(xr)
        xin = torch.cat(refs, 1); cin = xin.shape[1]
        dy = rnd(N, Cout, H, W)
        w = torch.zeros(Cout, cin, ks, ks, device=dev, requires_grad=True)
        F.conv2d(xin, w, None, padding=dil * (ks // 2), dilation=dil).backward(dy)
        d = ops.conv_desc(srcs, N, H, W, Cout, (ks, ks), dil, 1.0, ops.MATH_TF32)
        dw = torch.zeros(Cout, cin, ks, ks, device=dev)
        ops.conv_wgrad(d, dy.permute(0, 2, 3, 1).contiguous(), dw); torch.cuda.synchronize()
        out[tag] = rel(dw, w.grad)
    case("32to32_16x8", 1, 16, 8, [32], 32)
    case("32to16", 2, 64, 64, [32], 16)
    case("16to32", 2, 32, 32, [16], 32)
    case("64to64", 2, 32, 32, [64], 64)
    case("128to128", 2, 32, 24, [128], 128)
    case("128to256", 1, 16, 16, [128], 256)
    case("1x1_128to64", 2, 32, 32, [128], 64, ks=1)
    case("cat_affine_pool", 2, 32, 32, [16, 16], 32, affine=True, pool=True)
    case("ragged_48", 3, 37, 29, [32], 48)
    case("dil2", 2, 32, 32, [64], 128, dil=2)
    case("cin20_cout12", 2, 24, 24, [20], 12)
    case("big_c5", 2, 256, 256, [32, 32], 32)
    return out

def unet():
    from test_oracle import CASES, build_case
    from atomai_b200.losses_metrics import select_loss
    import golden_utils as gu
    out = {}
    for math in ("fp32", "tf32", "tf32_simtwgrad"):
        ab.set_math(math.split("_")[0], wgrad_tc=(math != "tf32_simtwgrad"))
        for name in CASES:
            try:
                net, sd, cfg, x, y, gold = build_case(name)
                net = net.to(dev); x = x.to(dev); y = y.to(dev)
                net.eval()
                with torch.no_grad(): le = net(x)
                from test_oracle import logits_view
                r = {"eval": float(np.abs(logits_view(le.cpu().numpy(), gold) - gold["logits_eval"]).max() / np.abs(gold["logits_eval"]).max())}
                net.train(); net.zero_grad()
                lt = net(x)
                r["train"] = float(np.abs(logits_view(lt.detach().cpu().numpy(), gold) - gold["logits_train"]).max() / np.abs(gold["logits_train"]).max())
                loss = select_loss("ce", cfg["nb_classes"])(lt, y)
                r["loss"] = [loss.item(), float(gold["loss_train"])]
                loss.backward()
                worst = ("", 0.0); worstn = ("", 0.0); te = tr = 0.0
                for k, p in net.named_parameters():
                    gg = p.grad.detach().cpu().numpy(); ref = gold["grad/" + k]
                    got = gu.sample_flat(gg, 97) if gg.size > 4096 else gg
                    te += float(((got.reshape(-1) - ref.reshape(-1)).astype(np.float64) ** 2).sum()); tr += float((ref.reshape(-1).astype(np.float64) ** 2).sum())
                    scale = max(float(gold["gradnorm/" + k]) / np.sqrt(gg.size), 1e-8)
                    e = float(np.abs(got.reshape(-1) - ref.reshape(-1)).max() / scale)
                    if e > worst[1]: worst = (k, e)
                    en = abs(float(np.linalg.norm(gg.astype(np.float64))) - float(gold["gradnorm/" + k])) / (float(gold["gradnorm/" + k]) + 1e-12)
                    if en > worstn[1]: worstn = (k, en)
                r["grad_worst_elem_over_rms"] = worst; r["gradnorm_worst_rel"] = worstn; r["grel"] = (te / tr) ** 0.5
                rb = 0.0
                for k, b in net.named_buffers():
                    if "running" in k:
                        rb = max(rb, float(np.abs(b.cpu().numpy() - gold["buf/" + k]).max()))
                r["running_abs"] = rb
                out[f"{math}/{name}"] = r
            except Exception as e:
                out[f"{math}/{name}"] = {"error": f"{type(e).__name__}: {e}", "tb": traceback.format_exc()[-800:]}
    return out

for nm, fn in [("stress", stress), ("wgrad_tc", wgrad), ("unet", unet)]:
    if len(sys.argv) > 1 and nm not in sys.argv[1:]: continue
    try:
        res[nm] = fn()
    except Exception as e:
        res[nm] = {"error": f"{type(e).__name__}: {e}", "tb": traceback.format_exc()[-1500:]}
    print(nm, json.dumps(res[nm])[:6000], flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "check2.json"), "w"), indent=1)
