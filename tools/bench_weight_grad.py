import os, sys, json, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package
pkg = load_package(); dev = "cuda:0"
def t(fn, n=10):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
M = 172032
for out_f, in_f in ((256, 256), (192, 256), (96, 256), (1024, 256), (256, 1024)):
    gy = torch.randn(M, out_f, device=dev); x = torch.randn(M, in_f, device=dev)
    gyt, xt = gy.t().contiguous(), x.t().contiguous()
    r = {"out": out_f, "in": in_f,
         "torch_dw_ms": t(lambda: gy.t() @ x),
         "transpose_gy_ms": t(lambda: gy.t().contiguous()), "transpose_x_ms": t(lambda: x.t().contiguous()),
         "my_transpose_gy_ms": t(lambda: pkg.ops.transpose2d(gy)), "my_transpose_x_ms": t(lambda: pkg.ops.transpose2d(x)),
         "transpose_ok": bool(torch.equal(pkg.ops.transpose2d(x), xt)),
         "kernel_dw_ms": t(lambda: pkg.linear_tf32x3(gyt, xt, None, split_weight_in_kernel=True)),
         "wgrad_kernel_ms": t(lambda: pkg.ops.linear_wgrad(gy, x)),
         "torch_bias_sum_ms": t(lambda: gy.sum(0))}
    ref = gy.double().t() @ x.double()
    for kbpc in (8, 16, 32, 64):
        pkg.set_option("wgrad_chunk", kbpc)
        gw_, _ = pkg.ops.linear_wgrad(gy, x)
        r[f"kb{kbpc}_ms"] = round(t(lambda: pkg.ops.linear_wgrad(gy, x)), 4)
        r[f"kb{kbpc}_err"] = (gw_.double() - ref).abs().max().item()
    pkg.set_option("linear_variant", 0)
    gw, gb = pkg.ops.linear_wgrad(gy, x)
    r["wgrad_err"] = (gw.double() - ref).abs().max().item()
    r["torch_err"] = ((gy.t() @ x).double() - ref).abs().max().item()
    r["bias_err"] = (gb.double() - gy.double().sum(0)).abs().max().item()
    r["torch_bias_err"] = (gy.sum(0).double() - gy.double().sum(0)).abs().max().item()
    print(json.dumps(r), flush=True)
