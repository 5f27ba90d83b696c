import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from __graft_entry__ import load_package, load_oracle
import test_modules as tm
pkg = load_package(); o = load_oracle()
def cpu_core(value, shapes, lsi, loc, w, step):
    return o.core_grid_sample(value.cpu(), shapes.cpu(), loc.cpu(), w.cpu()).to(value.device)
for name, core in (("cuda-core", None), ("oracle-core-on-gpu-tensors", cpu_core)):
    m, g = tm.build_small(pkg, core=core, dtype=torch.float64, device="cuda:0")
    srcs = [torch.from_numpy(g[f"src{i}"]).cuda() for i in range(3)]
    pos = [torch.from_numpy(g[f"pos{i}"]).cuda() for i in range(3)]
    with torch.no_grad():
        mem = m(srcs, pos)[0]
    print(name, np.abs(mem.cpu().numpy() - g["memory"]).max())
m, g = tm.build_small(pkg, core=tm.oracle_core(o), dtype=torch.float64, device="cpu")
srcs = [torch.from_numpy(g[f"src{i}"]) for i in range(3)]
pos = [torch.from_numpy(g[f"pos{i}"]) for i in range(3)]
with torch.no_grad():
    print("cpu", np.abs(m(srcs, pos)[0].numpy() - g["memory"]).max())
    src, p, shapes, lsi, levels = m.flatten_inputs(srcs, pos)
    a = m.encoder.layers[0].self_attn
    v_cpu = a.project_value(src); loc_cpu, w_cpu = a.sampling_inputs(src + p, pkg.modules.reference_points_for(levels, "cpu").expand(2, -1, -1, -1), shapes)
    mg = tm.build_small(pkg, dtype=torch.float64, device="cuda:0")[0]
    ag = mg.encoder.layers[0].self_attn
    v_g = ag.project_value(src.cuda()); loc_g, w_g = ag.sampling_inputs((src + p).cuda(), pkg.modules.reference_points_for(levels, "cuda").expand(2, -1, -1, -1), shapes.cuda())
    print("value_proj diff", (v_g.cpu() - v_cpu).abs().max().item(), "loc diff", (loc_g.cpu() - loc_cpu).abs().max().item(), "w diff", (w_g.cpu() - w_cpu).abs().max().item())
    print("refpoints diff", (pkg.modules.reference_points_for(levels, "cuda").cpu() - pkg.modules.reference_points_for(levels, "cpu")).abs().max().item())
