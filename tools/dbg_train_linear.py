import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package
pkg = load_package(); DEV = "cuda:0"
torch.manual_seed(4)
levels = [(8, 16), (16, 32), (32, 64)]
kw = dict(d_model=256, nhead=8, num_encoder_layers=2, dim_feedforward=1024, dropout=0.0, num_feature_levels=3, enc_n_points=4)
a = pkg.modules.MSDeformAttnTransformerEncoderOnly(**kw).to(DEV).eval()
b = pkg.modules.MSDeformAttnTransformerEncoderOnly(linear="tf32x3", **kw).to(DEV).eval()
for m in a.modules():
    if isinstance(m, pkg.modules.MSDeformAttn):
        torch.nn.init.normal_(m.sampling_offsets.weight, std=0.02)
        torch.nn.init.normal_(m.attention_weights.weight, std=0.05)
b.load_state_dict(a.state_dict())
srcs = [torch.randn(2, 256, h, w, device=DEV) for h, w in levels]
pos = [torch.randn(2, 256, h, w, device=DEV) for h, w in levels]
def run(model):
    ss = [t.clone().requires_grad_(True) for t in srcs]
    model.zero_grad()
    y = model(ss, pos)[0]
    return y, ss
ya, sa = run(a); cot = torch.randn_like(ya); (ya * cot).sum().backward()
ga = [t.grad.clone() for t in sa]; pa = {n: p.grad.clone() for n, p in a.named_parameters() if p.grad is not None}
yb, sb = run(b); (yb * cot).sum().backward()
print("fwd max diff", (ya - yb).abs().max().item())
for i, (u, v) in enumerate(zip(ga, sb)):
    d = (u - v.grad)
    print("src", i, "relL2", (d.norm() / u.norm()).item(), "max", d.abs().max().item(), "frac>1e-3max", (d.abs() > 1e-3 * u.abs().max()).float().mean().item())
for n, p in b.named_parameters():
    if p.grad is not None:
        d = pa[n] - p.grad
        print(n, "relL2", (d.norm() / pa[n].norm().clamp_min(1e-12)).item())
# same comparison torch vs torch with a 1e-6 perturbation of the inputs: how sensitive are the gradients?
srcs2 = [t + 1e-6 * torch.randn_like(t) for t in srcs]
ss = [t.clone().requires_grad_(True) for t in srcs2]; a.zero_grad(); y2 = a(ss, pos)[0]; (y2 * cot).sum().backward()
for i, (u, v) in enumerate(zip(ga, ss)):
    d = u - v.grad
    print("torch vs torch(+1e-6 noise) src", i, "relL2", (d.norm() / u.norm()).item(), "max", d.abs().max().item())
