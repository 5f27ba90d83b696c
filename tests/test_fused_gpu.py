"""GPU tests of the fused-producer kernels: softmax over the L*P logits and
`ref + offset / (W, H)` inside the forward / backward kernels, against the unfused composition the
reference module performs (ops/modules/ms_deform_attn.py:105-121) evaluated in fp64 with the oracle."""
import numpy as np
import pytest
import torch

from conftest import FWD_ABS_TOL, GRAD_REL_TOL, load_golden, rel_err
from test_modules import build_small

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def make_case(pkg, levels, batch, heads, points, shared_ref, seed):
    gen = torch.Generator().manual_seed(seed)
    L = len(levels)
    S = sum(h * w for h, w in levels)
    value = torch.randn(batch, S, heads, 32, generator=gen)
    ref = pkg.modules.reference_points_for(levels, "cpu")                      # [1, S, L, 2]
    if not shared_ref:
        ref = (ref + (torch.rand(batch, S, L, 2, generator=gen) - 0.5) * 0.01).contiguous()
    off = torch.randn(batch, S, heads, L, points, 2, generator=gen) * 2.0 + \
        torch.rand(batch, S, heads, L, points, 2, generator=gen) - 0.5
    logits = torch.randn(batch, S, heads, L * points, generator=gen) * 1.5
    grad_out = torch.randn(batch, S, heads * 32, generator=gen)
    shapes, lsi = pkg.synthetic.level_tensors(levels)
    return dict(value=value, ref=ref.contiguous(), off=off, logits=logits, grad_out=grad_out,
                shapes=shapes, lsi=lsi, levels=levels)


def reference_composition(oracle, c):
    """The reference module's arithmetic around the op (ms_deform_attn.py:105-112) on the CPU: the
    location `ref + off / (W, H)` in fp32 exactly as the module (and the fused kernel) computes it --
    a 1-ulp difference in a location can move a point into another bilinear cell, where the location
    gradient is discontinuous -- the softmax in fp64, the op in the fp64 oracle with fp32 geometry,
    and the chain rule back to offsets / logits through torch autograd."""
    L = len(c["levels"])
    N, S, M = c["off"].shape[:3]
    P = c["off"].shape[4]
    off = c["off"].clone().requires_grad_(True)
    logits = c["logits"].double().requires_grad_(True)
    w = torch.softmax(logits, -1).view(N, S, M, L, P)
    wh = c["shapes"].flip(-1).float()
    loc = c["ref"][:, :, None, :, None, :] + off / wh[None, None, None, :, None, :]      # fp32, IEEE
    a = (c["value"], c["shapes"], c["lsi"], loc.detach(), w.detach())
    out = oracle.forward(*a, geometry=np.float32)
    gv, gl, gw = oracle.backward(c["grad_out"], *a, geometry=np.float32)
    (loc.double() * torch.from_numpy(gl)).sum().backward()
    (w * torch.from_numpy(gw)).sum().backward()
    return out, gv, off.grad.double().numpy(), logits.grad.numpy()


@pytest.mark.parametrize("levels,batch,heads,points,shared_ref", [
    ([(8, 16), (16, 32), (32, 64)], 2, 8, 4, True),      # pixel-decoder shape, L*P = 12
    ([(12, 39), (24, 78)], 1, 4, 4, False),              # odd sizes, per-image reference points, L*P = 8
    ([(7, 9), (5, 3), (4, 4), (2, 2)], 2, 2, 4, True),   # L*P = 16
    ([(9, 13)], 3, 2, 4, True),                          # L*P = 4
])
@pytest.mark.parametrize("bwd_variant", [0, 20, 2])   # probe-gated default, merging kernel forced, per-row kernel forced
def test_fused_matches_unfused_composition(pkg, oracle, levels, batch, heads, points, shared_ref, bwd_variant):
    pkg.set_option("bwd_variant", bwd_variant)
    try:
        _fused_matches_unfused_composition(pkg, oracle, levels, batch, heads, points, shared_ref,
                                           launches=3 if bwd_variant == 0 else 2)
    finally:
        pkg.set_option("bwd_variant", 0)


def _fused_matches_unfused_composition(pkg, oracle, levels, batch, heads, points, shared_ref, launches):
    c = make_case(pkg, levels, batch, heads, points, shared_ref, seed=21)
    d = {k: (v.to(DEV) if torch.is_tensor(v) else v) for k, v in c.items()}
    value = d["value"].clone().requires_grad_(True)
    off = d["off"].clone().requires_grad_(True)
    logits = d["logits"].clone().requires_grad_(True)
    n0 = pkg.launch_count()
    out = pkg.MSDeformAttnFusedFunction.apply(value, d["shapes"], d["lsi"], d["ref"], off, logits)
    out.backward(d["grad_out"])
    torch.cuda.synchronize()
    assert pkg.launch_count() - n0 == launches   # forward + backward (default: merging + per-row kernel behind the probe)
    ref_out, ref_gv, ref_goff, ref_glog = reference_composition(oracle, c)
    assert np.abs(out.detach().double().cpu().numpy() - ref_out).max() <= 2 * FWD_ABS_TOL
    assert rel_err(value.grad.cpu().numpy(), ref_gv.reshape(value.shape)) <= GRAD_REL_TOL
    assert rel_err(off.grad.cpu().numpy(), ref_goff) <= GRAD_REL_TOL
    assert rel_err(logits.grad.cpu().numpy(), ref_glog) <= GRAD_REL_TOL

    # and against the unfused kernels fed by torch's own fp32 softmax / location arithmetic
    v2 = d["value"].clone().requires_grad_(True)
    o2 = d["off"].clone().requires_grad_(True)
    l2 = d["logits"].clone().requires_grad_(True)
    L, P = len(levels), points
    w = torch.softmax(l2, -1).view(*l2.shape[:3], L, P)
    wh = d["shapes"].flip(-1).float()
    loc = d["ref"][:, :, None, :, None, :] + o2 / wh[None, None, None, :, None, :]
    out2 = pkg.MSDeformAttnFunction.apply(v2, d["shapes"], d["lsi"], loc.contiguous(), w.contiguous(), 128)
    out2.backward(d["grad_out"])
    assert (out - out2).abs().max().item() <= 2e-5
    assert rel_err(value.grad.cpu().numpy(), v2.grad.cpu().numpy()) <= 1e-5
    assert rel_err(logits.grad.cpu().numpy(), l2.grad.cpu().numpy()) <= 1e-4
    assert rel_err(off.grad.cpu().numpy(), o2.grad.cpu().numpy()) <= 1e-4


def test_fused_encoder_matches_reference_golden(pkg):
    """Reference encoder (fp64, build container) vs mirror with fused=True on the sm_100a kernels."""
    m, g = build_small(pkg, dtype=torch.float32, device=DEV)
    for layer in m.encoder.layers:
        layer.self_attn.fused = True
    srcs = [torch.from_numpy(g[f"src{i}"]).float().cuda() for i in range(3)]
    srcs[2].requires_grad_(True)
    pos = [torch.from_numpy(g[f"pos{i}"]).float().cuda() for i in range(3)]
    n0 = pkg.launch_count()
    memory = m(srcs, pos)[0]
    assert np.abs(memory.detach().double().cpu().numpy() - g["memory"]).max() <= 2e-4
    (memory * torch.from_numpy(g["cotangent"]).float().cuda()).sum().backward()
    assert pkg.launch_count() - n0 == 6           # 2 layers x (fused forward + the two gated fused backward kernels)
    ref = g["grad_src2"]
    assert np.abs(srcs[2].grad.double().cpu().numpy() - ref).max() <= 2e-4 * max(1.0, np.abs(ref).max())


def test_fused_rejects_unsupported_shapes_and_module_falls_back(pkg):
    c = make_case(pkg, [(6, 7), (3, 4)], 1, 2, 3, True, seed=3)          # L*P = 6: no fused kernel
    d = {k: (v.to(DEV) if torch.is_tensor(v) else v) for k, v in c.items()}
    with pytest.raises(RuntimeError, match="fused MSDA covers"):
        pkg.ms_deform_attn_fused_forward(d["value"], d["shapes"], d["lsi"], d["ref"], d["off"], d["logits"])
    attn = pkg.modules.MSDeformAttn(64, 2, 2, 3, fused=True).to(DEV)         # same shape through the module
    q = torch.randn(1, 54, 64, device=DEV)
    ref = pkg.modules.reference_points_for([(6, 7), (3, 4)], DEV)
    attn_unfused = pkg.modules.MSDeformAttn(64, 2, 2, 3).to(DEV)
    attn_unfused.load_state_dict(attn.state_dict())
    with torch.no_grad():
        a = attn(q, ref, q, d["shapes"], d["lsi"])
        b = attn_unfused(q, ref, q, d["shapes"], d["lsi"])
    assert torch.equal(a, b)


def test_packed_projection_forward_is_bit_identical_to_separate_tensors(pkg):
    """msda_b200_fused_forward_strided_f32: offsets and logits read in place from one [N, Lq, 3*M*L*P]
    projection output give exactly the result of the packed tensors."""
    levels = [(16, 32), (8, 16), (4, 8)]
    c = make_case(pkg, levels, 2, 8, 4, True, seed=11)
    d = {k: (v.to(DEV) if torch.is_tensor(v) else v) for k, v in c.items()}
    N, Lq = d["off"].shape[:2]
    proj = torch.cat((d["off"].reshape(N, Lq, -1), d["logits"].reshape(N, Lq, -1)), -1).contiguous()
    want = pkg.ms_deform_attn_fused_forward(d["value"], d["shapes"], d["lsi"], d["ref"], d["off"], d["logits"])
    n0 = pkg.launch_count()
    got = pkg.ops.ms_deform_attn_fused_forward_packed(d["value"], d["shapes"], d["lsi"], d["ref"], proj,
                                                      len(levels), 4)
    assert pkg.launch_count() - n0 == 1
    assert torch.equal(got, want)
    # a per-query table added inside the kernel == the same table added to every image beforehand
    table = torch.randn(Lq, proj.shape[-1], device=DEV) * 0.5
    shifted = proj + table[None]
    want_t = pkg.ops.ms_deform_attn_fused_forward_packed(d["value"], d["shapes"], d["lsi"], d["ref"],
                                                         shifted.contiguous(), len(levels), 4)
    got_t = pkg.ops.ms_deform_attn_fused_forward_packed(d["value"], d["shapes"], d["lsi"], d["ref"], proj,
                                                        len(levels), 4, query_table=table)
    assert torch.equal(got_t, want_t)
    with pytest.raises(RuntimeError, match="query_table must be"):
        pkg.ops.ms_deform_attn_fused_forward_packed(d["value"], d["shapes"], d["lsi"], d["ref"], proj,
                                                    len(levels), 4, query_table=table[:-1])
    with pytest.raises(RuntimeError, match="projections must be"):
        pkg.ops.ms_deform_attn_fused_forward_packed(d["value"], d["shapes"], d["lsi"], d["ref"], proj[..., :-1],
                                                    len(levels), 4)
    # the C ABI: odd offsets stride / too narrow rows are shape errors
    lib = pkg._lib.lib
    args = lambda so, sl: (d["value"].data_ptr(), d["shapes"].data_ptr(), d["lsi"].data_ptr(), d["ref"].data_ptr(),
                           0 if d["ref"].shape[0] == 1 and N > 1 else Lq * 3 * 2, proj.data_ptr(), so,
                           proj.data_ptr(), sl, None, None, N, d["value"].shape[1], 8, 32, 3, Lq, 4,
                           want.data_ptr(), None)
    assert lib.msda_b200_fused_forward_strided_f32(*args(289, 288)) == -2
    assert lib.msda_b200_fused_forward_strided_f32(*args(190, 288)) == -2
    assert lib.msda_b200_fused_forward_strided_f32(*args(288, 95)) == -2
    one_table = list(args(288, 288))
    one_table[9] = table.data_ptr()                       # offsets table without a logits table
    assert lib.msda_b200_fused_forward_strided_f32(*one_table) == -1


def test_encoder_with_shared_position_embedding_folds_src_plus_pos_into_the_query_gemm(pkg):
    """Inference kernels (fused=True, linear="tf32x3") with a position embedding shared by the batch (batch
    stride 0, what the pixel decoder passes): `src + pos` is never formed -- the two query projections run as one
    GEMM on `src`, and the fused kernel adds the cached `pos W^T + b` row of each query -- and the result agrees with the
    same weights on the unfolded path (per-image copy of the embedding) and with torch's fp32 GEMMs."""
    torch.manual_seed(21)
    levels = [(8, 16), (16, 32), (32, 64)]
    kw = dict(d_model=256, nhead=8, num_encoder_layers=2, dim_feedforward=1024, dropout=0.0,
              num_feature_levels=3, enc_n_points=4)
    a = pkg.modules.MSDeformAttnTransformerEncoderOnly(**kw).to(DEV).eval()
    for m in a.modules():
        if isinstance(m, pkg.modules.MSDeformAttn):
            torch.nn.init.normal_(m.sampling_offsets.weight, std=0.02)
            torch.nn.init.normal_(m.attention_weights.weight, std=0.05)
            torch.nn.init.normal_(m.attention_weights.bias, std=0.05)
    b = pkg.modules.MSDeformAttnTransformerEncoderOnly(fused=True, linear="tf32x3", **kw).to(DEV).eval()
    b.load_state_dict(a.state_dict())
    srcs = [torch.randn(3, 256, h, w, device=DEV) for h, w in levels]
    one = [torch.randn(1, 256, h, w, device=DEV) for h, w in levels]
    shared = [p.expand(3, -1, -1, -1) for p in one]                  # batch stride 0
    copies = [p.expand(3, -1, -1, -1).contiguous() for p in one]     # same values, one copy per image
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        with torch.no_grad():
            want = a(srcs, copies)[0]
            n0 = pkg.launch_count()
            unfolded = b(srcs, copies)[0]
            n1 = pkg.launch_count()
            folded = b(srcs, shared)[0]
            n2 = pkg.launch_count()
            again = b(srcs, shared)[0]
            n3 = pkg.launch_count()
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old
    per_layer_unfolded = 1 + 6 * 2 + 2          # MSDA forward, 6 x (weight split + GEMM) on first use, 2 x add + LayerNorm
    per_layer_folded = 1 + 5 + 2                # offsets and logits are ONE GEMM; split weights are cached per layer
    assert n1 - n0 == 2 * per_layer_unfolded
    assert n2 - n1 == 2 * (per_layer_folded + 2)            # first call: + split of the stacked weight + the table GEMM
    assert n3 - n2 == 2 * per_layer_folded                  # afterwards table and split weights are cached
    assert torch.equal(folded, again)
    assert (folded - unfolded).abs().max().item() <= 2e-5, (folded - unfolded).abs().max().item()
    assert (folded - want).abs().max().item() <= 5e-5, (folded - want).abs().max().item()
    # the table follows its inputs: an in-place change of a projection weight or of the level embedding rebuilds it
    with torch.no_grad():
        b.encoder.layers[0].self_attn.attention_weights.bias.add_(0.25)
        b.level_embed.add_(0.1)
        a.load_state_dict(b.state_dict())
        want2 = a(srcs, copies)[0]
        folded2 = b(srcs, shared)[0]
    assert (folded2 - want2).abs().max().item() <= 5e-5
    assert (folded2 - folded).abs().max().item() > 1e-3


@pytest.mark.parametrize("levels,points,shared_ref,lq", [
    ([(12, 20), (6, 10)], 4, True, None),           # L*P = 8
    ([(16, 16), (8, 8), (4, 4), (2, 2)], 4, False, None),   # L*P = 16, per-image reference points
    ([(9, 13)], 4, True, None),                     # L*P = 4, odd sizes
    ([(16, 32), (8, 16), (4, 8)], 4, True, 301),    # fewer queries than pixels (decoder-style / a rank's row range)
])
def test_packed_projection_with_table_on_other_shapes(pkg, levels, points, shared_ref, lq):
    """Strided offsets / logits + per-query table against the packed tensors with the table added beforehand,
    over the level / point counts the fused kernels are instantiated for, per-image reference points and Lq != S."""
    c = make_case(pkg, levels, 2, 8, points, shared_ref, seed=5 + len(levels))
    d = {k: (v.to(DEV) if torch.is_tensor(v) else v) for k, v in c.items()}
    L = len(levels)
    if lq is not None:
        d["off"], d["logits"], d["ref"] = d["off"][:, :lq].contiguous(), d["logits"][:, :lq].contiguous(), d["ref"][:, :lq].contiguous()
    N, Lq = d["off"].shape[:2]
    proj = torch.cat((d["off"].reshape(N, Lq, -1), d["logits"].reshape(N, Lq, -1)), -1).contiguous()
    table = torch.randn(Lq, proj.shape[-1], device=DEV) * 0.3
    shifted = (proj + table[None]).contiguous()
    width = 2 * 8 * L * points
    want = pkg.ms_deform_attn_fused_forward(d["value"], d["shapes"], d["lsi"], d["ref"],
                                            shifted[..., :width].reshape(N, Lq, 8, L, points, 2).contiguous(),
                                            shifted[..., width:].reshape(N, Lq, 8, L * points).contiguous())
    got = pkg.ops.ms_deform_attn_fused_forward_packed(d["value"], d["shapes"], d["lsi"], d["ref"], proj, L, points,
                                                      query_table=table)
    assert torch.equal(got, want)
