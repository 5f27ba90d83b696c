"""Generate tests/golden/*.npz by running the REFERENCE's own CPU path.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

It imports ``ms_deform_attn_core_pytorch`` from
/root/reference/model/modeling/pixel_decoder/ops/functions/ms_deform_attn_func.py (by file
path, so ``model/__init__.py`` and its detectron2 import never run), evaluates it in
fp64 on fp32-representable seeded inputs, takes the three gradients with autograd, and
stores inputs + outputs.  The committed fixtures are what pins oracle/ (and through it
the CUDA path) to the reference; nothing on the GPU box reads /root/reference.

Cases cover SURVEY.md section 8(c): model-like and uniform locations, all points out of
bounds, exact borders / pixel centres / half-pixel borders, single level / single point,
D in {4, 8, 16, 32, 64}, odd and non-square levels such as (12, 39), Lq != S.
"""
import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF_FUNC = "/root/reference/model/modeling/pixel_decoder/ops/functions/ms_deform_attn_func.py"

sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402


def load_reference():
    spec = importlib.util.spec_from_file_location("ref_ms_deform_attn_func", REF_FUNC)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.ms_deform_attn_core_pytorch


def border_locations(levels, batch, heads, points, Lq):
    """Locations exactly at 0, 1, -1/W, 1+1/W, pixel centres, and the half-pixel borders."""
    L = len(levels)
    loc = torch.zeros(batch, Lq, heads, L, points, 2)
    for l, (H, W) in enumerate(levels):
        xs = [0.0, 1.0, -1.0 / W, 1.0 + 1.0 / W, 0.5 / W, (W - 0.5) / W, 0.25 / W, (W - 0.25) / W,
              1.5 / W, 0.5, -0.49 / W, 1.0 + 0.49 / W]
        ys = [0.0, 1.0, -1.0 / H, 1.0 + 1.0 / H, 0.5 / H, (H - 0.5) / H, 0.25 / H, (H - 0.25) / H,
              1.5 / H, 0.5, -0.49 / H, 1.0 + 0.49 / H]
        k = 0
        for q in range(Lq):
            for m in range(heads):
                for p in range(points):
                    loc[:, q, m, l, p, 0] = xs[k % len(xs)]
                    loc[:, q, m, l, p, 1] = ys[(k // len(xs) + k) % len(ys)]
                    k += 1
    return loc


def cases(syn):
    out = {}
    pyr = [(2, 3), (4, 6), (8, 12)]
    out["model_d32"] = syn.make_inputs(pyr, batch=1, mode="model", seed=1)
    out["model_d32_b2_m2"] = syn.make_inputs(pyr, batch=2, heads=2, mode="model", seed=8)
    out["uniform_d32"] = syn.make_inputs(pyr, batch=1, num_query=16, mode="uniform", seed=2)
    out["kitti_like_odd"] = syn.make_inputs([(2, 3), (3, 5), (6, 13)], batch=1, heads=2, mode="model", seed=3)
    for D in (4, 8, 16, 64):
        out[f"uniform_d{D}"] = syn.make_inputs([(5, 7), (9, 4)], batch=2, heads=2, channels=D,
                                               points=2, num_query=19, mode="uniform", seed=10 + D)
    out["single_level_point"] = syn.make_inputs([(7, 9)], batch=1, heads=3, channels=32, points=1,
                                                num_query=23, mode="uniform", seed=4)
    out["four_levels"] = syn.make_inputs([(1, 2), (2, 3), (4, 6), (8, 12)], batch=1, heads=2,
                                         mode="model", seed=5)
    # every point out of bounds -> output and gradients exactly zero
    oob = syn.make_inputs(pyr, batch=1, heads=2, num_query=11, mode="uniform", seed=6)
    oob["sampling_locations"] = oob["sampling_locations"] + 3.0
    out["all_oob"] = oob
    # exact borders / centres / half-pixel borders (forward is continuous there; gradients wrt
    # location are not, so only forward, grad_value and grad_attn_weight are pinned)
    b = syn.make_inputs(pyr, batch=1, heads=2, num_query=29, mode="uniform", seed=7)
    b["sampling_locations"] = border_locations(pyr, 1, 2, 4, 29)
    out["borders"] = b
    return out


def encoder_golden():
    """A small reference MSDeformAttnTransformerEncoderOnly (msdeformattn.py:26-99) in fp64 on CPU:
    weights, inputs, memory and the gradient of <memory, cotangent> w.r.t. the finest-level input.
    Pins the host-side mirror in uni-encoder-code_b200/modules.py (and, on the GPU, mirror + op)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import ref_import
    ns = ref_import.load()
    torch.manual_seed(1234)
    levels = [(2, 3), (4, 6), (8, 12)]
    enc = ns.MSDeformAttnTransformerEncoderOnly(d_model=64, nhead=2, num_encoder_layers=2,
                                                dim_feedforward=128, dropout=0.0,
                                                num_feature_levels=3, enc_n_points=4).double().eval()
    # default init puts every offset on an integer pixel step (ms_deform_attn.py:69-77), where
    # the gradient w.r.t. locations is discontinuous: randomise the offset projection
    gen = torch.Generator().manual_seed(99)
    with torch.no_grad():
        for layer in enc.encoder.layers:
            layer.self_attn.sampling_offsets.weight.copy_(
                torch.randn(layer.self_attn.sampling_offsets.weight.shape, generator=gen, dtype=torch.float64) * 0.05)
            layer.self_attn.sampling_offsets.bias.add_(
                torch.rand(layer.self_attn.sampling_offsets.bias.shape, generator=gen, dtype=torch.float64) * 0.6 + 0.2)
            layer.self_attn.attention_weights.weight.copy_(
                torch.randn(layer.self_attn.attention_weights.weight.shape, generator=gen, dtype=torch.float64) * 0.1)
    srcs = [torch.randn(2, 64, h, w, generator=gen, dtype=torch.float64) for h, w in levels]
    srcs[2].requires_grad_(True)
    pe = ns.PositionEmbeddingSine(32, normalize=True)
    pos = [pe(s).double() for s in srcs]
    memory, shapes, lsi, _ = enc(srcs, pos)
    cot = torch.randn(memory.shape, generator=gen, dtype=torch.float64)
    (memory * cot).sum().backward()
    out = {"state::" + k: v.detach().numpy() for k, v in enc.state_dict().items()}
    for i, (s_, p_) in enumerate(zip(srcs, pos)):
        out[f"src{i}"] = s_.detach().numpy()
        out[f"pos{i}"] = p_.detach().numpy()
    out.update(memory=memory.detach().numpy(), cotangent=cot.numpy(), grad_src2=srcs[2].grad.numpy(),
               spatial_shapes=shapes.numpy(), level_start_index=lsi.numpy())
    path = os.path.join(HERE, "encoder_small.npz")
    np.savez_compressed(path, **out)
    print(f"encoder_small: memory {tuple(memory.shape)} -> {os.path.getsize(path) / 1024:.1f} KiB")


PD_KW = dict(transformer_dropout=0.0, transformer_nheads=2, transformer_dim_feedforward=96,
             transformer_enc_layers=2, conv_dim=64, mask_dim=32, norm="GN",
             transformer_in_features=["res3", "res4", "res5"], common_stride=4)
PD_SHAPES = {"res2": (12, 4), "res3": (20, 8), "res4": (28, 16), "res5": (36, 32)}   # (channels, stride)


def pixel_decoder_golden():
    """A small reference MSDeformAttnPixelDecoder (msdeformattn.py:178-386, with the third-party stubs of
    tests/ref_stubs) in fp32 (its only mode: inputs are cast with .float()) on CPU for a 64x96 image: weights, the four backbone features and the three
    outputs of forward_features.  Pins uni-encoder-code_b200/pixel_decoder.py."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import ref_import
    ns = ref_import.load()
    torch.manual_seed(4321)
    shapes = {k: ns.ShapeSpec(channels=c, stride=s) for k, (c, s) in PD_SHAPES.items()}
    dec = ns.MSDeformAttnPixelDecoder(shapes, **PD_KW).eval()      # fp32: forward_features casts its inputs with .float()
    gen = torch.Generator().manual_seed(77)
    with torch.no_grad():
        for layer in dec.transformer.encoder.layers:
            a = layer.self_attn
            a.sampling_offsets.weight.copy_(torch.randn(a.sampling_offsets.weight.shape, generator=gen) * 0.05)
            a.sampling_offsets.bias.add_(torch.rand(a.sampling_offsets.bias.shape, generator=gen) * 0.6 + 0.2)
            a.attention_weights.weight.copy_(torch.randn(a.attention_weights.weight.shape, generator=gen) * 0.1)
    feats = {k: torch.randn(2, c, 64 // s, 96 // s, generator=gen) for k, (c, s) in PD_SHAPES.items()}
    with torch.no_grad():
        mask_feat, low, multi = dec.forward_features(feats)      # the reference casts inputs to fp32 (.float())
    out = {"state::" + k: v.detach().numpy() for k, v in dec.state_dict().items()}
    out.update({"feat::" + k: v.numpy() for k, v in feats.items()})
    out.update(mask_features=mask_feat.numpy(), lowest=low.numpy(),
               **{f"multi{i}": m.numpy() for i, m in enumerate(multi)})
    path = os.path.join(HERE, "pixel_decoder_small.npz")
    np.savez_compressed(path, **out)
    print(f"pixel_decoder_small: mask_features {tuple(mask_feat.shape)} -> {os.path.getsize(path) / 1024:.1f} KiB")


def main():
    encoder_golden()
    pixel_decoder_golden()
    core = load_reference()
    syn = load_package().synthetic if os.path.exists(
        os.path.join(ROOT, "uni-encoder-code_b200", "lib", "libmsda_b200.so")) else None
    if syn is None:
        raise SystemExit("build the package first (python __graft_entry__.py)")
    for name, inp in cases(syn).items():
        v = inp["value"].double().requires_grad_(True)
        loc = inp["sampling_locations"].double().requires_grad_(True)
        w = inp["attention_weights"].double().requires_grad_(True)
        out = core(v, inp["spatial_shapes"], loc, w)
        out.backward(inp["grad_output"].double())
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(
            path,
            value=inp["value"].numpy(), spatial_shapes=inp["spatial_shapes"].numpy(),
            level_start_index=inp["level_start_index"].numpy(),
            sampling_locations=inp["sampling_locations"].numpy(),
            attention_weights=inp["attention_weights"].numpy(),
            grad_output=inp["grad_output"].numpy(),
            ref_output=out.detach().numpy(), ref_grad_value=v.grad.numpy(),
            ref_grad_sampling_loc=loc.grad.numpy(), ref_grad_attn_weight=w.grad.numpy(),
            loc_grad_pinned=np.array(name != "borders"))
        print(f"{name}: out {tuple(out.shape)} -> {os.path.getsize(path) / 1024:.1f} KiB")


if __name__ == "__main__":
    main()
