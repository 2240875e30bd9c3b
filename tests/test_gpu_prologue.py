"""Prologue / epilogue kernels of one DiT forward against the eager arithmetic of the reference
(the same torch ops the oracle's embed / head_unpatchify / flow_to_x0 / add_noise use): patch gather + GEMM vs Conv3d,
fp64 sinusoid, the small-row linears of the time MLP incl. the all-layers modulation table, unpatchify + fp64 x0, and
add_noise.  Element-wise kernels are bit-exact; the GEMV differs from cuBLAS only in fp32 summation order."""
import pytest
import torch
import torch.nn.functional as F

from inferix_b200 import ops
from inferix_b200.scheduler import FlowMatchScheduler
from inferix_b200.wan_model import sinusoidal_embedding_1d

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda", 0)


def rel_l2(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


@pytest.mark.parametrize("hw,dim,world,rank", [((16, 16), 256, 1, 0), ((90, 160), 1536, 1, 0), ((90, 160), 1536, 8, 3)])
def test_patch_embedding_equals_conv3d(hw, dim, world, rank):
    g = torch.Generator(device=DEV).manual_seed(0)
    lat = torch.randn(1, 3, 16, hw[0], hw[1], device=DEV, generator=g).bfloat16()       # [B, F, C, H, W] as the pipeline holds it
    x = lat.permute(0, 2, 1, 3, 4)[0]                                                   # [C, F, H, W] view, not contiguous
    w = (torch.randn(dim, 16, 1, 2, 2, device=DEV, generator=g) * 0.1).bfloat16()
    b = (torch.randn(dim, device=DEV, generator=g) * 0.1).bfloat16()
    ref = F.conv3d(x.unsqueeze(0), w, b, stride=(1, 2, 2)).flatten(2).transpose(1, 2)[0]   # [F*hw, dim]
    ghw = (hw[0] // 2) * (hw[1] // 2)
    chunk = ghw // world
    ref = ref.view(3, ghw, dim)[:, rank * chunk:(rank + 1) * chunk].reshape(-1, dim)
    a = ops.patchify(x, (1, 2, 2), rank * chunk, chunk)
    out = ops.gemm(a, w.view(dim, -1), b)
    # exact statement in fp32 (what the oracle's CPU conv3d accumulates in): ours must be one bf16 rounding away from
    # it, and at least as close as cuDNN's bf16 Conv3d (measured 2.8e-3 from ours: cuDNN is the less exact of the two)
    ref32 = F.conv3d(x.unsqueeze(0).float(), w.float(), b.float(), stride=(1, 2, 2)).flatten(2).transpose(1, 2)[0]
    ref32 = ref32.view(3, ghw, dim)[:, rank * chunk:(rank + 1) * chunk].reshape(-1, dim)
    ours, cudnn = rel_l2(out, ref32), rel_l2(ref, ref32)
    print(f"patch embedding vs fp32 conv3d: ours {ours:.2e}, cuDNN bf16 {cudnn:.2e}; ours vs cuDNN {rel_l2(out, ref):.2e}")
    assert out.shape == ref.shape and ours <= 2.5e-3 and ours <= cudnn + 2e-4
    assert torch.equal(a.view(3, chunk, 16, 2, 2)[1, 5, :, 1, 0],
                       x[:, 1, 2 * ((rank * chunk + 5) // (hw[1] // 2)) + 1, 2 * ((rank * chunk + 5) % (hw[1] // 2))])


def test_sinusoid_and_time_mlp():
    dim, freq, layers = 1536, 256, 30
    g = torch.Generator(device=DEV).manual_seed(1)
    t = torch.tensor([1000.0, 937.5, 0.0, 522.25], device=DEV)
    sin = ops.sinusoidal_embedding(t.to(torch.float64), freq)
    ref = sinusoidal_embedding_1d(freq, t).bfloat16()
    assert (sin == ref).float().mean().item() >= 0.995 and (sin.float() - ref.float()).abs().max().item() <= 8e-3

    def lin(o, i, s):
        return ((torch.randn(o, i, device=DEV, generator=g) * s).bfloat16(), (torch.randn(o, device=DEV, generator=g) * 0.1).bfloat16())
    (w0, b0), (w2, b2), (wp, bp) = lin(dim, freq, 0.05), lin(dim, dim, 0.03), lin(6 * dim, dim, 0.03)
    h1 = ops.linear_small(ref, w0, b0)
    r1 = F.linear(ref, w0, b0)
    assert rel_l2(h1, r1) <= 2e-3
    e = ops.linear_small(r1, w2, b2, silu_input=True)
    re = F.linear(F.silu(r1), w2, b2)
    assert rel_l2(e, re) <= 2e-3
    table = (torch.randn(layers, 6 * dim, device=DEV, generator=g) * 0.03).bfloat16()
    mods = ops.linear_small(re, wp, bp, silu_input=True, mod_table=table)
    e0 = F.linear(F.silu(re), wp, bp)
    rm = table.unsqueeze(1) + e0.unsqueeze(0)                                           # modulation + e0 per layer (:412)
    assert mods.shape == rm.shape and rel_l2(mods, rm) <= 2e-3


@pytest.mark.parametrize("hw", [(16, 16), (90, 160)])
def test_unpatchify_x0_bit_exact(hw):
    from inferix_b200.wrapper import WanDiffusionWrapper
    frames, c = 3, 16
    gh, gw = hw[0] // 2, hw[1] // 2
    g = torch.Generator(device=DEV).manual_seed(2)
    tokens = torch.randn(frames * gh * gw, 4 * c, device=DEV, generator=g).bfloat16()
    xt = torch.randn(1, frames, c, hw[0], hw[1], device=DEV, generator=g).bfloat16()
    sched = FlowMatchScheduler(shift=5.0, sigma_min=0.0, extra_one_step=True)
    sched.set_timesteps(1000, training=True)
    t = torch.tensor([sched.timesteps[100].item(), 500.0, sched.timesteps[900].item() + 0.3], device=DEV)
    flow, x0 = ops.unpatchify_x0(tokens, xt[0], t.to(torch.float64), sched.timesteps.to(DEV), sched.sigmas.to(DEV), (2, 2))
    # reference: unpatchify (causal_model.py:1196-1219) then wrapper._convert_flow_pred_to_x0 (:259-283)
    u = tokens.view(frames, gh, gw, 1, 2, 2, c)
    ref_flow = torch.einsum("fhwpqrc->cfphqwr", u).reshape(c, frames, hw[0], hw[1]).permute(1, 0, 2, 3)   # [F, C, H, W]
    w = WanDiffusionWrapper.__new__(WanDiffusionWrapper)
    w.scheduler = sched
    ref_x0 = WanDiffusionWrapper._convert_flow_pred_to_x0(w, ref_flow, xt[0], t)
    assert torch.equal(flow, ref_flow.contiguous())
    assert torch.equal(x0, ref_x0)


def test_add_noise_bit_exact():
    sched = FlowMatchScheduler(shift=5.0, sigma_min=0.0, extra_one_step=True)
    sched.set_timesteps(1000, training=True)
    g = torch.Generator(device=DEV).manual_seed(3)
    x0 = torch.randn(3, 16, 90, 160, device=DEV, generator=g).bfloat16()
    noise = torch.randn(3, 16, 90, 160, device=DEV, generator=g).bfloat16()
    for t in (torch.full((3,), 750, dtype=torch.long, device=DEV), torch.tensor([937.5, 522.0, 3.0], device=DEV)):
        out = sched.add_noise(x0, noise, t)                                            # native path (CUDA bf16)
        _, sigma = sched._sigma_of(t, DEV)
        ref = ((1 - sigma) * x0 + sigma * noise).type_as(noise)                        # the reference's eager arithmetic
        assert torch.equal(out, ref)
    t1 = torch.full((1,), 522, dtype=torch.long, device=DEV)                           # CausVid passes ONE timestep (:233-236)
    _, sigma = sched._sigma_of(t1, DEV)
    assert torch.equal(sched.add_noise(x0, noise, t1), ((1 - sigma) * x0 + sigma * noise).type_as(noise))
