"""Pins the MAGI oracle's restatements of the reference's third-party CUDA kernels to the REAL kernels, on the GPU.

The MAGI reference calls five CUDA-only library kernels (dit_module.py:20-30,241-292,453,910,1000-1014): flash_attn's
`flash_attn_func` / `flash_attn_varlen_func` and `apply_rotary_emb`, flashinfer's `silu_and_mul` and `bmm_fp8`, and its
own Triton kernel `range_mod_triton`.  None of them runs on the CPU host where the goldens are generated, so
oracle/magi_oracle.py restates them from their published semantics and the reference-vs-oracle goldens could not cover
them (round-1 verdict: "that boundary is unpinned").  flash_attn, flashinfer and triton ARE in the GPU image, so this
file closes the loop there: every restatement is compared with the library kernel itself on the same inputs.

The file sorts last on purpose and every library call happens in a child process under a timeout: flashinfer and
triton JIT-compile on first use, and a library problem must read as a skip, never as a red (or hung) suite.
`range_mod_triton` is taken from the installed reference (baseline/_ref, tools/install_reference.sh) when present.
"""
import json
import subprocess
import sys
import textwrap
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent

PRELUDE = """
import json, math, sys, torch
sys.path.insert(0, %r)
from oracle import magi_oracle as mo
dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(7)
def rel_l2(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / b.norm()).item()
def same_frac(a, b):
    return (a.cpu() == b.cpu()).float().mean().item()
""" % str(ROOT)


def run_child(body: str, timeout: int = 240) -> dict:
    code = PRELUDE + textwrap.dedent(body)
    try:
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=timeout, cwd=str(ROOT))
    except subprocess.TimeoutExpired:
        pytest.skip(f"library kernel did not come up within {timeout} s (JIT compile?)")
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    if r.returncode != 0 or not lines:
        pytest.skip("library kernel unavailable on this box: " + (r.stderr or r.stdout).strip()[-300:])
    out = json.loads(lines[-1])
    print(out)
    return out


def test_flash_attn_func_is_what_gqa_attention_restates():
    """flash_attn_func (dit_module.py:1000-1014 branch) with 8 query / 2 KV heads vs mo.gqa_attention (fp32)."""
    out = run_child("""
        from flash_attn.flash_attn_interface import flash_attn_func
        q = torch.randn(1, 300, 8, 128, generator=g).bfloat16()
        k = torch.randn(1, 700, 2, 128, generator=g).bfloat16()
        v = torch.randn(1, 700, 2, 128, generator=g).bfloat16()
        lib = flash_attn_func(q.to(dev), k.to(dev), v.to(dev), deterministic=False)[0]
        print(json.dumps({"rel_l2": rel_l2(lib, mo.gqa_attention(q[0], k[0], v[0]))}))
    """)
    assert out["rel_l2"] <= 4e-3          # a bf16-P flash kernel against the fp32 statement


def test_flash_attn_varlen_func_is_what_varlen_attention_restates():
    """flash_attn_varlen_func (cross-attention, dit_module.py:960-996): segment i of q attends segment i of k / v."""
    out = run_child("""
        from flash_attn import flash_attn_varlen_func
        cu_q, cu_k = [0, 128, 256, 448], [0, 40, 75, 200]
        q = torch.randn(448, 8, 128, generator=g).bfloat16()
        k = torch.randn(200, 8, 128, generator=g).bfloat16()
        v = torch.randn(200, 8, 128, generator=g).bfloat16()
        lib = flash_attn_varlen_func(q.to(dev), k.to(dev), v.to(dev),
                                     torch.tensor(cu_q, dtype=torch.int32, device=dev),
                                     torch.tensor(cu_k, dtype=torch.int32, device=dev), 192, 125, deterministic=False)
        print(json.dumps({"rel_l2": rel_l2(lib, mo.varlen_attention(q, k, v, cu_q, cu_k))}))
    """)
    assert out["rel_l2"] <= 4e-3


def test_flash_attn_rotary_is_what_apply_rotary_restates():
    """flash_attn.layers.rotary.apply_rotary_emb (Triton; dit_module.py:910,927), non-interleaved, partial rotary dim."""
    out = run_child("""
        from flash_attn.layers.rotary import apply_rotary_emb
        x = torch.randn(1, 333, 6, 128, generator=g).bfloat16()
        ang = torch.randn(333, 32, generator=g)
        cos, sin = ang.cos(), ang.sin()
        lib = apply_rotary_emb(x.to(dev), cos.to(dev), sin.to(dev))
        ours = mo.apply_rotary(x.float(), cos, sin).bfloat16()
        print(json.dumps({"rel_l2": rel_l2(lib, ours), "same": same_frac(lib, ours)}))
    """)
    assert out["rel_l2"] <= 2e-3 and out["same"] >= 0.98     # fp32 rotate, one bf16 rounding on either side


def test_flashinfer_silu_and_mul_is_what_the_oracle_restates():
    """flashinfer.activation.silu_and_mul (gated MLP of the 24B model): silu(x[..., :d]) * x[..., d:]."""
    out = run_child("""
        import flashinfer
        x = (torch.randn(500, 2 * 1024, generator=g) * 2).bfloat16()
        lib = flashinfer.activation.silu_and_mul(x.to(dev))
        ours = mo.silu_and_mul(x)
        print(json.dumps({"rel_l2": rel_l2(lib, ours), "same": same_frac(lib, ours)}))
    """)
    assert out["rel_l2"] <= 2e-3 and out["same"] >= 0.98


def test_flashinfer_bmm_fp8_is_what_the_oracle_restates():
    """flashinfer.gemm.bmm_fp8 as PerTensorQuantizedFp8Linear calls it (dit_module.py:447-459): e4m3 x e4m3 (column-major
    weight), ONE scale per operand read through the scale pointers — including the reference's quirk of passing the
    per-channel `input_scale` vector, of which cuBLASLt reads element 0."""
    out = run_child("""
        from flashinfer.gemm import bmm_fp8
        m, k, n = 384, 512, 256
        a = (torch.randn(m, k, generator=g) * 0.5).to(torch.float8_e4m3fn)
        w = (torch.randn(n, k, generator=g) * 0.5).to(torch.float8_e4m3fn)              # nn.Linear layout [out, in]
        a_scale = torch.rand(k, generator=g) * 0.05 + 0.01                               # a VECTOR, like input_scale
        b_scale = torch.rand(1, generator=g) * 0.05 + 0.01
        lib = bmm_fp8(a.to(dev).reshape(1, m, k), w.to(dev).reshape(1, n, k).transpose(-2, -1),
                      a_scale.to(dev), b_scale.to(dev), dtype=torch.bfloat16)[0]
        ours = mo.bmm_fp8(a, w, a_scale, b_scale)
        print(json.dumps({"rel_l2": rel_l2(lib, ours), "same": same_frac(lib, ours)}))
    """)
    assert out["rel_l2"] <= 1e-3 and out["same"] >= 0.98


def test_reference_range_mod_triton_is_what_range_mod_restates(tmp_path):
    """The reference's own Triton kernel (dit_module.py:205-292), taken from the installed reference: its two
    definitions are cut out of baseline/_ref/.../dit_module.py into a scratch module (the module itself imports half
    the framework) and run on the GPU against mo.range_mod."""
    src_file = ROOT / "baseline" / "_ref" / "inferix" / "models" / "magi" / "dit" / "dit_module.py"
    if not src_file.exists():
        pytest.skip("baseline/_ref not installed (tools/install_reference.sh)")
    import ast
    src = src_file.read_text()
    tree = ast.parse(src)
    lines = src.splitlines()
    parts = []
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in ("range_mod_kernel_fwd", "range_mod_triton"):
            first = min([node.lineno] + [d.lineno for d in node.decorator_list])
            parts.append("\n".join(lines[first - 1:node.end_lineno]))
    if len(parts) != 2:
        pytest.skip("range_mod_triton not found in the installed reference")
    mod = tmp_path / "ref_range_mod.py"
    mod.write_text("import torch\nimport triton\nimport triton.language as tl\n\n\n" + "\n\n\n".join(parts) + "\n")
    out = run_child(f"""
        sys.path.insert(0, {str(tmp_path)!r})
        from ref_range_mod import range_mod_triton
        s, b, h, ranges = 640, 1, 3072, 4
        x = torch.randn(s, b, h, generator=g).bfloat16()
        cmap = torch.repeat_interleave(torch.arange(ranges), s // ranges).reshape(b, s).transpose(0, 1).contiguous()
        gat = torch.randn(b, ranges, h, generator=g).bfloat16()
        lib = range_mod_triton(x.to(dev), cmap.to(dev), gat.to(dev))
        ours = mo.range_mod(x, cmap, gat)
        xf, gf = x.float(), gat.float()
        lib32 = range_mod_triton(xf.to(dev), cmap.to(dev), gf.to(dev))
        print(json.dumps({{"rel_l2": rel_l2(lib, ours), "same": same_frac(lib, ours),
                          "fp32_equal": bool(torch.equal(lib32.cpu(), mo.range_mod(xf, cmap, gf)))}}))
    """)
    assert out["fp32_equal"]                                   # the form bias_modulate_add uses (fp32 in, :295-313)
    assert out["rel_l2"] <= 2e-3 and out["same"] >= 0.98


def test_native_pipeline_vs_reference_running_on_this_gpu():
    """The strongest pin available: the UNMODIFIED reference (baseline/_ref) running its own GPU path — flash-attn 2 +
    cuBLAS — on this GPU, against the native pipeline on the same synthetic weights, noise and re-noise stream
    (tools/ref_gpu_bench.py --parity: 4 blocks x 4 steps + clean pass, eviction from block 3).  Two bf16
    implementations: the bar is the one the CPU-reference goldens use (tests/test_gpu_pipeline.py)."""
    if not (ROOT / "baseline" / "_ref" / "inferix").exists():
        pytest.skip("baseline/_ref not installed (tools/install_reference.sh)")
    try:
        r = subprocess.run([sys.executable, str(ROOT / "tools" / "ref_gpu_bench.py"), "--parity"], capture_output=True,
                           text=True, timeout=300, cwd=str(ROOT))
    except subprocess.TimeoutExpired:
        pytest.skip("reference GPU run did not finish in 300 s")
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    if r.returncode != 0 or not lines:
        pytest.skip("reference GPU path unavailable on this box: " + (r.stderr or r.stdout).strip()[-300:])
    out = json.loads(lines[-1])
    print(out)
    assert out["finite"] and out["index_equal"]
    assert out["rel_l2"] <= 5e-3
