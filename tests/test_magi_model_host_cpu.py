"""Host logic of the native MAGI-1 VideoDiTModel (inferix_b200/magi_model.py) on CPU against goldens from the
reference's own VideoDiTModel (oracle/make_golden_magi_model.py): parameter names / dtypes / shapes, the embedding
prologue and range bookkeeping, the epilogue, the batch-folding of the unconditional pass and the CFG dispatcher
(cfg_number 1 and 3).  Kernels are the CPU test doubles of tests/fake_magi_ops.py; the GPU tests run the real ones."""
import types

import pytest
import torch

import fake_magi_ops
from oracle import magi_oracle as mo


def rel_l2(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm()).item()


def build_model(g, cfg_number, device="cpu"):
    from inferix_b200 import magi_model
    mc = types.SimpleNamespace(model_name="tiny", params_dtype=torch.bfloat16, layernorm_epsilon=1e-6,
                               apply_layernorm_1p=False, **g["model"])
    rc = types.SimpleNamespace(cfg_number=cfg_number, chunk_width=g["chunk_width"],
                               cfg_t_range=[0, 0.0217, 0.1000, 0.3, 0.999], prev_chunk_scales=[1.5] * 5,
                               text_scales=[7.5] * 5)
    ec = types.SimpleNamespace(cp_strategy="none", cp_size=1, fp8_quant=False, kv_offload=False, distill=False)
    model = magi_model.VideoDiTModel(types.SimpleNamespace(model_config=mc, runtime_config=rc, engine_config=ec))
    model.load_state_dict(mo.synth_model_state_dict(model, seed=g["seed"]), strict=True)
    return model.eval().to(device)


def new_ip(g):
    return types.SimpleNamespace(max_sequence_length=6 * g["clip"], max_batch_size=1, update_kv_cache=False)


def test_state_dict_matches_reference(golden_dir):
    g = torch.load(golden_dir / "magi_model.pt")
    model = build_model(g, 1)
    mine = {k: (str(v.dtype), tuple(v.shape)) for k, v in model.state_dict().items()}
    assert mine == g["state"]


def test_rotary_table_and_meta(golden_dir):
    """Prologue pieces that have closed forms: rope tail slicing, condition_map numbering, cumulative ranges."""
    g = torch.load(golden_dir / "magi_model.pt")
    model = build_model(g, 1)
    st = g["forward"][1]
    x, cond, cmap, y_flat, rope, meta = model.forward_pre_process(
        st["x"], st["t"], st["y"], torch.tensor([False]), st["mask"], st["kv_range"], **dict(st["kwargs"]))
    clip, r = g["clip"], st["kwargs"]["denoising_range_num"]
    assert x.shape == (r * clip, 1, 512) and x.dtype == torch.bfloat16
    assert rope.shape == (r * clip, 96) and rope.dtype == torch.float32
    assert cmap[:, 0].tolist() == [i // clip for i in range(r * clip)]
    assert meta.core_attn_params.np_q_range.tolist() == [[i * clip, (i + 1) * clip] for i in range(r)]
    assert meta.cross_attn_params.cu_seqlens_kv.tolist() == [0, 5, 17]
    assert y_flat.shape == (17, 512) and meta.clip_token_nums == clip and meta.slice_point == 1
    # the rotary table of the current frames is the tail of the table over history + current frames
    t_tot = (1 + r) * g["chunk_width"]
    rescale = (6 * 6 / 256) ** 0.5
    assert torch.equal(rope, model.rope.get_embed([t_tot, 6, 6], [t_tot, 6 / rescale, 6 / rescale])[-r * clip:])


def test_forward_sequence_matches_reference(golden_dir, monkeypatch):
    from inferix_b200 import magi_layer
    fake_magi_ops.install(monkeypatch, magi_layer)
    g = torch.load(golden_dir / "magi_model.pt")
    model = build_model(g, 1)
    ip = new_ip(g)
    for i, st in enumerate(g["forward"]):
        ip.update_kv_cache = st["update"]
        out = model(st["x"], st["t"], st["y"], caption_dropout_mask=torch.tensor([False]), xattn_mask=st["mask"],
                    kv_range=st["kv_range"], inference_params=ip, **dict(st["kwargs"]))
        err = rel_l2(out, st["out"])
        print(f"forward {i}: rel-L2 {err:.2e}")
        assert out.shape == st["out"].shape and out.dtype == torch.float32 and err <= 6e-3


@pytest.mark.parametrize("idx", [0, 1, 2, 3])
def test_cfg_dispatcher_matches_reference(golden_dir, monkeypatch, idx):
    from inferix_b200 import magi_layer
    fake_magi_ops.install(monkeypatch, magi_layer)
    g = torch.load(golden_dir / "magi_model.pt")
    st = g["dispatch"][idx]
    model = build_model(g, st["cfg_number"])
    ip = new_ip(g)
    pf = st["prefix"]
    ip.update_kv_cache = True
    model(pf["x"], pf["t"], pf["y"], caption_dropout_mask=torch.tensor([False]), xattn_mask=pf["mask"],
          kv_range=torch.tensor([[0, g["clip"]]], dtype=torch.int32), inference_params=ip, range_num=1,
          denoising_range_num=1, slice_point=0, chunk_width=g["chunk_width"], num_steps=12, distill_interval=4,
          extract_prefix_video_feature=True, fwd_extra_1st_chunk=False)
    kw = dict(st["kwargs"])
    out = model.forward_dispatcher(st["x"].clone(), st["t"], st["y"], st["mask"], st["kv_range"], ip, **kw)
    err = rel_l2(out, st["out"])
    print(f"dispatch cfg={st['cfg_number']} {st['kwargs']}: rel-L2 {err:.2e}")
    assert out.shape == st["out"].shape and err <= 8e-3
    # the two CFG copies of the result are identical and the non-denoised prefix is passed through
    assert torch.equal(out[0], out[1])
