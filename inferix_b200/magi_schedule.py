"""MAGI-1 sliding-window chunk scheduler — the index / timestep logic that feeds the kernels.

Pure host functions restating ``inferix/pipeline/magi/video_generate.py``: ``generate_sequences`` (:166-182),
``init_t`` (:185-231), ``init_intervel`` (:234-243), ``SampleTransport.get_timestep`` (:320-339),
``get_denoise_step_of_each_chunk`` (:341-360), the kv-range builders (:373-529) and
``generate_denoise_status_and_sequences`` / ``total_forward_step`` (:553-585).  Outputs are the reference's:
int32 ``kv_range`` rows ``[start_token, end_token)`` (bit-exact), float32 timestep tables.  The reference methods
read ``self.runtime_config`` / ``self.transport_inputs``; here those values are explicit arguments.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch

__all__ = ["generate_sequences", "init_t", "init_intervel", "get_timestep", "get_denoise_step_of_each_chunk",
           "kvrange_for_prefix_video", "default_kvrange", "noise2clean_kvrange", "kvrange_for_denoising_video",
           "denoise_status_and_sequences", "total_forward_step", "integrate"]


def generate_sequences(chunk_num: int, window_size: int, chunk_offset: int):
    """:166-182 — per denoising stage: chunks [clip_start, clip_end) and their window slots [t_start, t_end)."""
    start_index, end_index = chunk_offset, chunk_num + window_size - 1
    clip_start = [max(chunk_offset, i - window_size + 1) for i in range(start_index, end_index)]
    clip_end = [min(chunk_num, i + 1) for i in range(start_index, end_index)]
    t_start = [max(0, i - chunk_num + 1) for i in range(start_index, end_index)]
    t_end = [min(window_size, i - chunk_offset + 1) if i - chunk_offset < window_size else window_size
             for i in range(start_index, end_index)]
    return clip_start, clip_end, t_start, t_end


def init_t(t_schedule_config: Optional[Dict], num_steps: int, device="cpu", shortcut_mode: str = "") -> torch.Tensor:
    """:185-231 — timestep grid (0: noise, 1: clean) with the sd3 / square / piecewise transforms."""
    if num_steps == 12:
        base_t = torch.linspace(0, 1, 4 + 1, device=device) / 4
        accu_num = torch.linspace(0, 1, 4 + 1, device=device)
        base_t = base_t[:3] if shortcut_mode == "16,16,8" else torch.cat([base_t[:1], base_t[2:4]], dim=0)
        t = torch.cat([base_t + accu for accu in accu_num], dim=0)[: (num_steps + 1)]
    else:
        t = torch.linspace(0, 1, num_steps + 1, device=device)
    func = (t_schedule_config or {}).get("tSchedulerFunc", "sd3")
    if func == "sd3":
        shift = (t_schedule_config or {}).get("shift", 3.0)
        assert shift >= 1.0, "shift should >=1"
        shift_inv = 1.0 / shift
        t = t ** 2
        t = shift_inv * t / (1 + (shift_inv - 1) * t)
    elif func == "square":
        t = t ** 2
    elif func == "piecewise":
        mask = t < 0.875
        t[mask] = t[mask] * (0.5 / 0.875)
        t[~mask] = 0.5 + (t[~mask] - 0.875) * (0.5 / (1 - 0.875))
    return t


def init_intervel(num_steps: int, device="cpu", shortcut_mode: str = "") -> torch.Tensor:
    """:234-243."""
    base = torch.ones(num_steps, device=device)
    if num_steps % 3 == 0:
        pat = [1, 1, 2] if shortcut_mode == "16,16,8" else [2, 1, 1]
        base = torch.tensor(pat * (num_steps // 3), device=device)
    return base


def get_timestep(t_total: torch.Tensor, denoise_step_per_stage: int, start: int, end: int, denoise_idx: int,
                 has_clean_t: bool = False, clean_t: float = 1.0) -> torch.Tensor:
    """:320-339 — timesteps of the window's chunks, newest (noisiest) chunk last."""
    idx = [i * denoise_step_per_stage + denoise_idx for i in range(start, end)]
    idx.reverse()
    ts = t_total[idx]
    if has_clean_t:
        ts = torch.cat([torch.ones(1, device=t_total.device) * clean_t, ts], 0)
    return ts


def get_denoise_step_of_each_chunk(num_steps: int, denoise_step_per_stage: int, t_start: int, t_end: int,
                                   denoise_idx: int, has_clean_t: bool = False) -> List[int]:
    """:341-360."""
    steps = [i * denoise_step_per_stage + denoise_idx for i in range(t_start, t_end)]
    steps.reverse()
    return ([num_steps] + steps) if has_clean_t else steps


def kvrange_for_prefix_video(range_num: int, chunk_token_nums: int, clean_chunk_kvrange: int = -1,
                             noise2clean: Sequence[int] = (), batch_size: int = 1) -> torch.Tensor:
    """:373-391 — each prefix chunk attends the previous `prev_chunk_num` chunks."""
    if clean_chunk_kvrange != -1:
        prev = clean_chunk_kvrange
    elif len(noise2clean) > 0:
        prev = noise2clean[-1]
    else:
        prev = 8
    k_end = torch.linspace(1, range_num, steps=range_num).reshape((range_num, 1))
    k_start = torch.clamp(k_end - prev, min=0).reshape((range_num, 1))
    rng = torch.concat([k_start, k_end], dim=1)
    return torch.concat([rng + i * range_num for i in range(batch_size)], dim=0).to(torch.int32) * chunk_token_nums


def default_kvrange(slice_point: int, denoising_range_num: int, chunk_token_nums: int, batch_size: int = 1):
    """:455-467 — every denoising chunk attends everything from token 0 up to itself."""
    range_num = slice_point + denoising_range_num
    k_end = torch.linspace(slice_point + 1, range_num, steps=denoising_range_num).reshape((denoising_range_num, 1))
    k_start = torch.Tensor([0] * denoising_range_num).reshape((denoising_range_num, 1))
    rng = torch.concat([k_start, k_end], dim=1)
    return torch.concat([rng + i * range_num for i in range(batch_size)], dim=0).to(torch.int32) * chunk_token_nums


def noise2clean_kvrange(slice_point: int, denoising_range_num: int, chunk_token_nums: int, noise2clean: Sequence[int],
                        clean_chunk_kvrange: int, denoise_step_of_each_chunk: Sequence[int], num_steps: int,
                        batch_size: int = 1) -> torch.Tensor:
    """:469-510 — the noisier a chunk, the shorter the history it attends."""
    assert len(denoise_step_of_each_chunk) == denoising_range_num
    assert len(noise2clean) > 0
    if clean_chunk_kvrange == -1:
        clean_chunk_kvrange = noise2clean[-1]
    assert num_steps % len(noise2clean) == 0
    per_stage = num_steps // len(noise2clean)
    width = [clean_chunk_kvrange if s == num_steps else noise2clean[s // per_stage] for s in denoise_step_of_each_chunk]
    range_num = slice_point + denoising_range_num
    rows = []
    for i in range(batch_size):
        base = i * range_num
        for j in range(denoising_range_num):
            k_end = slice_point + j + 1
            k_start = max(0, k_end - width[j])
            rows.append(torch.Tensor([(base + k_start) * chunk_token_nums, (base + k_end) * chunk_token_nums]).reshape(1, 2))
    return torch.concat(rows, dim=0).to(torch.int32)


def kvrange_for_denoising_video(slice_point: int, denoising_range_num: int, chunk_token_nums: int,
                                denoise_step_of_each_chunk: Sequence[int], num_steps: int,
                                noise2clean: Sequence[int] = (), clean_chunk_kvrange: int = -1, batch_size: int = 1):
    """:512-529."""
    if len(noise2clean) == 0:
        return default_kvrange(slice_point, denoising_range_num, chunk_token_nums, batch_size)
    return noise2clean_kvrange(slice_point, denoising_range_num, chunk_token_nums, noise2clean, clean_chunk_kvrange,
                               denoise_step_of_each_chunk, num_steps, batch_size)


def denoise_status_and_sequences(cur_denoise_step: int, num_steps: int, chunk_num: int, window_size: int,
                                 chunk_offset: int = 0):
    """:553-574 — (denoise_step_per_stage, stage, idx), (chunk_offset, chunk_start, chunk_end, t_start, t_end)."""
    per_stage = num_steps // window_size
    stage, idx = cur_denoise_step // per_stage, cur_denoise_step % per_stage
    cs, ce, ts, te = generate_sequences(chunk_num, window_size, chunk_offset)
    return (per_stage, stage, idx), (chunk_offset, cs[stage], ce[stage], ts[stage], te[stage])


def total_forward_step(num_steps: int, chunk_num: int, window_size: int, chunk_offset: int = 0) -> int:
    """:576-585."""
    return (num_steps // window_size) * (chunk_num + window_size - 1 - chunk_offset)


def integrate(x_chunk: torch.Tensor, velocity: torch.Tensor, t_total: torch.Tensor, denoise_step_per_stage: int,
              t_start: int, t_end: int, i: int, chunk_width: int) -> torch.Tensor:
    """:531-551 — Euler step x += v * dt with a per-chunk dt."""
    dt = get_timestep(t_total, denoise_step_per_stage, t_start, t_end, i + 1) - \
        get_timestep(t_total, denoise_step_per_stage, t_start, t_end, i)
    n, c, t, h, w = x_chunk.shape
    x = x_chunk.reshape(n, c, -1, chunk_width, h, w)
    v = velocity.reshape(n, c, -1, chunk_width, h, w)
    assert x.size(2) == dt.size(0)
    return (x + v * dt.reshape(1, 1, -1, 1, 1, 1)).reshape(n, c, t, h, w)
