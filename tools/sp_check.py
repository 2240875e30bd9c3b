"""Run under torchrun on N GPUs: the sequence-parallel pipeline must reproduce the single-GPU pipeline.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/sp_check.py
Rank 0 prints one JSON line {"world":N,"rel_l2":...,"index_trace_equal":true}.
"""
import json
import os
import sys
import types
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from inferix_b200.kvcache_manager import KVCacheManager, KVCacheRequest  # noqa: E402
from inferix_b200.parallel import ParallelConfig  # noqa: E402
from inferix_b200.pipeline import CausalInferencePipeline, DecodeMode  # noqa: E402
from inferix_b200.synthetic import TINY, synth_state_dict  # noqa: E402
from inferix_b200.wan_model import CausalWanModel  # noqa: E402
from inferix_b200.wrapper import WanDiffusionWrapper  # noqa: E402


def run(pc, dev, noise, context, seed):
    model = CausalWanModel(**TINY, local_attn_size=6, sink_size=1, parallel_config=pc)
    model.load_state_dict(synth_state_dict(TINY, seed=0))
    model = model.to(torch.bfloat16).to(dev)
    args = types.SimpleNamespace(denoising_step_list=[1000, 500], warp_denoising_step=True, num_frame_per_block=3,
                                 context_noise=0)
    pipe = CausalInferencePipeline(args, dev, generator=WanDiffusionWrapper(model=model, timestep_shift=5.0,
                                                                          parallel_config=pc), parallel_config=pc)
    gen = torch.Generator().manual_seed(seed)
    pipe.renoise_fn = lambda x: torch.randn(x.shape, generator=gen, dtype=torch.float32).to(x.dtype).to(x.device)
    trace = []
    hook = model.blocks[0].register_forward_hook(lambda m, i, o: trace.append(pipe.kv_cache_meta[0]["_ifx_plan"]))
    out = pipe.inference(noise=noise.to(dev), text_prompts=context.to(dev), kv_cache_manager=KVCacheManager(dev),
                         kv_cache_requests=[KVCacheRequest("r")], decode_mode=DecodeMode.NO_DECODE)
    hook.remove()
    return out, trace


def main():
    world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    torch.set_grad_enabled(False)
    g = torch.Generator().manual_seed(1)
    noise = torch.randn(1, 12, 16, 16, 16, generator=g).bfloat16()
    context = torch.randn(1, 20, TINY["text_dim"], generator=g).bfloat16()
    out_sp, trace_sp = run(ParallelConfig(ring_size=world, world_size=world, rank=rank, local_rank=local), dev, noise,
                           context, 99)
    torch.cuda.synchronize()
    dist.barrier()
    if rank == 0:
        out_1, trace_1 = run(ParallelConfig(), dev, noise, context, 99)
        rel = ((out_sp.float() - out_1.float()).norm() / out_1.float().norm()).item()
        print(json.dumps({"world": world, "rel_l2": rel, "bit_equal": bool(torch.equal(out_sp, out_1)),
                          "index_trace_equal": trace_sp == trace_1, "blocks": len(trace_1)}), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
