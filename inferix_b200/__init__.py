"""inferix_b200 — B200-native (sm_100a) block-diffusion denoising hot path behind Inferix's Python surface.

Import map for a user of the reference:
    inferix.kvcache_manager.*                         -> inferix_b200.kvcache_manager.*
    inferix.models.attention.{attention,flash_attention} -> inferix_b200.attention.*
    inferix.models.self_forcing.causal_model.*        -> inferix_b200.wan_model.*
    inferix.models.self_forcing.wrapper.WanDiffusionWrapper -> inferix_b200.wrapper.WanDiffusionWrapper
    inferix.models.schedulers.flow_match.FlowMatchScheduler -> inferix_b200.scheduler.FlowMatchScheduler
    inferix.pipeline.self_forcing.CausalInferencePipeline -> inferix_b200.pipeline.CausalInferencePipeline
    inferix.models.wan_base.ParallelConfig            -> inferix_b200.parallel.ParallelConfig
    inferix.pipeline.causvid.CausalInferencePipeline / models.causvid.* -> inferix_b200.causvid.*
    inferix.models.magi.dit.dit_module.{TransformerLayer,TransformerBlock,...} -> inferix_b200.magi_layer.*
    inferix.models.magi.dit.dit_model.VideoDiTModel   -> inferix_b200.magi_model.VideoDiTModel
    inferix.pipeline.magi.video_generate.SampleTransport (+ index helpers) -> inferix_b200.magi_pipeline / magi_schedule
    inferix.distributed.parallelism.context_parallel (Ulysses) -> inferix_b200.magi_cp, inferix_b200.ulysses_scheduler
    inferix.distributed.parallel_state / dist_utils   -> inferix_b200.parallel_state / dist_utils
    inferix.models.wan_base.utils.fm_solvers_unipc.FlowUniPCMultistepScheduler -> inferix_b200.unipc
    inferix.pipeline.self_forcing.CausalDiffusionInferencePipeline -> inferix_b200.diffusion_pipeline
    inferix.models.attention.{distributed.CoreAttention,backends.collect_supported_attn} -> inferix_b200.attention
Everything computes through libinferix_b200.so (include/inferix_b200.h); there is no CPU fallback.
"""
__version__ = "0.1.0"
