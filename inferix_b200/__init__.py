"""inferix_b200 — B200-native (sm_100a) block-diffusion denoising hot path behind Inferix's Python surface.

Import map for a user of the reference:
    inferix.kvcache_manager.*                         -> inferix_b200.kvcache_manager.*
    inferix.models.attention.{attention,flash_attention} -> inferix_b200.attention.*
    inferix.models.self_forcing.causal_model.*        -> inferix_b200.wan_model.*
    inferix.models.self_forcing.wrapper.WanDiffusionWrapper -> inferix_b200.wrapper.WanDiffusionWrapper
    inferix.models.schedulers.flow_match.FlowMatchScheduler -> inferix_b200.scheduler.FlowMatchScheduler
    inferix.pipeline.self_forcing.CausalInferencePipeline -> inferix_b200.pipeline.CausalInferencePipeline
    inferix.models.wan_base.ParallelConfig            -> inferix_b200.parallel.ParallelConfig
Everything computes through libinferix_b200.so (include/inferix_b200.h); there is no CPU fallback.
"""
__version__ = "0.1.0"
