from .causvid_kv_cache_manager import CausVidKVCacheManager, KVCacheManagerFactory
from .magi_kv_cache_manager import InferenceParams, KVMetaArgs, MagiKVCacheManager
from .self_forcing_kv_cache_manager import SelfForcingKVCacheManager, SelfForcingKVCacheManagerFactory

__all__ = ["CausVidKVCacheManager", "KVCacheManagerFactory", "MagiKVCacheManager", "InferenceParams", "KVMetaArgs", "SelfForcingKVCacheManager",
           "SelfForcingKVCacheManagerFactory"]
