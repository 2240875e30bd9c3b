from .causvid_kv_cache_manager import CausVidKVCacheManager, KVCacheManagerFactory
from .self_forcing_kv_cache_manager import SelfForcingKVCacheManager, SelfForcingKVCacheManagerFactory

__all__ = ["CausVidKVCacheManager", "KVCacheManagerFactory", "SelfForcingKVCacheManager",
           "SelfForcingKVCacheManagerFactory"]
