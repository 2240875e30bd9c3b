"""Mirror of inferix.kvcache_manager (reference kvcache_manager/__init__.py) on the native paged cache."""
from .kvcache_manager import (KVCacheManager, KVCacheRequest, KVCacheRequestSpec, KVCaches, KVCacheSpec,
                              KVCacheTensorSpec, align, cdiv, get_dtype_size)

__all__ = ["KVCacheManager", "KVCacheRequest", "KVCacheRequestSpec", "KVCaches", "KVCacheSpec", "KVCacheTensorSpec",
           "align", "cdiv", "get_dtype_size"]
