"""Times the HBM/L2-bound row kernels at the 720p block shape (developer probe)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from inferix_b200 import ops  # noqa: E402
from inferix_b200._lib import RopeGrid  # noqa: E402
from oracle import wan_oracle as wo  # noqa: E402
from tools.gpu_probe import timeit  # noqa: E402

S, C = 10800, 1536
qkv = torch.randn(S, 3 * C, device="cuda").bfloat16()
w = torch.ones(C, device="cuda").bfloat16()
tab = ops.rope_table(wo.rope_freqs(128), "cuda")
g = RopeGrid(3, 45, 80, 21, 0, 3600)
q = torch.empty(S, C, device="cuda", dtype=torch.bfloat16)
k, v = torch.empty_like(q), torch.empty_like(q)
ms = timeit(lambda: ops.qk_norm_rope_append(qkv, w, w, tab, g, 12, 128, q_out=q, k_out=k, v_out=v), iters=20)
print("qk_norm_rope_append 720p: %.1f us  %.0f GB/s" % (ms * 1e3, 6 * S * C * 2 / ms / 1e6))
x = torch.randn(S, C, device="cuda").bfloat16()
m = torch.randn(3, 6, C, device="cuda").bfloat16()
o = torch.empty_like(x)
ms = timeit(lambda: ops.ln_modulate(x, o, shift=m[:, 0], scale=m[:, 1], tokens_per_frame=3600), iters=20)
print("ln_modulate 720p: %.1f us  %.0f GB/s" % (ms * 1e3, 2 * S * C * 2 / ms / 1e6))
ms = timeit(lambda: ops.rmsnorm(x, w, o), iters=20)
print("rmsnorm 720p: %.1f us  %.0f GB/s" % (ms * 1e3, 2 * S * C * 2 / ms / 1e6))
