"""SpectraNet block tail at stage-1 size: unfused (layernorm_stream + pooled 1x1 GEMM) vs fused (acb_gemm_ln_bf16)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from applecider_b200 import ops  # noqa: E402

stage = int(sys.argv[1]) if len(sys.argv) > 1 else 1
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
L, K, N = {1: (1024, 384, 128), 2: (256, 768, 256), 3: (64, 1536, 512)}[stage]
M = B * L
bn = 128 if N == 128 else 256
parts = (K + bn - 1) // bn
torch.manual_seed(0)
y = torch.randn(M, K, device="cuda").to(torch.bfloat16)
stats = torch.empty(M, parts, 2, device="cuda")
yf = y.float()
for j in range(parts):
    stats[:, j, 0] = yf[:, j * bn:(j + 1) * bn].sum(1)
    stats[:, j, 1] = (yf[:, j * bn:(j + 1) * bn] ** 2).sum(1)
del yf
g, be = torch.ones(K, device="cuda"), torch.zeros(K, device="cuda")
wd = (torch.randn(N, K, device="cuda") * K ** -0.5).to(torch.bfloat16)
bd = torch.zeros(N, device="cuda")


def timeit(f, n=10):
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        o = f()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, o


t_ln, yn = timeit(lambda: ops.layernorm(y, g, be, 1e-5, post_act=ops.ACT_GELU))
t_g, ref = timeit(lambda: ops.gemm(yn, wd, bd, pool4=True))
t_f, got = timeit(lambda: ops.gemm_ln(y, wd, bd, stats, parts, g, be, 1e-5, pool4=True))
print(f"stage {stage}: M={M} K={K} N={N}: layernorm {t_ln:.3f} ms + pooled GEMM {t_g:.3f} ms = {t_ln + t_g:.3f} ms; fused {t_f:.3f} ms "
      f"({M * K * 2 / t_f / 1e6:.0f} GB/s of the activation read); max|diff| {(got.float() - ref.float()).abs().max().item():.4f}")
