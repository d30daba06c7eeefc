"""Time individual tcgen05 GEMM shapes (CUDA events, L2-cold by rotating buffers)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from applecider_b200 import ops

def run(M, N, K, act=0, bn=None, reps=5, res=False, out_dtype=torch.bfloat16):
    a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    w = (torch.randn(N, K, device="cuda") * K**-0.5).to(torch.bfloat16)
    b = torch.randn(N, device="cuda")
    r = torch.randn(M, N, device="cuda").to(torch.bfloat16) if res else None
    out = torch.empty(M, N, device="cuda", dtype=out_dtype)
    for _ in range(2):
        ops.gemm(a, w, b, act=act, out=out, bn=bn, res=r, res_mode=(ops.RES_ADD if res else ops.RES_NONE))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        ops.gemm(a, w, b, act=act, out=out, bn=bn, res=r, res_mode=(ops.RES_ADD if res else ops.RES_NONE))
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    byts = (M*K + N*K + M*N*(2 if res else 1)) * 2
    print(f"M={M} N={N} K={K} act={act} bn={bn} res={res}: {ms:.3f} ms  {2*M*N*K/ms/1e9:.1f} TFLOP/s  {byts/ms/1e6:.0f} GB/s", flush=True)

if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("all", "ffn1"):
        run(237482, 512, 128, act=1)
    if which == "all":
        run(237482, 512, 128, act=1, bn=128)
        run(237482, 384, 128)
        run(237482, 128, 512, res=True)
        run(921600, 384, 96, act=2)
        run(921600, 96, 384, res=True)
        run(8192, 8192, 8192, bn=256)
        run(8192, 8192, 8192, bn=128)
    if which == "cnx":  # ConvNeXt stage-0 MLP pair (for ncu: first launch after warm-up)
        run(921600, 384, 96, act=2)
        run(921600, 96, 384, res=True)
    if which == "down":  # SpectraNet 1x1 downsample GEMMs (stages 1-3), tile-width comparison
        for bn in (None, 128, 64):
            run(4194304, 128, 384, bn=bn)
        for bn in (None, 128):
            run(1048576, 256, 768, bn=bn)
            run(262144, 512, 1536, bn=bn)
