import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import applecider_b200 as ab
from applecider_b200 import synth, ops
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
from applecider_b200 import spectra as _sp
_sp.FUSE_STAGE1 = len(sys.argv) > 2 and sys.argv[2] == 'fuse1'
cfg = ab.default_config(); cfg["model"]["SpectraNet"]["compute_dtype"] = "bf16"
m = ab.SpectraNet(cfg); m.load_state_dict(synth.det_state_dict(m, 0)); m = m.cuda().eval()
x = synth.spectra(B, seed=1).cuda()
with torch.no_grad():
    for _ in range(2): m((x, None, None))
    torch.cuda.synchronize()
    ops.profile_start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): m((x, None, None))
    e1.record(); torch.cuda.synchronize()
    reg = ops.profile_stop()
print("B", B, "ms/fwd", e0.elapsed_time(e1) / 3, {k: sum(v) / len(v) for k, v in reg.items()})
import ctypes
from applecider_b200 import _lib
l = _lib.lib()
buf = (ctypes.c_ulonglong * 4)()
l.acb_debug_timing(1, buf)
with torch.no_grad():
    m((x, None, None))
l.acb_debug_timing(0, buf)
n = max(1, buf[3])
print("conv_ln phases (cycles/CTA): prologue %.0f  mainloop %.0f  epilogue %.0f  ctas %d" % (buf[0] / n, buf[1] / n, buf[2] / n, buf[3]))
