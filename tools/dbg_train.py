import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import torch
torch.backends.cudnn.allow_tf32=False; torch.backends.cuda.matmul.allow_tf32=False
from test_gpu_train import _pair
from applecider_b200 import synth, fn, ops
from applecider_b200.train import SpectraConvs
DEV='cuda'
prod, oracle = _pair("SpectraNet")
L=1000; B=2
s = synth.spectra(B, seed=93, L=L)
# ---- mine: stage 0 and 1 manually with retained grads
sig = s.to(DEV).view(B, L); h = sig.view(B, L, 1)
mine = {}
Lc = L
for si in range(2):
    blk = prod.all_stages[si][0]
    params = [c.weight for c in blk.convs] + [c.bias for c in blk.convs]
    y = SpectraConvs.apply(h, None, blk, B, Lc, torch.float32, *params); y.retain_grad(); mine[f'conv{si}']=y
    ln = fn.layernorm(y, blk.norm.weight, blk.norm.bias, blk.norm.eps); ln.retain_grad(); mine[f'ln{si}']=ln
    a = fn.act(ln, ops.ACT_GELU); a.retain_grad(); mine[f'act{si}']=a
    z = fn.linear(a, blk.downsample.weight.view(blk.out_channels,-1), blk.downsample.bias); z.retain_grad(); mine[f'down{si}']=z
    p = fn.MaxPool.apply(z.view(B,Lc,blk.out_channels), B, Lc, blk.out_channels, 4); p.retain_grad(); mine[f'pool{si}']=p
    Lc//=4; h = p
g = torch.randn(h.shape, generator=torch.Generator().manual_seed(3))
h.backward(g.to(DEV))
# ---- oracle
x = s.clone(); ref={}
for si in range(2):
    blk = oracle.all_stages[si][0]
    y = torch.cat([c(x) for c in blk.convs],1); y.retain_grad(); ref[f'conv{si}']=y
    ln = blk.norm(y.transpose(1,2)).transpose(1,2); ln.retain_grad(); ref[f'ln{si}']=ln
    a = torch.nn.functional.gelu(ln); a.retain_grad(); ref[f'act{si}']=a
    z = blk.downsample(a); z.retain_grad(); ref[f'down{si}']=z
    p = torch.nn.functional.max_pool1d(z,4); p.retain_grad(); ref[f'pool{si}']=p
    x = p
x.backward(g.transpose(1,2))
for k in mine:
    r = ref[k].grad.transpose(1,2).reshape(mine[k].grad.shape) if ref[k].grad.dim()==3 else ref[k].grad
    m = mine[k].grad.cpu()
    fr = ref[k].detach().transpose(1,2).reshape(mine[k].shape)
    print(f"{k:8s} fwd rel {((mine[k].detach().cpu()-fr).abs().max()/fr.abs().max()).item():.2e}  grad rel {((m-r).abs().max()/r.abs().max()).item():.2e}")
print("---- direct check of stage-0 maxpool backward")
z0 = mine['down0'].detach().view(B, L, 64)
dy = mine['pool0'].grad
zz = z0.clone().requires_grad_(True)
torch.nn.functional.max_pool1d(zz.transpose(1,2), 4).transpose(1,2).backward(dy)
mg = mine['down0'].grad.view(B, L, 64)
print("mine vs torch-gpu on same inputs:", (mg - zz.grad).abs().max().item(), zz.grad.abs().max().item())
rg = ref['down0'].grad.transpose(1,2)
print("torch-gpu vs oracle-cpu:", (zz.grad.cpu() - rg).abs().max().item())
w = z0.view(B, 250, 4, 64)
mx = w.amax(2, keepdim=True)
print("windows with ties:", ((w == mx).sum(2) > 1).sum().item(), "of", B*250*64)
zr = ref['down0'].detach().transpose(1,2)
am_m = w.argmax(2).cpu(); am_r = zr.reshape(B,250,4,64).argmax(2)
print("argmax disagreements mine-vs-oracle forward values:", (am_m != am_r).sum().item())
