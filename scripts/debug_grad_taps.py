"""GPU debugging aid: bisect the segmenter backward pass by comparing d(loss)/d(tap activations) between the
CUDA product and the oracle, plus BN backward at model scale."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import pnp_b200
from pnp_b200 import functional as F, runtime as rt, layers as L, source_segmenter as seg
from oracle.pnp_graphs import OracleSegmenter, init_numpy_params, synthetic_images, synthetic_labels
from oracle.tf14_numpy import label_decomp
from oracle import tf14_torch as T

B = 2
dev = "cuda"


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


rt.set_conv_backend(sys.argv[1] if len(sys.argv) > 1 else "simt")
ws, bns = OracleSegmenter.layout()
P = init_numpy_params(ws, bns, 0, 0.05)
ck = {"cross_flag": True, "miu_cross": 1.0, "dice_flag": True, "miu_dice": 1.0, "regularizer": 1e-4}
net = seg.Full_DRN(channels=3, n_class=5, batch_size=B, cost_kwargs=dict(ck))
rt.load_state_dict(P)
oracle = OracleSegmenter(P, B)
x = synthetic_images(B, 1234)
lab = synthetic_labels(B, 99)
y = torch.from_numpy(label_decomp(5, lab))

out = oracle.forward(x, 1.0, True)
cost, reg, wce, dice = oracle.losses(out["logits"], y)
keys = ["logits", "c9_2", "b8", "b7", "c6_2", "c4_2"]
go = torch.autograd.grad(cost, [out[k] for k in keys], retain_graph=True)

for v in rt.global_variables():
    if v.pnp_trainable:
        v.requires_grad_(True)
logits, taps = net.forward(x.to(dev), keep_prob=1.0, main_bn=True, adapt_bn=True, return_taps=True)
taps["logits"] = logits
for k in keys:
    taps[k].retain_grad()
wce_g, dice_g = net.losses(logits, y.to(dev))
one = torch.tensor(1.0, device=dev)
torch.autograd.backward([wce_g, dice_g], [one, one])
print("loss  wce %.3e dice %.3e" % (rel(wce_g.reshape(1), wce.reshape(1)), rel(dice_g.reshape(1), dice.reshape(1))))
for k, g in zip(keys, go):
    print("fwd %-8s %.3e   d/d%-8s %.3e   max|g| %.3e" % (k, rel(taps[k], out[k]), k, rel(taps[k].grad, g), float(g.abs().max())))

print("== conv_bn_relu backward at model scale (B=2, 32x32, 512->512, train-mode BN), simt")
rt.reset_default_graph()
rt.set_conv_backend("simt")
g = torch.Generator().manual_seed(5)
xx = torch.randn(2, 32, 32, 512, generator=g)
ww = torch.randn(3, 3, 512, 512, generator=g) * 0.05
bn = T.BNState(512, torch.float64)
xo, wo = xx.double().requires_grad_(True), ww.double().requires_grad_(True)
yo = T.conv_bn_relu2d(xo, wo, 1.0, bn, is_train=True, leak=True)
r = torch.randn(yo.shape, generator=g)
(yo * r.double()).sum().backward()
xg, wg = xx.to(dev).requires_grad_(True), ww.to(dev).requires_grad_(True)
yy = L.conv_bn_relu2d(xg, wg, 1.0, is_train=True, scope="t", leak=True)
yy.backward(r.to(dev))
v = rt.graph.vars
print("y %.2e dx %.2e dw %.2e dgamma %.2e dbeta %.2e" % (rel(yy, yo), rel(xg.grad, xo.grad), rel(wg.grad, wo.grad),
                                                          rel(v["t/gamma"].grad, bn.gamma.grad), rel(v["t/beta"].grad, bn.beta.grad)))

print("== per-layer backward check inside the model (BN backward + dgrad recomputed in torch fp64 from the saved tensors)")
import torch.nn.functional as TF
rt.reset_default_graph()
rt.set_conv_backend("simt")
net = seg.Full_DRN(channels=3, n_class=5, batch_size=B, cost_kwargs=dict(ck))
rt.load_state_dict(P)
for v in rt.global_variables():
    if v.pnp_trainable:
        v.requires_grad_(True)
names = {id(v): n for n, v in rt.graph.vars.items()}


def hook(sv, dy, g, dz, dx):
    cfg, geom, W = sv["cfg"], sv["geom"], sv["W"]
    name = names.get(id(W), "?")
    msg = "%-22s" % name
    if cfg.bn is not None and cfg.bn_training:
        y, z = sv["y"].double(), sv["z"].double()
        mean, invstd, gamma = sv["mean"].double(), sv["invstd"].double(), cfg.bn.gamma.double()
        gr = dy.double() * torch.where(y > 0, 1.0, 0.2)
        xh = (z - mean) * invstd
        red = (0, 1, 2)
        dzr = gamma * invstd * (gr - gr.mean(red) - xh * (gr * xh).mean(red))
        msg += " g %.2e dz %.2e" % (rel(g, gr), rel(dz, dzr))
    if dx is not None and sv["p"] == 0 and sv["skip_c"] == 0:
        # dgrad reference from OUR dz through torch's conv (fp64)
      with torch.enable_grad():
        xin = torch.zeros(sv["x_shape"], dtype=torch.float64, device=dz.device, requires_grad=True)
        pt, pb = F.same_pad(geom.H, geom.kh, geom.stride, geom.dil)
        pl, pr = F.same_pad(geom.W, geom.kw, geom.stride, geom.dil)
        yy = TF.conv2d(TF.pad(xin.permute(0, 3, 1, 2), (pl, pr, pt, pb)), W.detach().double().permute(3, 2, 0, 1), stride=geom.stride,
                       dilation=geom.dil).permute(0, 2, 3, 1)
        (gx,) = torch.autograd.grad(yy, xin, dz.double())
        msg += " dx|dz %.2e" % rel(dx, gx)
    print(msg)


F.DEBUG_HOOK = hook
logits = net.forward(x.to(dev), keep_prob=1.0, main_bn=True, adapt_bn=True)
wce_g, dice_g = net.losses(logits, y.to(dev))
torch.autograd.backward([wce_g, dice_g], [one, one])
F.DEBUG_HOOK = None
