"""GPU debugging aid (not part of the product or the test-suite): localises discrepancies by comparing our
kernels with torch's own CUDA convolution ops at full-size shapes, and SIMT vs tcgen05 sub-steps."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as TF
import pnp_b200
from pnp_b200 import functional as F, runtime as rt, layers as L
from pnp_b200._C import ConvGeom

dev = "cuda"
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def ref_conv(x, w, s, d, pt, pl, pb, pr):
    xn = TF.pad(x.permute(0, 3, 1, 2).double(), (pl, pr, pt, pb))
    return TF.conv2d(xn, w.permute(3, 2, 0, 1).double(), stride=s, dilation=d).permute(0, 2, 3, 1)


def case(B, H, W, Cin, Cout, k, s, d, tag):
    g = torch.Generator(device="cpu").manual_seed(1)
    x = torch.randn(B, H, W, Cin, generator=g).to(dev)
    w = (torch.randn(k, k, Cin, Cout, generator=g) * 0.1).to(dev)
    pt, pb = F.same_pad(H, k, s, d)
    pl, pr = F.same_pad(W, k, s, d)
    xr = x.double().requires_grad_(True)
    wr = w.double().requires_grad_(True)
    yr = ref_conv(xr, wr, s, d, pt, pl, pb, pr)
    r = torch.randn(yr.shape, generator=g).to(dev)
    (yr * r.double()).sum().backward()
    for backend in ("simt", "tc3"):
        rt.set_conv_backend(backend)
        xg = x.clone().requires_grad_(True)
        wg = w.clone().requires_grad_(True)
        y = L.conv2d(xg, wg, 1.0, strides=[1, s, s, 1]) if d == 1 else L.dilate_conv2d(xg, wg, 1.0, rate=d)
        y.backward(r)
        print("%-28s %-5s y %.2e dx %.2e dw %.2e" % (tag, backend, rel(y, yr), rel(xg.grad, xr.grad), rel(wg.grad, wr.grad)))


print("== full-size conv flavours vs torch fp64")
case(2, 256, 256, 3, 16, 3, 1, 1, "conv1_1")
case(2, 256, 256, 16, 16, 3, 1, 1, "g1.res")
case(2, 128, 128, 16, 32, 3, 1, 1, "g2.res.a")
case(2, 128, 128, 32, 32, 3, 1, 1, "g2.res.b")
case(2, 64, 64, 32, 64, 3, 1, 1, "g3.res.a")
case(2, 64, 64, 64, 64, 3, 1, 1, "g3.res.b")
case(2, 32, 32, 64, 128, 3, 1, 1, "g4.res.a")
case(2, 32, 32, 256, 512, 3, 1, 1, "g7.res.a")
case(2, 32, 32, 512, 512, 3, 1, 2, "g8.dr")
case(2, 256, 256, 32, 64, 3, 1, 1, "cls_1.a")
case(2, 256, 256, 64, 64, 3, 2, 1, "cls_1_3")
case(2, 128, 128, 128, 128, 5, 2, 1, "cls_2_3")
case(2, 16, 16, 512, 512, 5, 4, 1, "cls_5_3")
case(2, 256, 256, 5, 16, 3, 2, 1, "mask_cls_1")

print("== SYMMETRIC convs (g10, output) vs oracle-on-GPU (manual mirror pad + torch fp64)")
from oracle import tf14_torch as T
for (B, H, Cin, Cout, k, tag) in ((2, 32, 512, 2560, 3, "g10"), (2, 256, 40, 5, 5, "output")):
    g = torch.Generator(device="cpu").manual_seed(2)
    x = torch.randn(B, H, H, Cin, generator=g).to(dev)
    w = (torch.randn(k, k, Cin, Cout, generator=g) * 0.05).to(dev)
    xr, wr = x.double().requires_grad_(True), w.double().requires_grad_(True)
    yr = T.conv2d_raw(xr, wr, padding="SYMMETRIC")
    r = torch.randn(yr.shape, generator=g).to(dev)
    (yr * r.double()).sum().backward()
    for backend in ("simt", "tc3"):
        rt.set_conv_backend(backend)
        xg, wg = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
        y = L.conv2d(xg, wg, 1.0, padding="SYMMETRIC")
        y.backward(r)
        print("%-28s %-5s y %.2e dx %.2e dw %.2e" % (tag, backend, rel(y, yr), rel(xg.grad, xr.grad), rel(wg.grad, wr.grad)))

print("== tcgen05 dgrad with accumulate: determinism + value")
rt.set_conv_backend("tc3")
g = torch.Generator(device="cpu").manual_seed(3)
for (B, H, Cin, Cout) in ((2, 16, 64, 128), (2, 16, 64, 64), (8, 32, 256, 512)):
    dz = torch.randn(B, H, H, Cout, generator=g).to(dev)
    w = (torch.randn(3, 3, Cin, Cout, generator=g) * 0.1).to(dev)
    base = torch.randn(B, H, H, Cin, generator=g).to(dev)
    geom = ConvGeom(B, H, H, Cin, H, H, Cout, 3, 3, 1, 1, 1, 1)
    rt.set_conv_backend("simt")
    ref = F.conv_dgrad_raw(dz, w, geom, into=base.clone())
    ref0 = F.conv_dgrad_raw(dz, w, geom)
    rt.set_conv_backend("tc3")
    outs = [F.conv_dgrad_raw(dz, w, geom, into=base.clone()) for _ in range(4)]
    o0 = F.conv_dgrad_raw(dz, w, geom)
    torch.cuda.synchronize()
    print("B%d H%d %d<-%d  acc: %s  no-acc %.2e  identical runs: %s" % (
        B, H, Cin, Cout, ["%.2e" % rel(o, ref) for o in outs], rel(o0, ref0), all(torch.equal(outs[0], o) for o in outs[1:])))

print("== elementwise backward pieces vs torch autograd (maxpool, PS, mirror pad, seg loss) at full size")
x = torch.randn(2, 256, 256, 16, device=dev)
xr = x.double().requires_grad_(True)
yr = TF.max_pool2d(xr.permute(0, 3, 1, 2), 2).permute(0, 2, 3, 1)
r = torch.randn_like(yr)
(yr * r).sum().backward()
xg = x.clone().requires_grad_(True)
y = L.max_pool2d(xg, 2)
y.backward(r.float())
print("maxpool y %.2e dx %.2e" % (rel(y, yr), rel(xg.grad, xr.grad)))
x = torch.randn(2, 32, 32, 2560, device=dev)
xr = x.double().requires_grad_(True)
yr = T.PS(xr, 8, 40, 2)
r = torch.randn_like(yr)
(yr * r).sum().backward()
xg = x.clone().requires_grad_(True)
from pnp_b200 import ops
y = ops.PS(xg, 8, 40, 2)
y.backward(r.float())
print("PS y %.2e dx %.2e" % (rel(y, yr), rel(xg.grad, xr.grad)))
print("done")
