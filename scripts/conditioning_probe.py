import sys, torch, time
sys.path.insert(0, '/root/repo')
from oracle import fcd_oracle as O
from tests._util import load_golden, rel_l2, rel_err
torch.set_num_threads(8)
# perception conditioning
f = load_golden("perception.pt")
net = O.vgg16_features(1234)
sd = {k: v.double() for k, v in net.state_dict().items()}
for key, layers, pb in (("perband", 1, True), ("rgb5", 5, False)):
    d = f[key]
    def run(eps):
        g = d["g"].double().clone()
        if eps: g = g * (1 + eps * torch.randn(g.shape, generator=torch.Generator().manual_seed(0)).double())
        g.requires_grad_(True)
        cm = d["cmap"].double().clone().requires_grad_(True)
        v = O.perception_loss(sd, d["t"].double(), g, cm, layers, pb)
        v.backward()
        return v.item(), g.grad, cm.grad
    v0, g0, c0 = run(0.0)
    for eps in (1e-6, 1e-5, 3e-5):
        v1, g1, c1 = run(eps)
        print(key, eps, "value rel", abs(v1 - v0) / v0, "dg rel-L2", rel_l2(g1, g0), "max", rel_err(g1, g0), "dcmap rel-L2", rel_l2(c1, c0))
# segmentor at 13x256x256 B=2 conditioning (fp64)
def pair(B, C, H, W, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, C, H, W, generator=g)
    y = x + 0.3 * torch.randn(B, C, H, W, generator=g)
    y[:, :, H // 4:H // 2, W // 3:2 * W // 3] = torch.randn(B, C, H // 2 - H // 4, 2 * W // 3 - W // 3, generator=g)
    return x, y
C, H, W, B = 13, 256, 256, 2
x, y = pair(B, C, H, W, 521 + H)
r = torch.randn(B, 1, H, W, generator=torch.Generator().manual_seed(522)).double()
def runS(eps, dtype=torch.float64):
    sd = O.clone_sd(O.make_state_dict(O.segmentor_spec(C, 1, True), 12), dtype=dtype, requires_grad=True)
    xx = x.to(dtype).clone()
    if eps: xx = xx * (1 + eps * torch.randn(xx.shape, generator=torch.Generator().manual_seed(0)).to(dtype))
    c = O.segmentor(sd, xx, y.to(dtype), bilinear=True, train=True)
    (c * r.to(dtype)).sum().backward()
    return c.detach(), {k: v.grad for k, v in sd.items() if v.requires_grad}
t = time.time()
c0, g0 = runS(0.0)
print("S fp64 run", time.time() - t)
for eps in (1e-5, 6e-5):
    c1, g1 = runS(eps)
    worst = max((rel_l2(g1[k], g0[k]), k) for k in g0 if g0[k].numel() >= 64 and g0[k].abs().max() > 1e-4)
    print("S eps", eps, "cmap max rel", rel_err(c1, c0), "worst grad rel-L2", worst, "inc.double_conv.1.bias", rel_l2(g1["inc.double_conv.1.bias"], g0["inc.double_conv.1.bias"]))
c32, g32 = runS(0.0, torch.float32)
worst = max((rel_l2(g32[k], g0[k]), k) for k in g0 if g0[k].numel() >= 64 and g0[k].abs().max() > 1e-4)
print("S fp32 oracle vs fp64 oracle: cmap", rel_err(c32, c0), "worst grad", worst, "inc.double_conv.1.bias", rel_l2(g32["inc.double_conv.1.bias"], g0["inc.double_conv.1.bias"]))
