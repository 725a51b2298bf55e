"""Compare every intermediate gradient of the CUDA Generator backward with torch autograd on the oracle."""
import sys, os, torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fcdgan_b200 as fb
from fcdgan_b200 import engine as E
from oracle import fcd_oracle as O
from tests._util import load_golden, rel_err
DEV = "cuda:0"
f = load_golden("g13_train.pt")
sd0 = O.make_state_dict(O.generator_spec(13), f["seed"])
net = fb.Generator(13); net.load_state_dict(sd0); net.to(DEV).train()
x = f["x"].to(DEV).requires_grad_(True)
E.DEBUG_CAPTURE = []
y = net(x)
(y * f["r"].to(DEV)).sum().backward()
cap = E.DEBUG_CAPTURE
# oracle with retained intermediate grads (fp64 on CPU)
sd = O.clone_sd(sd0, dtype=torch.float64, requires_grad=True)
xx = f["x"].double().clone().requires_grad_(True)
keep = []
def K(t):
    t.retain_grad(); keep.append(t); return t
def bn(p, t): return O._bn(sd, p, t, True)
b1z = K(O._conv(sd, "block1.0", xx, padding=4))
b1 = K(F.prelu(b1z, sd["block1.1.weight"]))
h = b1
for i in range(2, 7):
    p = f"block{i}"
    z1 = K(O._conv(sd, f"{p}.conv1", h, padding=1))
    a = K(F.prelu(bn(f"{p}.bn1", z1), sd[f"{p}.prelu.weight"]))
    z2 = K(O._conv(sd, f"{p}.conv2", a, padding=1))
    o = K(h + bn(f"{p}.bn2", z2))
    h = o
z7 = K(O._conv(sd, "block7.0", h, padding=1))
s7 = K(b1 + bn("block7.1", z7))
yy = O._conv(sd, "block8", s7, padding=4)
(yy * f["r"].double()).sum().backward()
# oracle order of (da, dz) pairs in reverse: s7/z7, then per block 6..2: (o, z2), (a, z1); then (b1, b1z)
ref = [(s7.grad, z7.grad)]
# keep = [b1z, b1, (z1,a,z2,o)*5, z7, s7]
for i in range(5, -1 + 1, -1):
    pass
blocks = [keep[2 + 4 * j: 6 + 4 * j] for j in range(5)]
for z1, a, z2, o in reversed(blocks):
    ref.append((o.grad, z2.grad)); ref.append((a.grad, z1.grad))
ref.append((b1.grad, b1z.grad))
names = ["s7"] + sum([[f"blk{i}.out", f"blk{i}.a"] for i in (6, 5, 4, 3, 2)], []) + ["b1"]
assert len(cap) == 2 * len(ref), (len(cap), len(ref))
for k, (rda, rdz) in enumerate(ref):
    da, dz = cap[2 * k][1].cpu().double(), cap[2 * k + 1][1].cpu().double()
    e = (dz - rdz).abs()
    loc = (e == e.max()).nonzero()[0].tolist()
    print(f"{names[k]:10s} da err {rel_err(da, rda):.2e}  dz err {rel_err(dz, rdz):.2e}  dz max at {loc}")
