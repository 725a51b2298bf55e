import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fcdgan_b200 as fb
from oracle import fcd_oracle as O
from tests._util import load_golden, rel_err
DEV = "cuda:0"
f = load_golden("g13_train.pt")
net = fb.Generator(13); net.load_state_dict(O.make_state_dict(O.generator_spec(13), f["seed"])); net.to(DEV).train()
x = f["x"].to(DEV).requires_grad_(True)
y = net(x)
print("y err", rel_err(y, f["y"]))
(y * f["r"].to(DEV)).sum().backward()
e = (x.grad.cpu() - f["dx"]).abs()
print("dx err", rel_err(x.grad, f["dx"]), "max at", (e == e.max()).nonzero()[0].tolist(), "shape", list(e.shape))
print("err by row", e.amax(dim=(0, 1, 3))[:30])
print("err by col", e.amax(dim=(0, 1, 2))[:30])
for k, p in net.named_parameters():
    ref = f["grads"][k]
    if "full" in ref:
        print(k, rel_err(p.grad, ref["full"]))
    else:
        print(k, "norm", p.grad.norm().item(), ref["norm"], "head err", (p.grad.flatten()[:16].cpu() - ref["head"]).abs().max().item(), ref["head"].abs().max().item())
