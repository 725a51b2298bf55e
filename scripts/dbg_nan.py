"""Poison every torch.empty buffer with NaN and report kernels that leave their outputs partially unwritten."""
import sys, os, traceback, weakref, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fcdgan_b200 as fb
from fcdgan_b200 import engine as E
from oracle import fcd_oracle as O
from tests._util import load_golden, rel_err
DEV = "cuda:0"
_orig_empty, _orig_empty_like = torch.empty, torch.empty_like
live = []
def _site():
    st = traceback.extract_stack()[:-2]
    return " <- ".join(f"{os.path.basename(f.filename)}:{f.lineno}" for f in st[-4:])
def empty(*a, **k):
    t = _orig_empty(*a, **k)
    if t.is_cuda and t.is_floating_point():
        t.fill_(float("nan"))
        live.append((weakref.ref(t), _site()))
    return t
def empty_like(x, **k):
    t = _orig_empty_like(x, **k)
    if t.is_cuda and t.is_floating_point():
        t.fill_(float("nan"))
        live.append((weakref.ref(t), _site()))
    return t
torch.empty, torch.empty_like = empty, empty_like
reported = set()
_orig_call = E._call
def call(name, *args):
    r = _orig_call(name, *args)
    for ref, site in live:
        t = ref()
        if t is None or id(t) in reported: continue
        n = torch.isnan(t).sum().item()
        if 0 < n < t.numel():
            reported.add(id(t))
            nz = torch.isnan(t).nonzero()
            print(f"[partial] after {name}: {tuple(t.shape)} {t.dtype} nan={n}/{t.numel()} first={nz[0].tolist()} last={nz[-1].tolist()} alloc@ {site}")
    return r
E._call = call

which = sys.argv[1] if len(sys.argv) > 1 else "g"
if which == "g":
    f = load_golden("g13_train.pt")
    net = fb.Generator(13); net.load_state_dict(O.make_state_dict(O.generator_spec(13), f["seed"])); net.to(DEV).train()
    x = f["x"].to(DEV).requires_grad_(True)
    y = net(x)
    print("y err", rel_err(y, f["y"]))
    (y * f["r"].to(DEV)).sum().backward()
    print("dx err", rel_err(x.grad, f["dx"]))
elif which == "s":
    f = load_golden("s4_bilinear_odd.pt")
    net = fb.Segmentor(4, 1, True); net.load_state_dict(O.make_state_dict(O.segmentor_spec(4, 1, True), f["seed"])); net.to(DEV).train()
    x = f["x"].to(DEV).requires_grad_(True); y = f["y"].to(DEV).requires_grad_(True)
    c = net(x, y)
    print("cmap err", rel_err(c, f["cmap"]))
    (c * f["r"].to(DEV)).sum().backward()
    print("dx err", rel_err(x.grad, f["dx"]), rel_err(y.grad, f["dy"]))
elif which == "d":
    f = load_golden("d3_odd.pt")
    net = fb.Discriminator_SRGAN_simple(3); net.load_state_dict(O.make_state_dict(O.discriminator_spec(3), f["seed"])); net.to(DEV).train()
    x = f["x"].to(DEV).requires_grad_(True); y = f["y"].to(DEV).requires_grad_(True)
    c = net(x, y)
    print("out err", rel_err(c, f["out"]))
    (c * f["r"].to(DEV)).sum().backward()
    print("dx err", rel_err(x.grad, f["dx"]), rel_err(y.grad, f["dy"]))
