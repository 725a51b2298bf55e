"""Repeat a statistical GPU test in one process and print the outcome of every repetition (how often, by how much):
python scripts/flake_probe.py <reps> <test_module>::<test_function>[arg] ...  with engine switches from FCD_ENGINE / --set."""
import importlib
import sys
import traceback

import torch

sys.path.insert(0, ".")
from fcdgan_b200 import engine as E  # noqa: E402

reps = int(sys.argv[1])
for spec in sys.argv[2:]:
    mod, _, rest = spec.partition("::")
    fn, _, arg = rest.partition("[")
    arg = arg.rstrip("]")
    f = getattr(importlib.import_module(mod), fn)
    for batched in (True, False):
        E.set_batch_branches(batched)
        fails = []
        for i in range(reps):
            try:
                f(arg) if arg else f()
            except AssertionError as e:
                fails.append(str(e).splitlines()[0][:160])
            except Exception:
                fails.append(traceback.format_exc().splitlines()[-1][:160])
            torch.cuda.synchronize()
        print(f"{spec} batch_branches={batched}: {len(fails)}/{reps} failed", flush=True)
        for m in fails:
            print("   ", m, flush=True)
