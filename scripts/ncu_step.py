"""One eager iteration of a bench.py workload (default config 2: G + D fwd/bwd + optimizer steps, batch 16 of 13 x 256 x 256)
between cudaProfilerStart/Stop, for the ncu launch list:
  ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file launches.csv \
      python scripts/ncu_step.py [config] [precision] [streams]
(bench.py replays the same launches from one CUDA graph; the list is taken in eager mode so every kernel is visible)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import fcdgan_b200 as fb
from fcdgan_b200.steps import drive

config = sys.argv[1] if len(sys.argv) > 1 else "2"
dev = torch.device("cuda:0")
fb.set_precision(sys.argv[2] if len(sys.argv) > 2 else "parity")
fb.set_streams(int(sys.argv[3]) if len(sys.argv) > 3 else 1)
B = bench.CONFIGS[config]["B"]
genfn, nets, _ = bench.build_ours(config, dev, B, False)
data = bench.synth_for(config, B, 1234, device=dev)
for _ in range(3):
    drive(genfn(*data))
torch.cuda.synchronize()
torch.cuda.profiler.start()
drive(genfn(*data))
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
