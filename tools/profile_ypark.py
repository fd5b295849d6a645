"""Y-type parking sweep stage as bench.py times it: 256 sweeps x 858 candidates, host arrays in."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
torch.cuda.set_device(0)
r = bench.ypark_microbench(None, torch.device("cuda", 0), False)
print({k: (round(v, 3) if isinstance(v, float) else v) for k, v in r.items() if not isinstance(v, dict)})
