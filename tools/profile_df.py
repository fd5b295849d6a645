import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
r = bench.distance_field_microbench(torch.device("cuda", 0), False)
print({k: r[k] for k in ("value", "ms_per_field", "relaxation_launches", "reachable_cells")})
