"""Relit frames/s against the tile size (bench.py's relight leg, variance 0.5)."""
import argparse, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from rise_sdf_b200 import relight, nerfacc
nerfacc.KEEP_BUDGET_BYTES = int(os.environ.get('KEEP_GB', '1')) << 30
torch.cuda.set_device(0)
args = argparse.Namespace(steps=20, warmup=5)
for mt in [int(a) for a in sys.argv[1:]] or [32768, 65536, 131072]:
    relight.MAX_TILE = mt
    r = bench.run_relight(args, torch.device("cuda:0"), 1, 0, variance=0.5)
    print("mean_rgb", repr(r["mean_rgb"]))
    print(mt, round(r["value"], 3), "frames/s", round(r["ms_per_frame"], 1), "ms/frame; e2e", round(r["e2e"]["value"], 3),
          "reserved GB", round(torch.cuda.memory_reserved() / 2**30, 1))
    torch.cuda.empty_cache()
