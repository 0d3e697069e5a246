"""bench.py's split-train leg on its own (value, e2e, launches)."""
import argparse, json, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
torch.cuda.set_device(0)
args = argparse.Namespace(steps=int(sys.argv[1]) if len(sys.argv) > 1 else 20, warmup=5)
for _ in range(2):
    r = bench.run_split_train(args, torch.device("cuda:0"), 1, 0)
    print(json.dumps({k: r[k] for k in ("value", "ms_per_step", "e2e", "gpu_launches_per_step", "primary_samples_per_step")}))
    torch.cuda.empty_cache()
