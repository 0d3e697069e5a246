"""Which torch ops (by name and input shape) still run inside the split-sum training step, by device time."""
import torch
from torch.profiler import ProfilerActivity, profile

from rise_sdf_b200 import synthetic as syn
from rise_sdf_b200.split_mixed_occ import SplitMixedOCCModel, split_mixed_occ_config
from rise_sdf_b200.train import SplitTrainer

dev = torch.device("cuda:0")
torch.manual_seed(42)
model = SplitMixedOCCModel(split_mixed_occ_config()).to(dev)
with torch.no_grad():
    model.geometry.network.layers[0].weight_v[:, 3:].normal_(0.0, 0.05)
model.train()
trainer = SplitTrainer(model)
trainer.global_step = 20001
model.update_step(0, 20000)
gj = torch.Generator().manual_seed(7)
model.occupancy_grid._update(0, model.occ_eval_fn, occ_thre=0.001, jitter=torch.rand(128 ** 3, 3, generator=gj))
batch = tuple(t.to(dev) for t in syn.training_rays(4096, seed=7))
for _ in range(4):
    trainer.step(*batch, update=False)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True, with_stack=False) as prof:
    trainer.step(*batch, update=False)
    torch.cuda.synchronize()
print(prof.key_averages(group_by_input_shape=True).table(sort_by="device_time_total", row_limit=70, max_name_column_width=48,
                                                         max_shapes_column_width=70))
print("==== by host time, no shape grouping")
ka = prof.key_averages()
for e in sorted(ka, key=lambda e: -e.self_cpu_time_total)[:70]:
    print(f"{e.self_cpu_time_total / 1e3:8.2f} ms self-host  {e.count:5d} x  {e.key[:90]}")
