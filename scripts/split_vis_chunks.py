"""Split-sum training step against the visibility-round schedule (nerfacc.VISIBILITY_CHUNKS)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rise_sdf_b200 import synthetic as syn, nerfacc
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
for chunks in ((64, 128, 256), (), (128,), (32, 64, 128, 256), (64, 256)):
    nerfacc.VISIBILITY_CHUNKS = chunks
    for _ in range(3):
        trainer.step(*batch, update=False)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(6):
        trainer.step(*batch, update=False)
    torch.cuda.synchronize()
    print(chunks, f"{(time.perf_counter() - t0) / 6 * 1e3:.2f} ms/step")
