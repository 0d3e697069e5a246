import sys, time, torch
sys.path.insert(0, '/root/repo')
from rise_sdf_b200 import synthetic as syn
from rise_sdf_b200.relight import EnvSet, render_frame_shard, synthetic_envs
from rise_sdf_b200.split_mixed_occ import SplitMixedOCCModel, split_mixed_occ_config
dev = torch.device('cuda')
torch.manual_seed(42)
model = SplitMixedOCCModel(split_mixed_occ_config()).to(dev)
with torch.no_grad():
    model.geometry.network.layers[0].weight_v[:, 3:].normal_(0.0, 0.05); model.variance.variance.fill_(0.5)
model.train(); model.update_step(0, 80000)
model.occupancy_grid._update(0, model.occ_eval_fn, occ_thre=0.001, jitter=torch.rand(128**3, 3, generator=torch.Generator().manual_seed(7)))
model.eval(); model.background_color = torch.ones(3, device=dev)
t=time.time(); envs = EnvSet(model, synthetic_envs()); torch.cuda.synchronize(); print('EnvSet (2 maps: latlong->cube + build_mips)', time.time()-t)
rays = syn.frame_rays(3).to(dev)
tile = rays[320000-16384:320000+16384].contiguous()
envs.use(0)
with torch.no_grad():
    for _ in range(2): o = model.forward_(tile, relighting=True)
    torch.cuda.synchronize(); t=time.time(); o = model.forward_(tile, relighting=True); torch.cuda.synchronize(); print('tile 32768 rays', time.time()-t, 'samples', int(o['num_samples']))
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        o = model.forward_(tile, relighting=True); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))
