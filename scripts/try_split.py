import sys, time, torch
sys.path.insert(0, '/root/repo')
from rise_sdf_b200 import synthetic as syn
from rise_sdf_b200.split_mixed_occ import SplitMixedOCCModel, split_mixed_occ_config
torch.manual_seed(0)
cfg = split_mixed_occ_config()
cfg["light"]["envlight_config"]["base_res"] = int(sys.argv[1]) if len(sys.argv) > 1 else 64
m = SplitMixedOCCModel(cfg).cuda()
with torch.no_grad():
    m.geometry.network.layers[0].weight_v[:, 3:].normal_(0.0, 0.05)
    m.variance.variance.fill_(0.5)
m.train(); m.update_step(0, 20000); print('stage', m.stage, 'eps', m.geometry._finite_difference_eps)
t=time.time(); m.emitter.build_mips(); torch.cuda.synchronize(); print('build_mips', time.time()-t, [tuple(s.shape) for s in m.emitter.specular])
t=time.time(); m.emitter.build_mips(); torch.cuda.synchronize(); print('build_mips again', time.time()-t)
rays, rgb, fg, bg = syn.training_rays(1024, seed=3); m.background_color = bg.cuda()
m.randomized = False
m.occupancy_grid.binaries = syn.analytic_grid('ball')[None].cuda()
# eval / relight
m.eval()
with torch.no_grad():
    for rel in (False, True):
        t=time.time(); out = m(rays.cuda(), relighting=rel); torch.cuda.synchronize(); print('eval relight', rel, time.time()-t, {k: tuple(v.shape) for k,v in out.items() if k.startswith('comp_rgb')}, float(out['comp_rgb_full'].mean()), float(out['opacity'].mean()))
# train step
m.train(); m.randomized=False
out = m(rays.cuda())
loss = ((out['comp_rgb_full'] - rgb.cuda())**2).mean() + ((out['comp_rgb_phys_full'] - rgb.cuda())**2).mean() + 0.1*((out['sdf_grad_samples'].norm(dim=-1)-1)**2).mean() + out['sdf_laplace_samples'].mean()*0.01 + out['normals_orientation_loss_map'].mean()
loss.backward(); torch.cuda.synchronize()
print('train loss', float(loss), 'grads:', {n: float(p.grad.norm()) for n,p in m.named_parameters() if p.grad is not None and ('params' in n or 'base' in n or 'variance' in n or 'layers.0.weight' in n)})
