import sys, time, numpy as np, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from rise_sdf_b200 import synthetic as syn
from rise_sdf_b200.split_mixed_occ import SplitMixedOCCModel, split_mixed_occ_config
from oracle import split as osplit
from helpers import split_oracle_params
torch.manual_seed(0)
cfg = split_mixed_occ_config(); cfg["light"]["envlight_config"]["base_res"] = 64
m = SplitMixedOCCModel(cfg).cuda()
with torch.no_grad():
    m.geometry.network.layers[0].weight_v[:, 3:].normal_(0.0, 0.05); m.variance.variance.fill_(0.5)
    m.geometry.encoding.encoding.encoding.params.uniform_(-0.02, 0.02)
m.eval(); m.update_step(0, 20000)
with torch.no_grad(): m.emitter.build_mips()
R = int(sys.argv[1]) if len(sys.argv) > 1 else 192
rays, rgb, fg, bg = syn.training_rays(R, seed=3); m.background_color = bg.cuda()
grid = syn.analytic_grid('ball'); m.occupancy_grid.binaries = grid[None].cuda()
m.render_step_size = 1.732*2*1.5/256
# FD normal accuracy: fused (no_grad) vs torch fp32 path (grad enabled)
pts = ((torch.rand(20000, 3) * 2 - 1) * 0.9).cuda()
with torch.no_grad(): s1, g1, f1 = m.geometry(pts, with_grad=True, with_feature=True)
m.train(); s2, g2, f2 = m.geometry(pts.clone(), with_grad=True, with_feature=True); m.eval()
print('sdf fused-vs-torch max', float((s1 - s2).abs().max()), 'grad abs err max/mean', float((g1-g2).abs().max()), float((g1-g2).abs().mean()), '|grad| mean', float(g2.norm(dim=-1).mean()))
P = split_oracle_params(m)
t = time.time(); osplit.build_mips(P); print('oracle build_mips', time.time() - t)
for i, (a, b) in enumerate(zip(m.emitter.specular, P.specular)):
    print('mip', i, tuple(b.shape), float((a.detach().cpu() - b).abs().max()))
print('diffuse', float((m.emitter.diffuse.detach().cpu() - P.diffuse).abs().max()))
for rel in (False, True):
    with torch.no_grad(): out = m(rays.cuda(), relighting=rel)
    t = time.time(); ref = osplit.forward(P, rays, grid.numpy(), m.render_step_size, stage=1, relighting=rel, background=bg); print('oracle fwd', time.time() - t, ref['num_samples'], int(out['num_samples'].sum()), len(ref['valid_indices']))
    for k in ('comp_rgb', 'comp_rgb_phys', 'comp_normal', 'opacity', 'depth', 'comp_albedo', 'comp_roughness', 'comp_metallic', 'comp_rgb_full', 'comp_rgb_phys_full', 'comp_spec_rgb_phys'):
        a, b = out[k].cpu().numpy(), ref[k].numpy(); e = np.abs(a - b)
        print(f'  relight={rel} {k:22s} max {e.max():.2e} mean {e.mean():.2e} p99 {np.quantile(e, .99):.2e} scale {np.abs(b).max():.2f}')
