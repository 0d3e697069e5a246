import sys, numpy as np, torch
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from oracle import neus as oneus
from rise_sdf_b200 import synthetic as syn
from rise_sdf_b200.neus import NeuSModel, neus_blender_config
from helpers import oracle_params_from_model, rel_l2
torch.manual_seed(0)
m = NeuSModel(neus_blender_config(), fused_render=True).cuda()
with torch.no_grad(): m.geometry.encoding.encoding.params.uniform_(-0.05,0.05); m.geometry.network.layers[0].weight_v[:, 3:].normal_(0.0, 0.05)
m.train(); m.randomized=False; m.cos_anneal_ratio=0.37
grid = syn.analytic_grid("ball"); m.occupancy_grid.binaries = grid[None].cuda()
m.render_step_size = 1.732*2*1.5/256
rays, rgb, fg, bg = syn.training_rays(256, seed=2); m.background_color = bg.cuda()
out = m(rays.cuda()); loss,_ = oneus.loss(out, rgb.cuda(), fg.cuda()); loss.backward()
res={}
for dt in (torch.float64, torch.float32):
    P = oracle_params_from_model(m).to(dt)
    for t in P.tensors(): t.requires_grad_(True)
    ref = oneus.forward(P, rays, grid.numpy(), m.render_step_size, 0.37, background=bg, training=True, create_graph=True, dtype=dt)
    rl,_ = oneus.loss(ref, rgb.to(dt), fg.to(dt)); rl.backward()
    names=[("geometry.encoding.encoding.params",P.table),("variance.variance",P.variance)]
    for i,l in enumerate(P.geo_mlp):
        for n,t in l.items(): names.append((f"geometry.network.layers.{2*i}.{n}",t))
    for i,l in enumerate(P.tex_mlp):
        for n,t in l.items(): names.append((f"texture.network.layers.{2*i}.{n}",t))
    res[dt]={n:t.grad.double().numpy() for n,t in names}
    print(dt, float(rl), float(loss))
sd=dict(m.named_parameters())
print("%-45s %10s %10s %10s"%("param","gpu-vs-f64","cpu32-vs-f64","|g|"))
for n in res[torch.float64]:
    g=sd[n].grad.cpu().double().numpy(); r64=res[torch.float64][n]; r32=res[torch.float32][n]
    print("%-45s %10.2e %10.2e %10.2e"%(n, rel_l2(g,r64), rel_l2(r32,r64), np.linalg.norm(r64)))
# split layer0 weight_v by column blocks
g=sd["geometry.network.layers.0.weight_v"].grad.cpu().double().numpy(); r=res[torch.float64]["geometry.network.layers.0.weight_v"]; r32=res[torch.float32]["geometry.network.layers.0.weight_v"]
print("xyz cols", rel_l2(g[:,:3],r[:,:3]), rel_l2(r32[:,:3],r[:,:3]), "feat cols", rel_l2(g[:,3:],r[:,3:]), rel_l2(r32[:,3:],r[:,3:]))
for l in range(16):
    print(l, "%.2e %.2e"%(rel_l2(g[:,3+2*l:5+2*l],r[:,3+2*l:5+2*l]), rel_l2(r32[:,3+2*l:5+2*l],r[:,3+2*l:5+2*l])), end=" | ")
