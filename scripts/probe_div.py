import torch, numpy as np
p = (torch.rand(1<<20, 3) * 2 - 1) * 1.5
cpu = (p - (-1.5)) / (1.5 - (-1.5))
gpu = ((p.cuda() - (-1.5)) / (1.5 - (-1.5))).cpu()
rec = ((p + 1.5) * np.float32(1.0/3.0))
rec2 = (p + 1.5) * torch.tensor(1.0, dtype=torch.float32).div(3.0)
print("cpu==gpu", float((cpu != gpu).float().mean()), "gpu==p*(1/3 fp32)", float((gpu != rec).float().mean()), float((gpu != rec2).float().mean()))
print("cpu== true div", float((cpu != torch.from_numpy((p.numpy()+np.float32(1.5))/np.float32(3))).float().mean()))
a = torch.rand(1<<20)*5; b=torch.rand(1<<20)+0.5
print("tensor/tensor gpu==cpu", float(((a.cuda()/b.cuda()).cpu() != a/b).float().mean()))
print("x/2.0", float(((a.cuda()/2.0).cpu() != a/2.0).float().mean()))
