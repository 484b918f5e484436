"""Token construction (csrc/tokens.cu) against the reference's per-scene loop run with PyTorch ops on the same GPU, B=32 (GPU box)."""
import torch, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from situation3d_b200 import tokens
g = torch.Generator().manual_seed(0)
B, m, C = 32, 2400, 256
coords, feats = [], []
for s in range(B):
    xy = torch.randint(-20, 20, (m, 2), generator=g) * 16; z = torch.randint(0, 6, (m, 1), generator=g) * 16
    c = torch.unique(torch.cat([xy, z], 1).to(torch.int32), dim=0)
    coords.append(c.cuda()); feats.append(torch.randn(c.shape[0], C, generator=g).cuda())
flush = torch.empty(1 << 28, dtype=torch.float32, device='cuda')
def run():
    return tokens.column_tokens(coords, feats, [16, 16, 16], 256, 0.02)
run(); torch.cuda.synchronize()
import time
ts = []
for _ in range(5):
    torch.cuda.synchronize(); t0 = time.perf_counter(); run(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
print("column_tokens B=32, ~%d voxels/scene, C=256: %.3f ms wall (incl. the count read-back)" % (sum(c.shape[0] for c in coords) // B, 1e3 * sorted(ts)[2]))
# the reference's loop on the GPU for comparison (same ops as sqa_module.py:297-315)
def ref_loop():
    sf, sp = [], []
    for c, f in zip(coords, feats):
        rc = c[:, [0, 1]]
        u, idx = rc.unique(dim=0, return_inverse=True)
        r = torch.zeros(u.size(0), f.size(1), device=f.device).scatter_reduce_(0, idx.unsqueeze(-1).expand_as(f), f, reduce='mean')
        si = torch.randperm(u.size(0))[:256] if 256 < u.size(0) else torch.cat([torch.randperm(u.size(0)), torch.randint(0, u.size(0), (256 - u.size(0),))])
        sf.append(r[si].unsqueeze(0)); sp.append(((u[si] + torch.tensor([16, 16], device=f.device) / 2) * 0.02).unsqueeze(0))
    return torch.cat(sf), torch.cat(sp)
ref_loop(); torch.cuda.synchronize(); ts = []
for _ in range(5):
    torch.cuda.synchronize(); t0 = time.perf_counter(); ref_loop(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
print("reference loop (PyTorch ops on the same GPU): %.3f ms wall" % (1e3 * sorted(ts)[2]))
