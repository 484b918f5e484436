"""scene_feat_linear (Linear 256->768 + GELU over 32 x 256 tokens): fused tcgen05 kernel vs the PyTorch modules, device time (GPU box)."""
import os
import torch, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from situation3d_b200.heads import SceneFeatLinear
m = SceneFeatLinear().eval().cuda(); x = torch.randn(32, 256, 256, device='cuda')
flush = torch.empty(1 << 28, dtype=torch.float32, device='cuda')
def t(fn):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(5):
        flush.zero_(); s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); e.synchronize(); ts.append(s.elapsed_time(e))
    return sorted(ts)[2] * 1e3
with torch.no_grad():
    a = t(lambda: m(x))
    ref = torch.nn.Sequential(torch.nn.Linear(256, 768), torch.nn.GELU()).cuda().eval()
    b = t(lambda: ref(x))
print("scene_feat_linear B=32 x 256 tokens: fused tcgen05 %.1f us (%.0f TFLOP/s), PyTorch Linear + GELU (cuBLAS fp32 + eltwise) %.1f us" % (a, 2 * 8192 * 256 * 768 / a / 1e6, b))
