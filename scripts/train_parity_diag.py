"""Row-layout training path vs the reference wiring at full size: per-parameter gradient differences (GPU box)."""
import copy, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from situation3d_b200.backbone_module import Pointnet2Backbone
from situation3d_b200.synthetic import make_batch
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = os.environ.get("CUDNN_TF32", "0") == "1"
torch.manual_seed(0)
net = Pointnet2Backbone(input_feature_dim=129).cuda()
a, b = copy.deepcopy(net).train(), copy.deepcopy(net).train()
b.train_layout = "reference"
pc = torch.from_numpy(make_batch(2, 40000, 129)).cuda()
oa, ob = a({"point_clouds": pc}), b({"point_clouds": pc})
la, lb = oa["fp2_features"].square().mean(), ob["fp2_features"].square().mean()
la.backward(); lb.backward()
print("loss", float(la), float(lb))
# a second run of the reference wiring itself: its scatter-adds are float atomics, so it differs from ITSELF run to run
b2 = copy.deepcopy(net).train()
b2.train_layout = "reference"
lb2 = b2({"point_clouds": pc})["fp2_features"].square().mean()
lb2.backward()
self_rows = []
for (n1, p1), (_, p2) in zip(b2.named_parameters(), b.named_parameters()):
    d = (p1.grad - p2.grad)
    self_rows.append((float(d.abs().max()) / (float(p2.grad.abs().max()) + 1e-30), float(d.norm() / (p2.grad.norm() + 1e-30)), n1))
print("reference wiring vs itself: worst max-rel %.3e (%s), worst l2-rel %.3e" % (max(self_rows)[0], max(self_rows)[2], max(r[1] for r in self_rows)))
rows = []
for (n1, p1), (_, p2) in zip(a.named_parameters(), b.named_parameters()):
    d = (p1.grad - p2.grad)
    rows.append((float(d.abs().max()) / (float(p2.grad.abs().max()) + 1e-30), float(d.norm() / (p2.grad.norm() + 1e-30)), float(p2.grad.abs().max()), n1))
for r in sorted(rows, reverse=True)[:12]:
    print("max-rel %.3e  l2-rel %.3e  max|grad| %.3e  %s" % r)
for k in ("sa1_features", "sa2_features", "sa4_features", "fp2_features"):
    print(k, float((oa[k] - ob[k]).abs().max()), float(ob[k].abs().max()))
