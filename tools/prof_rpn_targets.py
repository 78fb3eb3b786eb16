"""Per-kernel CUDA times of build_rpn_targets at B=8 (developer aid, torch.profiler)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from objectdetection_b200 import config
from objectdetection_b200.data_processor import PreprareTrainData
from torch.profiler import profile, ProfilerActivity
conf = config(); rs = np.random.RandomState(0)
P8 = PreprareTrainData(conf); A = P8.anchors.shape[0]; anc = P8.anchors.cpu().numpy()
g8 = np.zeros((8, 100, 4))
for b in range(8):
    g8[b] = np.round(np.clip(anc[rs.choice(A, 100, replace=False)] + rs.normal(0, 3, (100, 4)), 0, 1024))
    bad = (g8[b, :, 2] <= g8[b, :, 0]) | (g8[b, :, 3] <= g8[b, :, 1]); g8[b, bad] = [100, 100, 164, 164]
g8c = torch.from_numpy(g8).cuda()
pp8 = torch.stack([torch.randperm(A, device="cuda") for _ in range(8)]).int()
for _ in range(3): P8.build_rpn_targets(g8c, perm_pos=pp8, perm_neg=pp8)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(5): P8.build_rpn_targets(g8c, perm_pos=pp8, perm_neg=pp8)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=12, max_name_column_width=60))
