"""Mid-size check of the symmetric-slab top-k Fit (1, 2, 3 shards) against the matrix-store TopK.
Run on the GPU box: python tools/check_symmetric_slabs.py [slab_rows].  It caught a block-level race
in the top-k kernels that the ml-100k tests were too small to hit."""
import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import recommend_sys_b200 as rs
from recommend_sys_b200.shard import union_topk_device
os.environ["RS_KNN_SLAB_ROWS"] = sys.argv[1] if len(sys.argv) > 1 else "1024"
d = rs.core.synth_ratings(20000, 4000, 1_500_000, 0x5EED0009)
train = rs.NewTrainSet(d)
n, k = train.UserCount, 100
full = rs.NewKNN(rs.Parameters({"sim": rs.MSD, "userBased": True, "simPath": "tensor"}))
full.Fit(train)
wi, ws = full.TopK(k)
full.Close()
base = {"sim": rs.MSD, "userBased": True, "store": "topk", "topk": k, "simPath": "tensor"}
def run(S):
    parts = []
    for r in range(S):
        p = rs.NewKNN(rs.Parameters(dict(base, shardCount=S, shardIndex=r)))
        p.Fit(train); parts.append(p.TopK(k)); p.Close()
    if S == 1: return parts[0]
    ui, us = union_topk_device(torch.from_numpy(np.stack([a for a, _ in parts])).cuda(), torch.from_numpy(np.stack([b for _, b in parts])).cuda())
    torch.cuda.synchronize()
    return ui.cpu().numpy(), us.cpu().numpy()
for S in (1, 1, 2, 3):
    gi, gs = run(S)
    bad = np.where((gi != wi).any(axis=1))[0]
    print("shards", S, "rows differing from matrix TopK:", len(bad), bad[:10])
    if len(bad):
        r = bad[0]; c = np.where(gi[r] != wi[r])[0][:5]
        print("  row", r, "cols", c, "got", gi[r][c], gs[r][c], "want", wi[r][c], ws[r][c])
