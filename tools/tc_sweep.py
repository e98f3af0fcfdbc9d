"""Tensor-kernel tuning sweep (run on the GPU box):  python tools/tc_sweep.py WORKLOAD "cluster:sup:debug" ...
Prints the similarity-kernel time and the int8 roofline fraction per setting.  debug != 0 are
timing experiments (results invalid): 1 = no TMA loads, 2 = no MMAs."""
import json, os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch
import bench
import recommend_sys_b200 as rs

wl = sys.argv[1]
users, items, nnz, n_test, sim, knn_type, user_based, k = bench.WORKLOADS[wl]
train, test = bench.make_data(wl)
n_left = train.UserCount if user_based else train.ItemCount
n_right = train.ItemCount if user_based else train.UserCount
left = train.innerUsers if user_based else train.innerItems
right = train.innerItems if user_based else train.innerUsers
dev = torch.device("cuda", 0)
d_left = torch.from_numpy(left).to(dev); d_right = torch.from_numpy(right).to(dev)
d_rating = torch.from_numpy(train.Ratings).to(dev)
peak = 2 * json.loads((Path(bench.ROOT) / "MEASURED_PEAKS.json").read_text())["bf16_tflops"]
g = {"cosine": 3, "msd": 4, "pearson": 6}[sim]
ops = n_left * (n_left - 1) / 2 * 2 * g * n_right
for spec in sys.argv[2:]:
    parts = spec.split(":")
    cl, sup, dbg = parts[0], (parts[1] if len(parts) > 1 else ""), (parts[2] if len(parts) > 2 and parts[2] else "0")
    os.environ["RS_KNN_TC_CLUSTER"] = cl; os.environ["RS_KNN_TC_SUP"] = sup or "x"; os.environ["RS_KNN_TC_DEBUG"] = dbg
    if len(parts) > 3:
        os.environ["RS_KNN_TC_SYNC"] = parts[3]
    else:
        os.environ.pop("RS_KNN_TC_SYNC", None)
    h = rs.core._Handle(sim=sim, knn_type=knn_type, k=k, device=0, pearson_mode="sums", sim_path="tensor")
    for it in range(3):
        h.fit_device(d_left.data_ptr(), d_right.data_ptr(), d_rating.data_ptr(), len(left), n_left, n_right, train.GlobalMean)
        h.synchronize()
        if it == 0: h.profile_reset()
    p = h.profile()
    ms = p["sim_kernel_ms"] / max(1, p["sim_launches"])
    print(f"{wl} {spec}: sim {ms:.3f} ms  {ops / ms / 1e9:.1f} TOP/s  frac {ops / ms / 1e9 / peak:.3f}", flush=True)
    h.close()
