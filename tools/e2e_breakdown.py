"""Where does the end-to-end step spend its time?  (run on the GPU box)"""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import recommend_sys_b200 as rs
import bench

train, test = bench.make_data("ml1m_item_pearson_k40")
params = rs.Parameters({"sim": rs.Pearson, "userBased": False, "k": 40})
def t(): return time.perf_counter()
for it in range(4):
    t0 = t(); est = rs.NewKNNWithMean(params); t1 = t()
    est.Fit(train); t2 = t()
    iu = train.convert_users(test.Users); ii = train.convert_items(test.Items); t3 = t()
    out = est._h.predict_batch(ii, iu); t4 = t()
    est.Close(); t5 = t()
    print(f"iter {it}: ctor {1e3*(t1-t0):.2f}  Fit {1e3*(t2-t1):.2f}  convert {1e3*(t3-t2):.2f}  predict {1e3*(t4-t3):.2f}  close {1e3*(t5-t4):.2f} ms")
# inside Fit
h = rs.core._Handle(sim="pearson", knn_type="centered")
for it in range(3):
    t0 = t(); h.fit(train.innerItems, train.innerUsers, train.Ratings, train.ItemCount, train.UserCount, train.GlobalMean); t1 = t()
    print(f"raw fit {1e3*(t1-t0):.2f} ms", h.profile())

print("--- fresh handle per iteration, pieces ---")
for it in range(5):
    t0 = t(); hh = rs.core._Handle(sim="pearson", knn_type="centered"); t1 = t()
    hh.fit(train.innerItems, train.innerUsers, train.Ratings, train.ItemCount, train.UserCount, train.GlobalMean); t2 = t()
    m = hh.means(); t3 = t()
    out = hh.predict_batch(ii, iu); t4 = t()
    hh.close(); t5 = t()
    print(f"iter {it}: create {1e3*(t1-t0):.2f}  fit {1e3*(t2-t1):.2f}  means {1e3*(t3-t2):.2f}  predict {1e3*(t4-t3):.2f}  close {1e3*(t5-t4):.2f} ms")
