"""K6 alone on rendered C4 arcs (for ncu): python scripts/prof_peaks.py [n_arcs]"""
import os
import sys


sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from rascal_b200 import util  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
arcs = [bench.c4_arc(bench.c4_input(i), i) for i in range(n)]
pb = util.PeakBatch(arcs)
import torch  # noqa: E402

for rep in range(2):
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record()
    pb.find(height=150.0, distance=3, prominence=100.0, max_peaks=256)
    e[1].record()
    pb.refine(5)
    e[2].record()
    torch.cuda.synchronize()
    print("find %.3f ms  refine %.3f ms" % (e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])))
print(sum(len(r) for r in pb.fetch_refined()), "refined peaks")
