"""Throughput of K3's generic path (sample_size > deg + 1: Householder least squares per hypothesis)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rascal_b200 import engine
from tests.conftest import load_golden
from tests.test_gpu_ransac_refit import setup_from_golden

g = load_golden("c3_deg4")
for ssz in (5, 6, 8):
    s = setup_from_golden(g)
    s.sample_size = ssz
    rb = engine.RansacBatch([s])
    n = 20_000_000
    rb.run(0, 1_000_000, seed=1)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    rb.run(0, n, seed=2)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    r = rb.results()[0]
    print("sample_size %d: %.3g hyp/s  counters %s" % (ssz, n / dt, engine.counters_dict(r)))
