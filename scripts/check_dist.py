"""Run under torchrun (N ranks, NCCL): the sharded single-spectrum fit must not depend on the number of GPUs.
Every rank runs fit(max_tries=...) and fit(n_hypotheses=...) through the facade with use_distributed(); rank 0
repeats them alone (no process group use) and compares winner, counters and completed tries."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from bench import c3_input
from rascal_b200.calibrator import Calibrator, LineList

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
w = c3_input()


def new(seed, distributed):
    c = Calibrator(w["peaks"], spectrum=np.zeros(w["npix"]))
    c.set_calibrator_properties(seed=seed)
    c.set_hough_properties(**w["hough"])
    c.set_ransac_properties(**w["ransac"])
    c.set_atlas(LineList(w["atlas"]), candidate_tolerance=10.0)
    if distributed:
        c.use_distributed(None)
    return c


out = {}
for name, kw in (("max_tries", dict(max_tries=700)), ("n_hypotheses", dict(n_hypotheses=3_000_001))):
    c = new(11, True)
    c.do_hough_transform()
    r = c.fit(progress=False, **kw, **w["fit"])
    out[name] = (c.best_hypothesis[0], c.best_cost, dict(c.ransac_counters), getattr(c, "completed_tries", None), np.array(r[0]))
dist.barrier()
if rank == 0:
    for name, kw in (("max_tries", dict(max_tries=700)), ("n_hypotheses", dict(n_hypotheses=3_000_001))):
        c = new(11, False)
        c.do_hough_transform()
        r = c.fit(progress=False, **kw, **w["fit"])
        a = out[name]
        assert a[0] == c.best_hypothesis[0] and a[1] == c.best_cost, (name, a[:2], c.best_hypothesis[0], c.best_cost)
        assert a[2] == c.ransac_counters, (name, a[2], c.ransac_counters)
        assert a[3] == getattr(c, "completed_tries", None)
        assert np.array_equal(a[4], np.array(r[0]))
        print("%s: %d ranks == 1 rank: winner %d cost %.6e counters %s tries %s" % (name, world, a[0], a[1], a[2], a[3]))
dist.barrier()
dist.destroy_process_group()
