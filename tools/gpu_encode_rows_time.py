import sys, os, time
sys.path.insert(0, os.getcwd())
import torch
from btb_phylo_b200 import _capi, build, device as dev
import bench
build.build()
n, L = 50000, 30000
seqs = bench.make_alignment_cuda(n, L, torch.device("cuda", 0))
plan = dev.DistancePlan(seqs, length=L, engine=_capi.ENGINE_UMMA)
def ms(fn, reps=5):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
print("full encode        %.3f ms" % ms(plan.encode))
for r0 in (0, 256, 14080, 25000, 40000):
    print("rows [%5d, n)     %.3f ms" % (r0, ms(lambda: plan.encode(r0, n))))
