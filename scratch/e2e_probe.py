import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pdmpflux_b200 as p
import bench
print('numa cpus', bench.pin_to_gpu_numa_node(0), len(os.sched_getaffinity(0)))
torch.cuda.set_device(0)
s = bench.make_sampler(p, "c2")
for slab in (None, None, 1 << 30):
    if slab: os.environ["PDMPFLUX_SLAB_BYTES"] = str(slab)
    for i in range(2):
        r = bench.e2e(p, s, "c2", 4096, 1000, 1, torch.device("cuda"))
        print(slab, "%.1f Mev/s %.1f ms" % (r["value"] / 1e6, r["ms_per_step"]), flush=True)
