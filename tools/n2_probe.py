import faulthandler, os, sys
faulthandler.enable()
sys.path.insert(0, os.getcwd())
import torch, torch.distributed as dist
import pdmpflux_b200 as p
rank=int(os.environ["RANK"]); world=int(os.environ["WORLD_SIZE"]); local=int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); p.lib().pdmpflux_set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
print(rank, "pg ok", flush=True)
comm = p.dist.Comm(torch.device("cuda", local))
print(rank, "comm ok", flush=True)
x = torch.full((4, 7), float(rank + 1), dtype=torch.float64, device="cuda")
comm.all_reduce(x, torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
print(rank, x[0,0].item(), flush=True)
comm.close()
dist.destroy_process_group()
