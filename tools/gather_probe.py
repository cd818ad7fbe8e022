import sys, os as _os; sys.path.insert(0, _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), ".."))
import os, time, numpy as np, torch, torch.distributed as dist
from apples_b200 import parallel
rank=int(os.environ['RANK']); world=int(os.environ['WORLD_SIZE']); lr=int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(lr); dev='cuda:%d'%lr
dist.init_process_group('nccl', device_id=torch.device(dev))
n=125000
loc=(np.arange(n,dtype=np.int32), np.random.rand(n), np.random.rand(n), np.random.rand(n), np.zeros(n,np.int32))
for it in range(6):
    dist.barrier(); torch.cuda.synchronize()
    t0=time.time()
    ts=[torch.from_numpy(a).to(dev, non_blocking=True) for a in loc]; torch.cuda.synchronize(); t1=time.time()
    rec=parallel.pack_records(*ts); torch.cuda.synchronize(); t2=time.time()
    full=parallel.gather_records(rec, n*world); torch.cuda.synchronize(); t3=time.time()
    parts=parallel.unpack_records(full); torch.cuda.synchronize(); t4=time.time()
    out=parallel.gather_placements(loc, n*world, device=dev); t5=time.time()
    if rank==0: print('h2d %.2f pack %.2f gather %.2f unpack %.2f | whole gather_placements %.2f ms'%tuple(1e3*x for x in (t1-t0,t2-t1,t3-t2,t4-t3,t5-t4)), flush=True)
dist.destroy_process_group()
