"""Developer tool (torchrun, N GPUs): cost of the pieces of the sharded layer's exchanges on the Reddit-shape slab:
symmetric-memory barrier, local top-k vs top-k writing to peers (unicast / multicast), NVLS reduce vs NCCL
reduce_scatter.   torchrun --nproc-per-node N tools/exchange_lab.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "spgemm-prunning_b200")):
    sys.path.insert(0, p)
import torch, torch.distributed as dist
import maxk_cuda_kernels as K
from sharded import PeerGather, PeerReduce

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n, k = 232965, 32
m = (n + world - 1) // world
x = torch.rand(m, 256, device=dev)

def timed(fn, reps=30):
    for _ in range(5): fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / reps], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)

pg = PeerGather(world * m, k, dev, None, True)
pg_uc = PeerGather(world * m, k, dev, None, False)
res = {}
res["barrier"] = timed(lambda: pg.sets[0]["hv"].barrier(channel=0))
res["topk_local"] = timed(lambda: K.topk_cbsr(x, k))
st = pg.sets[0]
res["topk_multicast"] = timed(lambda: K.topk_cbsr_to_peers(x, k, st["val_ptrs"], st["sel_ptrs"], rank * m, mc_val_ptr=st["mc_val"], mc_sel_ptr=st["mc_sel"]))
su = pg_uc.sets[0]
res["topk_unicast"] = timed(lambda: K.topk_cbsr_to_peers(x, k, su["val_ptrs"], su["sel_ptrs"], rank * m))
res["topk_multicast+barrier"] = timed(lambda: (K.topk_cbsr_to_peers(x, k, st["val_ptrs"], st["sel_ptrs"], rank * m, mc_val_ptr=st["mc_val"], mc_sel_ptr=st["mc_sel"]), st["hv"].barrier(channel=0)))
vals = torch.empty(m, k, device=dev); sel = torch.empty(m, k, dtype=torch.uint8, device=dev)
vf = torch.empty(world * m, k, device=dev); sf = torch.empty(world * m, k, dtype=torch.uint8, device=dev)
res["topk_local+2xall_gather"] = timed(lambda: (K.topk_cbsr(x, k, out_values=vals, out_sel=sel), dist.all_gather_into_tensor(vf, vals), dist.all_gather_into_tensor(sf, sel)))
pr = PeerReduce(world * m, k, dev, None)
gs = torch.empty(m, k, device=dev)
s0 = pr.sets[0]
res["nvls_reduce"] = timed(lambda: K.nvls_reduce(s0["mc"] + rank * m * k * 4, gs))
res["barrier+nvls_reduce"] = timed(lambda: (s0["h"].barrier(channel=0), K.nvls_reduce(s0["mc"] + rank * m * k * 4, gs)))
part = torch.zeros(world * m, k, device=dev)
res["nccl_reduce_scatter"] = timed(lambda: dist.reduce_scatter_tensor(gs, part))
res["memset_partial"] = timed(lambda: s0["buf"].zero_())
if rank == 0:
    print("world %d, slab %d rows:" % (world, m), {k_: round(v * 1e3, 1) for k_, v in res.items()}, "(microseconds)")
dist.destroy_process_group()
