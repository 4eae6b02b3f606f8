"""pm_comm_gather_records alone: every rank's sorted records of a 16 GiB S-planted shard (matches of >= 4 pattern bytes)
to rank 0.  torchrun --nproc-per-node N scripts/gather_bench.py [GiB per rank]   (NCCL_* environment as given)"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
import patternmatching_b200 as pm

DATA = os.path.join(ROOT, "oracle", "_ref", "data")
gib = float(sys.argv[1]) if len(sys.argv) > 1 else 16
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
d = pm.Dictionary().add_file(os.path.join(DATA, "snort.dict")).add_file(os.path.join(DATA, "et.dict")).compile()
eng = pm.Engine(d, device=lr)
n = int(gib * (1 << 30))
buf = torch.empty(n, dtype=torch.uint8, device=dev); out = torch.empty(n, dtype=torch.int16, device=dev)
eng.generate("planted", rank * n, n, buf)
cap = n // 256
rec = torch.empty(cap, dtype=torch.int64, device=dev)
cnt = eng.scan_device_records(buf, n, out, rec, cap, min_len=4, pos_base=rank * n)
del buf, out
tot = torch.tensor([cnt], dtype=torch.int64, device=dev); dist.all_reduce(tot)
all_cap = int(tot.item()) + 16
allrec = torch.empty(all_cap if rank == 0 else 1, dtype=torch.int64, device=dev)
comm = pm.Comm.from_torch(dist, lr)
st = torch.cuda.current_stream().cuda_stream
for _ in range(2):
    comm.gather_records(rec, cnt, allrec, all_cap, root=0, cuda_stream=st)
torch.cuda.synchronize(); dist.barrier()
a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5):
    counts, t = comm.gather_records(rec, cnt, allrec, all_cap, root=0, cuda_stream=st)
b.record(); torch.cuda.synchronize()
ms = torch.tensor([a.elapsed_time(b) / 5], dtype=torch.float64, device=dev); dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    inb = 8 * (t - counts[0])
    pos = allrec[:t] >> 24
    print(json.dumps({"ranks": world, "records_total": int(t), "bytes_into_rank0": int(inb), "ms": round(float(ms), 3),
                      "GBps_into_rank0": round(inb / float(ms) / 1e6, 1), "sorted": bool((pos[1:] > pos[:-1]).all().item()),
                      "env": {k: v for k, v in os.environ.items() if k.startswith("NCCL_") or k.startswith("PM_COMM")}}))
comm.free(); dist.barrier(); dist.destroy_process_group()
