"""gpu_read_block with every rank of a box calling at once: aggregate GB/s per PM_HOST_IDS mode / thread count.
    torchrun --nproc-per-node N scripts/e2e_ranks.py [MiB per rank] [modes, e.g. host,2,3,device] [threads per rank, 0 = default]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
import patternmatching_b200 as pm

DATA = os.path.join(ROOT, "oracle", "_ref", "data")
mib = int(sys.argv[1]) if len(sys.argv) > 1 else 128
modes = sys.argv[2].split(",") if len(sys.argv) > 2 else ["host", "device"]
threads = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0]
rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
d = pm.Dictionary().add_file(os.path.join(DATA, "snort.dict")).add_file(os.path.join(DATA, "et.dict")).compile()
gen = pm.Engine(d, device=lr)
n = mib << 20
buf = torch.empty(n, dtype=torch.uint8, device=dev)
gen.generate("planted", rank * n, n, buf)
hin = pm.PinnedBuffer(n); hout = pm.PinnedBuffer(8 * n)
hin.array(np.uint8)[:] = buf.cpu().numpy()
pats = [d.pattern(pid)[4] for pid in range(1, d.n_patterns + 1)]
res = {}
for t in threads:
    for mode in modes:
        os.environ["PM_HOST_IDS"] = mode
        if t:
            os.environ["PM_HOST_THREADS"] = str(t)
        m = pm.MpsGpu("sfx")
        for i, p in enumerate(pats):
            m.add_pattern(p, 0x7F0000000000 + 64 * (i + 1))
        m.compile()
        for _ in range(2):
            m.reset(); m.read_block_ptr(hin.ptr, n, hout.ptr)
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(4):
            m.reset(); m.read_block_ptr(hin.ptr, n, hout.ptr)
        dt = torch.tensor([(time.perf_counter() - t0) / 4], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        res[f"threads={t or 'default'} ids={mode}"] = round(world * n / float(dt.item()) / 1e9, 2)
        m.free()
if rank == 0:
    print(json.dumps({"ranks": world, "MiB_per_rank": mib, "cores": os.cpu_count(), "aggregate_GBps": res}, indent=1))
if world > 1:
    dist.barrier(); dist.destroy_process_group()
