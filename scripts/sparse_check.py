"""Dense scan vs sparse mode (in-kernel flags + bitmap compaction) on one device-resident stream (development aid; a
short command line for the ncu launch list).   python scripts/sparse_check.py [bytes] [min_len]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import patternmatching_b200 as pm
DATA = os.path.join(ROOT, "oracle", "_ref", "data")
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 4 << 30
min_len = int(sys.argv[2]) if len(sys.argv) > 2 else 4
d = pm.Dictionary().add_file(os.path.join(DATA, "snort.dict")).add_file(os.path.join(DATA, "et.dict")).compile()
eng = pm.Engine(d)
dev = torch.device("cuda:0")
buf = torch.empty(n, dtype=torch.uint8, device=dev)
out = torch.empty(n, dtype=torch.int16, device=dev)
eng.generate("planted", 0, n, buf)
cap = max(n // 128, 1 << 20)
rec = torch.empty(cap, dtype=torch.int64, device=dev)
for _ in range(2):
    eng.scan_device(buf, n, out)
    cnt = eng.scan_device_records(buf, n, out, rec, cap, min_len=min_len)
torch.cuda.synchronize()
e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
e0.record()
for _ in range(3):
    eng.scan_device(buf, n, out)
e1.record()
for _ in range(3):
    cnt = eng.scan_device_records(buf, n, out, rec, cap, min_len=min_len)
e2.record()
torch.cuda.synchronize()
print(f"{n} bytes: dense {e0.elapsed_time(e1) / 3:.3f} ms, sparse (min_len {min_len}) {e1.elapsed_time(e2) / 3:.3f} ms, {cnt} records", flush=True)
