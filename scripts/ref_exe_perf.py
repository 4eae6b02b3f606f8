"""The reference's own program (main.c / measure.c, patched per INTEGRATION.md and linked against libpm_b200.so:
oracle/_ref/exe_gpu_big, 16 MiB chunks, page-locked static buffers) on a synthetic stream file: its own CSV, i.e. its own
clock() around the matching loop and its own success rates against its reliable Aho-Corasick.

    python scripts/ref_exe_perf.py [MiB]        (default 256; the CPU rows make three passes at ~20 MB/s)
"""
import json, os, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import patternmatching_b200 as pm

DATA = os.path.join(ROOT, "oracle", "_ref", "data")
mib = int(sys.argv[1]) if len(sys.argv) > 1 else 256
n = mib << 20
d = pm.Dictionary().add_file(os.path.join(DATA, "snort.dict")).add_file(os.path.join(DATA, "et.dict")).compile()
eng = pm.Engine(d)
buf = torch.empty(n, dtype=torch.uint8, device="cuda:0")
eng.generate("planted", 0, n, buf)
tmp = tempfile.mkdtemp()
path = os.path.join(tmp, "planted.stream")
buf.cpu().numpy().tofile(path)
del buf, eng
out = os.path.join(tmp, "out.csv")
open(out, "w").close()
# The reference times its rows with clock() (measure.c:290-297): CPU seconds of the WHOLE process, i.e. of every thread.
# With the ids translated by the engine's host threads (the default, fastest by the wall clock) its column shows the sum
# of their CPU time; PM_HOST_IDS=device (ids translated on the GPU, arriving by DMA: no host thread works) shows what a
# single-threaded caller sees.  Both are reported.
res = {"stream_MiB": mib, "modes": {}}
for mode in ("host", "device"):
    env = dict(os.environ, PM_HOST_IDS=mode)
    open(out, "w").close()
    t0 = time.time()
    r = subprocess.run([os.path.join(ROOT, "oracle", "_ref", "exe_gpu_big"), "-v", "-d", os.path.join(DATA, "snort.dict"), "-d", os.path.join(DATA, "et.dict"),
                        "-s", path, "-o", out], capture_output=True, text=True, env=env)
    wall = time.time() - t0
    rows = [l.split(",") for l in open(out).read().splitlines()]
    m = {"exe_wall_seconds": round(wall, 1), "rc": r.returncode, "rows": []}
    for row in rows[1:]:
        secs = float(row[1])
        m["rows"].append({"algorithm": row[0], "time_secs_clock": secs, "GBps_by_its_own_clock": round(n / secs / 1e9, 3) if secs > 0 else None,
                          "false_pos": float(row[3]), "false_neg": float(row[4]), "partial": float(row[5]), "total_mem": int(row[2])})
    res["modes"]["PM_HOST_IDS=" + mode] = m
    if r.returncode:
        print(r.stdout[-1500:], r.stderr[-1500:])
print(json.dumps(res, indent=1))
