#!/usr/bin/env python
"""ref_digest.py -- TEST / BASELINE INFRASTRUCTURE: the reference's own Aho-Corasick (oracle/_ref/libpmref.so, unmodified
sources) on a seeded synthetic stream, in a process of its own (the reference keeps global state: one dictionary per
process).  Prints one JSON line with positions, matches and the two digest sums of the first --bytes bytes.

    python oracle/ref_digest.py --dict a.dict [--dict b.dict] --kind ab|uniform|planted|almost|ascii --bytes N [--workers K]
"""
import argparse
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dict", action="append", required=True)
    ap.add_argument("--kind", default="ab")
    ap.add_argument("--bytes", type=int, default=64 << 20)
    ap.add_argument("--offset", type=int, default=0)
    ap.add_argument("--workers", type=int, default=os.cpu_count() or 1)
    a = ap.parse_args()
    from reflib import Reference
    from oracle_lib import Oracle
    ref = Reference(a.dict, algo_mask=1)
    gen = Oracle()                      # regenerates the same seeded stream on the CPU
    for p in a.dict:
        gen.add_dict_file(p)
    gen.compile()
    stream = gen.gen(a.kind, a.offset, a.bytes)
    r = ref.scan_parallel(stream, a.workers)
    print(json.dumps({"positions": r["positions"], "matches": r["matches"], "hsum_longest": r["hsum_longest"],
                      "hsum_all": r["hsum_all"], "n_patterns": ref.n_patterns, "seconds": r["max_loop_seconds"]}))


if __name__ == "__main__":
    main()
