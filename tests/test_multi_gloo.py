"""world_size-2 test of the multi-GPU layer's host logic on CPU (gloo): shard planning with halo,
count / digest reduction and the variable-length gather of sorted record lists.  The per-shard scan is
the ORACLE here (there is no GPU in this container); on the GPU box the same layer runs over NCCL
(tests/test_gpu_parity.py::test_large_stream_properties covers the sharded scan itself)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import TINY_DICT


def _worker(rank, world, port, n_total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, here); sys.path.insert(0, os.path.dirname(here))
    from oracle_lib import Oracle
    from patternmatching_b200 import multi
    o = Oracle(); o.add_dict_bytes(TINY_DICT + b"a\nb\nab\nba\naab\n"); o.compile()
    halo = o.max_pat_len - 1
    sh = multi.plan_shards(n_total, world, halo=halo, align=64)[rank]
    stream = o.gen("ab", sh.lo - sh.halo, sh.n + sh.halo)          # the rank regenerates only its own bytes
    dense = (o.scan(stream)[sh.halo:] + 1).astype(np.uint16)      # halo walked, not reported
    s = o.summary(stream, skip=sh.halo, pos_base=sh.lo)
    red = multi.reduce_summary(dict(positions=s.positions, matches=s.matches, hsum_longest=s.hsum_longest,
                                    hsum_all=s.hsum_all), dist, torch.device("cpu"))
    rec = multi.gather_records(torch.from_numpy(multi.dense_to_records(dense, sh.lo)), dist, torch.device("cpu"))
    if rank == 0:
        q.put((red, rec.numpy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [64 * 3 + 17, 10000])
def test_two_rank_sharded_scan_equals_single_scan(n_total):
    from oracle_lib import Oracle
    from patternmatching_b200 import multi
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + n_total) % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_total, q)) for r in range(2)]
    for p in procs:
        p.start()
    red, rec = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    o = Oracle(); o.add_dict_bytes(TINY_DICT + b"a\nb\nab\nba\naab\n"); o.compile()
    stream = o.gen("ab", 0, n_total)
    s = o.summary(stream)
    assert red == dict(positions=s.positions, matches=s.matches, hsum_longest=s.hsum_longest, hsum_all=s.hsum_all)
    want = multi.dense_to_records((o.scan(stream) + 1).astype(np.uint16), 0)
    assert np.array_equal(rec, want)
    assert np.all(np.diff(rec >> 24) > 0)


def test_plan_shards_cover_and_align():
    from patternmatching_b200 import multi
    for n, w in ((1 << 30, 8), (12345678, 4), (4096, 8), (0, 2)):
        sh = multi.plan_shards(n, w)
        assert sh[0].lo == 0 and sh[-1].hi == n and sh[0].halo == 0
        for a, b in zip(sh[:-1], sh[1:]):
            assert a.hi == b.lo and a.lo % 4096 == 0
        assert all(s.halo == min(352, s.lo) for s in sh)
