"""Multi-GPU layer: one process per GPU, contiguous shards with a left halo, replicated dictionary.

The path shards naturally: the reference's matcher state after max_pat_len-1 bytes
from the root reports exactly what a continuous scan reports (quirk Q8, Core/src/measure.c:262-306),
so rank r scans global offsets [lo_r, hi_r) plus the HALO bytes before lo_r and needs NO exchange
step.  torch.distributed is used only to gather results: per-rank counts / digests (all_reduce) and,
on request, the position-sorted (pos, pid) record lists (all_gather of counts, then a padded
all_gather -- rank order is position order, so concatenation is already sorted).  Works with the
"nccl" backend on GPU tensors and with "gloo" on CPU tensors (tests).
"""
from dataclasses import dataclass

import numpy as np

from . import HALO


@dataclass
class Shard:
    rank: int
    lo: int          # first global offset this rank reports on
    hi: int          # one past the last
    halo: int        # bytes before lo that the rank must also have (0 for the rank that owns offset 0)

    @property
    def n(self):
        return self.hi - self.lo


def plan_shards(n_total: int, world: int, halo: int = HALO, align: int = 4096):
    """Contiguous shards of (almost) equal size, boundaries aligned to `align` bytes."""
    per = -(-n_total // world)
    per = -(-per // align) * align
    shards = []
    for r in range(world):
        lo, hi = min(r * per, n_total), min((r + 1) * per, n_total)
        shards.append(Shard(r, lo, hi, min(halo, lo)))
    return shards


def reduce_summary(summary: dict, dist, device):
    """Sum positions / matches / digest sums of all ranks (digests are sums mod 2^64: two 32-bit limbs)."""
    import torch
    keys = ["positions", "matches", "hsum_longest", "hsum_all"]
    limbs = []
    for k in keys:
        v = int(summary[k]) & 0xFFFFFFFFFFFFFFFF
        limbs += [v & 0xFFFFFFFF, v >> 32]
    t = torch.tensor(limbs, dtype=torch.int64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t)
    out = {}
    vals = t.cpu().tolist()
    for i, k in enumerate(keys):
        out[k] = (vals[2 * i] + (vals[2 * i + 1] << 32)) & 0xFFFFFFFFFFFFFFFF
    return out


def gather_records(local_records, dist, device):
    """All ranks pass their position-sorted int64 record tensor (pos << 24 | pid); every rank gets the
    concatenation in rank order (= position order).  Variable lengths: all_gather the counts, pad to
    the maximum, all_gather, trim."""
    import torch
    world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        return local_records
    cnt = torch.tensor([local_records.numel()], dtype=torch.int64, device=device)
    counts = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(counts, cnt)
    counts = [int(c.item()) for c in counts]
    m = max(max(counts), 1)
    padded = torch.zeros(m, dtype=torch.int64, device=device)
    padded[:local_records.numel()] = local_records
    parts = [torch.zeros(m, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(parts, padded)
    return torch.cat([p[:c] for p, c in zip(parts, counts)])


def dense_to_records(dense_pids: np.ndarray, pos_base: int) -> np.ndarray:
    """Host helper (tests): dense uint16 result -> (pos << 24 | pid) records of the positions with a match."""
    idx = np.nonzero(dense_pids)[0]
    return ((idx.astype(np.int64) + pos_base) << 24) | dense_pids[idx].astype(np.int64)
