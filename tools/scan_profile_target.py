"""Small target for ncu: config 3 shape (10M x 128 fp32, L2, batch 1) through the streaming scan path, 12 queries."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import _pkg  # noqa: E402

_pkg.load()
from the_algorithm_b200.ann.brute_force import BruteForceIndex  # noqa: E402
from the_algorithm_b200.ann.common import FuturePool, L2  # noqa: E402

dev = torch.device("cuda", 0)
n, d = 10_000_000, 128
g = torch.Generator(device=dev)
g.manual_seed(1)
ix = BruteForceIndex(L2, FuturePool.immediate_pool(), capacity_hint=n)
for c0 in range(0, n, 1_000_000):
    ix.append_batch_device(torch.arange(c0, c0 + 1_000_000, device=dev, dtype=torch.int64),
                           torch.randn((1_000_000, d), generator=g, device=dev) / d ** 0.5)
ix.set_option("path", 1)
q = (torch.rand((12, d), generator=g, device=dev) * 2 - 1).cpu().numpy()
for i in range(12):
    ids, dist, cnt = ix.batch_query_with_distance(q[i:i + 1], 100)
print("ok", ids[0, :3], dist[0, :3])
