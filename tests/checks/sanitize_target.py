"""Tiny target for `compute-sanitizer --tool memcheck` (run under gpurun, wrapped in `timeout`): every kernel of the library on
small shapes -- append / update, prep, streaming scan, tensor-core filter (cta_group 1 and 2, seed + chunks + resident-query
mode), compaction, finalize, exact fallback, merge.  Exits non-zero if a result differs from the CPU oracle."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import _pkg  # noqa: E402

_pkg.load()
import oracle  # noqa: E402
from the_algorithm_b200.ann.brute_force import BruteForceIndex  # noqa: E402
from the_algorithm_b200.ann.common import ComposedQueryable, Cosine, FuturePool, InnerProduct, L2  # noqa: E402

rng = np.random.default_rng(0)
ok = True
gemm = "--no-gemm" not in sys.argv
for metric in (InnerProduct, Cosine, L2):
    n, d = 9000, 40
    corpus = (rng.standard_normal((n, d)) / 6).astype(np.float32)
    corpus[17] = 0.0
    ids = rng.permutation(n).astype(np.int64)
    q = rng.uniform(-1, 1, (300, d)).astype(np.float32)
    ix = BruteForceIndex.apply(metric, FuturePool.immediate_pool())
    ix.append_batch(ids[:5000], corpus[:5000])
    ix.append_batch(ids[5000:], corpus[5000:])
    if metric is not Cosine:   # with a special row present (zero row under Cosine) an update parks the tensor-core path
        ix.update_batch(ids[:3], corpus[:3])
    want = oracle.query_canonical(metric.ordinal, corpus, ids, q, 20)
    for path, cg, nq in ((1, 0, 9), (2, 1, 300), (2, 2, 300), (2, 2, 5), (3, 0, 2)):
        if path == 2 and not gemm:
            continue
        ix.set_option("path", path)
        if cg:
            ix.set_option("gemm_cta_group", cg)
        got = ix.batch_query_with_distance(q[:nq], 20)
        good = bool((got[0] == want[0][:nq]).all() and (got[1].view(np.uint32) == want[1][:nq].view(np.uint32)).all())
        print(metric.name, "path", path, "cg", cg, "b", nq, "ok", good, flush=True)
        ok &= good
    ix.set_option("path", 0)
    shards = [BruteForceIndex.apply(metric, FuturePool.immediate_pool()) for _ in range(2)]
    shards[0].append_batch(ids[:4000], corpus[:4000])
    shards[1].append_batch(ids[4000:], corpus[4000:])
    got = ComposedQueryable(shards).batch_query_with_distance(q[:7], 20)
    ok &= bool((got[0] == want[0][:7]).all())
    for s in shards:
        s.close()
    ix.close()
print("SANITIZE_TARGET_OK", ok)
sys.exit(0 if ok else 1)
