"""Tiny target for `compute-sanitizer --tool memcheck` (run under gpurun, wrapped in `timeout`): every kernel of the library on
small shapes -- append / update, prep, streaming scan, tensor-core filter (cta_group 1 and 2, seed + chunks + resident-query
mode), compaction, finalize, exact fallback, merge, the two-phase shard query (seed publish + seed merge, 70k-row shards), the
fused exchange/merge kernel and the plain-pair metric kernels.  Exits non-zero if a result differs from the CPU oracle."""
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
# ---- two-phase shard query (K5c), fused exchange/merge (K5b) and the metric kernels, on shapes with ragged tails ----
import torch  # noqa: E402

from the_algorithm_b200.ann.common import MetricUtil  # noqa: E402
from the_algorithm_b200.ann.exchange import ResultBlock, exchange_merge_blocks, result_block_bytes, slice_of  # noqa: E402

dev = torch.device("cuda", 0)
st = torch.cuda.current_stream().cuda_stream
n_per, d, b, k = 70_003, 24, 37, 33
corpus = (rng.standard_normal((3 * n_per, d)) / 5).astype(np.float32)
ids = rng.permutation(3 * n_per).astype(np.int64)
q = rng.uniform(-1, 1, (b, d)).astype(np.float32)
q_t = torch.from_numpy(q).to(dev)
for metric in (InnerProduct, Cosine, L2):
    shards = []
    for s_ in range(3):
        ix = BruteForceIndex.apply(metric, FuturePool.immediate_pool())
        ix.append_batch(ids[s_ * n_per:(s_ + 1) * n_per], corpus[s_ * n_per:(s_ + 1) * n_per])
        shards.append(ix)
    keys = torch.empty((3, b, k), dtype=torch.int32, device=dev)
    nb = result_block_bytes(b, k)
    off = (nb + 255) // 256 * 256
    bufs = [torch.zeros(2 * off, dtype=torch.uint8, device=dev) for _ in range(3)]
    locs = [ResultBlock(bufs[s_], b, k, 0) for s_ in range(3)]
    fins = [ResultBlock(bufs[s_], b, k, off) for s_ in range(3)]
    for s_, ix in enumerate(shards):
        ix.query_seed_device(q_t, k, keys[s_], st)
    for s_, ix in enumerate(shards):
        ix.query_finish_device(q_t, k, [keys[j].data_ptr() for j in range(3)], *locs[s_].tensors, st)
    for r in range(3):
        q0, q1 = slice_of(r, 3, b)
        exchange_merge_blocks([x.ptr for x in locs], [x.ptr for x in fins], b, k, q0, q1 - q0, 0, st)
    torch.cuda.synchronize()
    want = oracle.query_canonical(metric.ordinal, corpus, ids, q, k)
    good = all(bool((f.ids.cpu().numpy() == want[0]).all() and (f.dist.cpu().numpy().view(np.uint32) == want[1].view(np.uint32)).all())
               for f in fins)
    print(metric.name, "two-phase shards + fused exchange ok", good, flush=True)
    ok &= good
    a_, b_ = corpus[:1001], corpus[1001:2002]
    pd = metric.distances(a_, b_)
    wd = np.array([oracle.distance(metric.ordinal, a_[i], b_[i]) for i in range(len(a_))], np.float32)
    good = bool((pd.view(np.uint32) == wd.view(np.uint32)).all())
    print(metric.name, "distance pairs ok", good, flush=True)
    ok &= good
    for ix in shards:
        ix.close()
good = bool((MetricUtil.norm(corpus[:1001]).view(np.uint32) == oracle.normalize(corpus[:1001]).view(np.uint32)).all())
print("normalize ok", good, flush=True)
ok &= good
print("SANITIZE_TARGET_OK", ok)
sys.exit(0 if ok else 1)
