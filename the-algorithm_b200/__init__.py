"""the-algorithm_b200 -- B200-native exact nearest-neighbour engine for ONE path of sagspot/the-algorithm:
the dense scan behind com.twitter.ann.brute_force.BruteForceIndex (ann/src/main/scala/com/twitter/ann/...).

  csrc/        CUDA kernels (sm_100a) + the C ABI implementation (include/b200ann.h)
  lib/         built libb200ann.so (git-ignored)
  _capi.py     ctypes binding of the C ABI
  ann/         host-side mirror of the reference interface (Metric, Queryable, Appendable, BruteForceIndex,
               QueryableById, ShardedAppendable / ComposedQueryable) -- same names and argument meaning
  host/        the same mirror for the reference's own language: Scala class + JNI shim (cannot be compiled in
               this image: no JVM), and a C++ header

Import through the repo-root helper because the directory name has a hyphen:  import _pkg; _pkg.load()
"""
__all__ = ["ann"]
