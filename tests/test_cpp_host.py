"""The C++ host mirror (the-algorithm_b200/host/cpp/b200ann.hpp) compiles, links against libb200ann.so and behaves:
on a CPU box the constructor fails loudly with ANN_ERR_NO_DEVICE; on a B200 a 3-row known answer is checked."""
import shutil
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


def _build(tmp_path, built_lib):
    gxx = shutil.which("g++")
    if not gxx:
        pytest.skip("no g++")
    exe = tmp_path / "cpp_host_check"
    cmd = [gxx, "-std=c++17", "-O1", "-o", str(exe), str(ROOT / "tests" / "cpp_host_check.cpp"), f"-L{built_lib.parent}",
           "-lb200ann", f"-Wl,-rpath,{built_lib.parent}", "-I/usr/local/cuda/include", "-L/usr/local/cuda/lib64", "-lcudart", "-pthread"]
    subprocess.run(cmd, check=True)
    return exe


def test_cpp_host_mirror_links_and_fails_loudly_without_gpu(tmp_path, built_lib):
    exe = _build(tmp_path, built_lib)
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "no device" in r.stdout or "gpu ok" in r.stdout


@pytest.mark.gpu
def test_cpp_host_mirror_known_answer_on_gpu(tmp_path, built_lib):
    exe = _build(tmp_path, built_lib)
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "gpu ok" in r.stdout, r.stdout + r.stderr


def test_jni_shim_is_guarded_and_compiles_without_jdk(tmp_path):
    gcc = shutil.which("gcc")
    if not gcc:
        pytest.skip("no gcc")
    subprocess.run([gcc, "-c", "-o", str(tmp_path / "jni.o"), str(ROOT / "the-algorithm_b200" / "host" / "jni" / "b200ann_jni.c")],
                   check=True)


def test_scala_binding_names_every_abi_entry_point_it_uses():
    scala = (ROOT / "the-algorithm_b200" / "host" / "scala" / "GpuBruteForceIndex.scala").read_text()
    jni = (ROOT / "the-algorithm_b200" / "host" / "jni" / "b200ann_jni.c").read_text()
    for native, abi in (("create", "ann_create"), ("destroy", "ann_destroy"), ("appendBatch", "ann_append_batch"),
                        ("size", "ann_size"), ("queryBatch", "ann_query_batch"), ("lastError", "ann_last_error"),
                        ("knnJoin", "ann_knn_join"), ("distancePairs", "ann_distance_pairs"),
                        ("normalizeRows", "ann_normalize_rows"), ("querySeedDevice", "ann_query_seed_device"),
                        ("queryFinishDevice", "ann_query_finish_device"), ("shardedCreate", "ann_sharded_create"),
                        ("shardedDestroy", "ann_sharded_destroy"), ("shardedAppendBatch", "ann_sharded_append_batch"),
                        ("shardedSize", "ann_sharded_size"), ("shardedQueryBatch", "ann_sharded_query_batch"),
                        ("saveDirectory", "ann_save_directory"), ("loadDirectory", "ann_load_directory"),
                        ("shardedSaveDirectory", "ann_sharded_save_directory"),
                        ("shardedLoadDirectory", "ann_sharded_load_directory")):
        assert f"def {native}(" in scala and abi in jni
    for trait in ("extends Appendable[T, BruteForceRuntimeParams.type, D]", "with Queryable[T, BruteForceRuntimeParams.type, D]"):
        assert trait in scala
    sharded = (ROOT / "the-algorithm_b200" / "host" / "scala" / "GpuShardedBruteForceIndex.scala").read_text()
    for used in ("shardedCreate", "shardedAppendBatch", "shardedQueryBatch", "shardedSaveDirectory", "shardedLoadDirectory", "shardedDestroy"):
        assert f"B200AnnNative.{used}(" in sharded
    for trait in ("extends Appendable[Long, BruteForceRuntimeParams.type, D]", "with Queryable[Long, BruteForceRuntimeParams.type, D]"):
        assert trait in sharded
