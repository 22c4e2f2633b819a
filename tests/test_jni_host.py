"""The JNI boundary (the-algorithm_b200/host/jni/b200ann_jni.c <-> host/scala/*.scala) without a JDK.

The image has no JVM and no jni.h, so nothing would otherwise compile the shim's body.  Here it is compiled against a
stand-in header (tests/jni_stub/jni.h: the JNI types and function-table members the shim uses) and every
Java_com_twitter_ann_brute_1force_B200AnnNative_00024_* entry point is EXECUTED against a mock JNIEnv
(tests/jni_host_check.c): on a CPU box `create` must fail loudly, on a B200 the known answers of the C++ host check are
reproduced through the JNI symbols, the sharded handle and knnJoin agree with the single index and the reference's
directory format round-trips.  A static check keeps the Scala `@native` declarations and the C functions in step.
"""
import re
import shutil
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
SHIM = ROOT / "the-algorithm_b200" / "host" / "jni" / "b200ann_jni.c"
SCALA = ROOT / "the-algorithm_b200" / "host" / "scala"


def _build(tmp_path, built_lib):
    gcc = shutil.which("gcc")
    if not gcc:
        pytest.skip("no gcc")
    exe = tmp_path / "jni_host_check"
    cmd = [gcc, "-std=gnu11", "-O1", "-Wall", "-Wextra", "-Werror", f"-I{ROOT / 'tests' / 'jni_stub'}", "-o", str(exe),
           str(ROOT / "tests" / "jni_host_check.c"), f"-L{built_lib.parent}", "-lb200ann", f"-Wl,-rpath,{built_lib.parent}", "-lm"]
    subprocess.run(cmd, check=True)
    return exe


def test_jni_shim_body_compiles_and_fails_loudly_without_gpu(tmp_path, built_lib):
    exe = _build(tmp_path, built_lib)
    r = subprocess.run([str(exe), str(tmp_path)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "no device" in r.stdout or "gpu ok" in r.stdout


@pytest.mark.gpu
def test_jni_entry_points_known_answers_on_gpu(tmp_path, built_lib):
    exe = _build(tmp_path, built_lib)
    r = subprocess.run([str(exe), str(tmp_path)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "gpu ok" in r.stdout, r.stdout + r.stderr
    # the directories the shim wrote are the reference's layout (BruteForceIndex.scala:142-161, ShardedSerialization.scala:28-38)
    assert (tmp_path / "jni_single" / "BruteForceFileData").stat().st_size > 0
    assert (tmp_path / "jni_single" / "_SUCCESS").exists()
    assert (tmp_path / "jni_sharded" / "shard_0" / "BruteForceFileData").exists()
    assert (tmp_path / "jni_sharded" / "shard_1" / "BruteForceFileData").exists()


_SCALA_TO_JNI = {"Int": "jint", "Long": "jlong", "ByteBuffer": "jobject", "String": "jstring", "Array[Int]": "jintArray",
                 "Unit": "void"}


def _scala_natives():
    text = (SCALA / "GpuBruteForceIndex.scala").read_text()
    out = {}
    for m in re.finditer(r"@native def (\w+)\((.*?)\): ([\w\[\]]+)", text, re.S):
        params = [p.split(":")[1].strip() for p in m.group(2).split(",") if p.strip()]
        out[m.group(1)] = ([_SCALA_TO_JNI[p] for p in params], _SCALA_TO_JNI[m.group(3)])
    return out


def _jni_functions():
    text = SHIM.read_text()
    out = {}
    for m in re.finditer(r"NATIVE\((\w+), (\w+)\)\(JNIEnv \*env, jobject self(.*?)\)\s*\{", text, re.S):
        params = [p.strip().rsplit(" ", 1)[0].strip() for p in m.group(3).split(",") if p.strip()]
        out[m.group(2)] = (params, m.group(1))
    return out


def test_scala_natives_and_jni_functions_agree_in_name_arity_and_types():
    scala, jni = _scala_natives(), _jni_functions()
    assert len(scala) >= 25 and set(scala) == set(jni), sorted(set(scala) ^ set(jni))
    for name, sig in scala.items():
        assert jni[name] == sig, (name, sig, jni[name])


def test_every_native_the_scala_classes_call_is_declared():
    declared = set(_scala_natives())
    for f in SCALA.glob("*.scala"):
        for used in re.findall(r"B200AnnNative\.(\w+)\(", f.read_text()):
            assert used in declared, (f.name, used)


def test_stub_header_says_what_it_is():
    text = (ROOT / "tests" / "jni_stub" / "jni.h").read_text()
    assert "TEST STAND-IN, not the JDK header" in text


def test_scala_sources_are_bracket_balanced():
    """No scalac here: at least the brackets of the shipped Scala sources balance (comments and string literals stripped)."""
    for f in SCALA.glob("*.scala"):
        t = re.sub(r"//.*", "", f.read_text())
        t = re.sub(r"/\*.*?\*/", "", t, flags=re.S)
        t = re.sub(r's?"(\\.|[^"\\])*"', '""', t)
        for o, c in ("()", "[]", "{}"):
            assert t.count(o) == t.count(c), (f.name, o, t.count(o), t.count(c))
