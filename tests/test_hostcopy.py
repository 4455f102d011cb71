"""The threaded staging copy behind vkhel_vector_copy_from_host
(vkhel_b200/csrc/hostcopy.cu; the reference's memcpy into its mapped buffer,
src/vector.c:262-268): exact for every size and offset, from several application
threads at once, with 1, 4 (default) and 8 threads per copy.  Host only."""
import os
import subprocess

import pytest

from conftest import ROOT

OBJ = os.path.join(ROOT, "build", "obj", "hostcopy.o")
BIN = os.path.join(ROOT, "build", "bin", "hostcopy_check")


@pytest.fixture(scope="module")
def checker():
    if not os.path.exists(OBJ):
        subprocess.check_call(["make", "-C", ROOT, "-s", "build/obj/hostcopy.o"])
    os.makedirs(os.path.dirname(BIN), exist_ok=True)
    subprocess.check_call(
        ["/usr/bin/g++", "-O1", "-std=c++17",
         os.path.join(ROOT, "tests", "native", "hostcopy_check.cpp"), OBJ,
         "-o", BIN, "-L/usr/local/cuda/lib64", "-lcudart_static", "-lpthread",
         "-ldl", "-lrt"])
    return BIN


@pytest.mark.parametrize("threads_per_copy", ["1", "4", "8"])
@pytest.mark.parametrize("app_threads", [1, 4])
def test_host_copy_exact(checker, threads_per_copy, app_threads):
    env = dict(os.environ, VKHEL_COPY_THREADS=threads_per_copy)
    res = subprocess.run([checker, str(app_threads), "120"], env=env,
                         capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    assert res.stdout.strip() == "ok"


def test_host_copy_under_thread_sanitizer(tmp_path):
    """the pool's protocol (generation|index word, claims by compare-and-swap)
    has no data race that ThreadSanitizer can see; skipped where the
    sanitizer runtime is not installed"""
    obj = str(tmp_path / "hostcopy_tsan.o")
    exe = str(tmp_path / "hostcopy_tsan")
    inc = ["-I" + os.path.join(ROOT, d)
           for d in ("include", "include/vkhel", "vkhel_b200/csrc")]
    try:
        subprocess.check_call(
            ["/usr/local/cuda/bin/nvcc", "-O1", "-g", "-std=c++17",
             "-Wno-deprecated-gpu-targets", "-Xcompiler",
             "-fPIC,-fsanitize=thread"] + inc +
            ["-c", os.path.join(ROOT, "vkhel_b200", "csrc", "hostcopy.cu"),
             "-o", obj], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        subprocess.check_call(
            ["/usr/bin/g++", "-O1", "-g", "-fsanitize=thread",
             os.path.join(ROOT, "tests", "native", "hostcopy_check.cpp"), obj,
             "-o", exe, "-L/usr/local/cuda/lib64", "-lcudart_static",
             "-lpthread", "-ldl", "-lrt"],
            stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    except (OSError, subprocess.CalledProcessError):
        pytest.skip("no ThreadSanitizer toolchain")
    env = dict(os.environ, VKHEL_COPY_THREADS="4")
    res = subprocess.run([exe, "3", "40"], env=env, capture_output=True,
                         text=True, timeout=600)
    if "FATAL: ThreadSanitizer" in res.stderr:
        pytest.skip("ThreadSanitizer cannot run here: " + res.stderr[:120])
    assert "ThreadSanitizer: data race" not in res.stderr, res.stderr[-3000:]
    assert res.returncode == 0 and res.stdout.strip().endswith("ok")
