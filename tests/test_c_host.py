"""tests/c/abi_smoke.c: a plain C program (no Python, no ctypes) linked against libola_gpu.so -- the host a Rust / C
maintainer writes.  Compiled here with gcc; without a GPU it must fail loudly at ola_gpu_init, with one it commits, proves,
verifies and drives a prove session."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path):
    import olavm_b200

    olavm_b200.load()
    exe = str(tmp_path / "abi_smoke")
    libdir = os.path.join(ROOT, "olavm_b200")
    subprocess.check_call(["gcc", "-O2", "-Wall", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "c", "abi_smoke.c"), "-o", exe,
                           "-L", libdir, "-lola_gpu", "-Wl,-rpath," + libdir])
    return exe


def test_c_host_links_and_fails_loudly_without_a_gpu(tmp_path):
    exe = _build(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "FAIL " not in r.stdout
    assert "PASSED" in r.stdout.splitlines()[-1]


@pytest.mark.gpu
def test_c_host_commits_proves_verifies_on_the_gpu(tmp_path):
    exe = _build(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=900)
    failed = [ln for ln in r.stdout.splitlines() if ln.startswith("FAIL")]
    assert not failed and r.returncode == 0, "\n".join(failed) + r.stderr[-500:]
    assert r.stdout.splitlines()[-1] == "PASSED", r.stdout[-800:]
    assert "ola_prove (Cmp + RangeCheck" in r.stdout and "session proof" in r.stdout
