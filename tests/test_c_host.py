"""tests/c/abi_smoke.c: a plain C program (no Python, no ctypes) linked against libola_gpu.so -- the host a Rust / C
maintainer writes.  Compiled here with gcc; without a GPU it must fail loudly at ola_gpu_init, with one it commits, proves,
verifies and drives a prove session."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path, name="abi_smoke"):
    import olavm_b200

    olavm_b200.load()
    exe = str(tmp_path / name)
    libdir = os.path.join(ROOT, "olavm_b200")
    subprocess.check_call(["gcc", "-O2", "-Wall", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "c", name + ".c"), "-o", exe,
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


def _trace_file(tmp_path, orc):
    import numpy as np

    from workload import trace_json as wj

    path = str(tmp_path / "trace.json")
    with open(path, "w") as f:
        f.write(wj.records_to_json(wj.system_records(orc, np.random.default_rng(4), n_iter=3), orc))
    return path


def test_ola_prove_file_parses_then_fails_loudly_without_a_gpu(tmp_path, orc):
    """tests/c/ola_prove_file.c = `ola prove -i trace.json -o proof.bin`: the parse is host code and succeeds anywhere; without a
    CUDA device the program stops at ola_gpu_init (exit 3), it never proves on the CPU."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present: covered by the gpu test")
    exe = _build(tmp_path, "ola_prove_file")
    r = subprocess.run([exe, _trace_file(tmp_path, orc), str(tmp_path / "proof.bin")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 3, r.stdout + r.stderr
    assert "table  4: 2^16 rows x 12 columns" in r.stdout and "ola_gpu_init failed" in r.stderr
    bad = tmp_path / "bad.json"
    bad.write_text('{"exec":[{"clk":"seven"}]}')
    r = subprocess.run([exe, str(bad), str(tmp_path / "proof.bin")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 2 and "not a Trace" in r.stderr


@pytest.mark.gpu
def test_ola_prove_file_proves_a_trace_file_and_the_proof_file_verifies(tmp_path, orc):
    import olavm_b200

    exe = _build(tmp_path, "ola_prove_file")
    out = tmp_path / "proof.bin"
    r = subprocess.run([exe, _trace_file(tmp_path, orc), str(out)], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-500:] + r.stderr[-500:]
    assert "Prove done!" in r.stdout and "Verify succeed!" in r.stdout
    proof = out.read_bytes()
    ok, why = olavm_b200.verify_proof(list(range(12)), proof)
    assert ok, why
    assert orc.stark_verify(list(range(12)), proof)[0]
