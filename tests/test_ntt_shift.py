"""The shift-twiddle register rounds of the forward NTT network (olavm_b200/csrc/ntt_shift.cuh, w96.cuh).

A radix-2^K block of the Cooley-Tukey network is computed as "scale row m by theta^m, then a 2^K-point transform whose
twiddles are powers of two" in a 96-bit lazy representation.  tools/microbench/bfly16.cu holds the check against the
reference semantics -- K stages of butterflies  (a, b) -> (a + w b, a - w b)  with the twiddles the tile kernels use
(cfft/serial.rs butterflies; twiddles as in ntt.cu: c_u * omega_{2^(u+1)}^{bitrev_u(q)}), canonical values compared:
  * CPU: the same header compiled by g++ (portable arithmetic, with the range assumptions of the device code asserted);
  * GPU: the PTX carry-chain versions inside a kernel, 2000 chained rounds, against the radix-2 form the round-1 kernels use.
The full transforms are compared with the oracle in tests/test_gpu_parity.py (both forms: OLA_NTT_SHIFT=0 / 1)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tools", "microbench", "bfly16.cu")
INC = os.path.join(ROOT, "olavm_b200", "csrc")

P = 2**64 - 2**32 + 1


def test_small_roots_of_unity_are_the_powers_of_two_the_shift_form_uses():
    # goldilocks_field.rs:77 POWER_OF_TWO_GENERATOR; omega_{2^k} = g^(2^(32-k))
    g = 1753635133440165772
    w64 = pow(g, 2 ** (32 - 6), P)
    assert w64 == pow(2, 39, P)
    for s in range(6):  # omega_{2^(s+1)} = 2^(39 * 2^(5-s))
        assert pow(g, 2 ** (32 - (s + 1)), P) == pow(2, (39 << (5 - s)) % 192, P)
    assert pow(2, 96, P) == P - 1 and pow(2, 64, P) == 2**32 - 1  # t^3 = -1, t^2 = t - 1


def test_shift_form_equals_reference_butterflies_on_the_host(tmp_path):
    exe = str(tmp_path / "bfly16_host")
    subprocess.check_call(["g++", "-O2", "-x", "c++", "-std=c++17", "-I", INC, "-o", exe, SRC])
    r = subprocess.run([exe, "host"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count(": ok") == 7 and "FAILED" not in r.stdout, r.stdout


@pytest.mark.gpu
def test_shift_form_equals_radix2_form_on_the_gpu(tmp_path):
    exe = str(tmp_path / "bfly16")
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-I", INC, "-o", exe, SRC])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and r.stdout.splitlines()[-1] == "PASSED", r.stdout[-2000:] + r.stderr[-500:]
    assert r.stdout.count("equality of canonical outputs") >= 8 and "differ): FAILED" not in r.stdout


@pytest.mark.gpu
def test_radix2_form_still_matches_the_oracle():
    """OLA_NTT_SHIFT=0 keeps the round-1 butterflies (the A/B baseline of profiles/r02m_*): the transform and
    commitment parity tests run once more in a process with that form selected."""
    env = dict(os.environ, OLA_NTT_SHIFT="0")
    r = subprocess.run(["python", "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_parity.py"), "-m", "gpu", "-q", "-x", "-k",
                        "ntt or lde or commit"], capture_output=True, text=True, timeout=1200, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-500:]
