"""The device code replaces two routines of the portable math / small-matrix headers by cheaper forms that must give the
same bits: the FMA short cut of ll_sincosf for small angles and the register-only 3x3 column-pivoted QR solve.  Both are
host-compilable; these tests build the C/C++ checkers under tests/csrc and run them over millions of cases."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))


def _run(src, compiler, flags, tmp_path):
    exe = str(tmp_path / os.path.splitext(os.path.basename(src))[0])
    subprocess.check_call([compiler] + flags + ["-o", exe, os.path.join(HERE, "csrc", src), "-lm"])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    return [int(v) for v in r.stdout.split()]


def test_sincos_short_cut_is_bit_identical(tmp_path):
    accepted, rejected, mismatches = _run("check_sincos_shortcut.c", "gcc", ["-O2", "-ffp-contract=off"], tmp_path)
    assert mismatches == 0 and accepted > 10_000_000
    assert rejected < accepted // 1000  # the certificate almost always holds, so the short cut is the common path


def test_qr3_specialisation_is_bit_identical(tmp_path):
    cases, mismatches = _run("check_qr3.cpp", "g++", ["-std=c++11", "-O3", "-ffp-contract=off"], tmp_path)
    assert cases >= 3_000_000 and mismatches == 0
