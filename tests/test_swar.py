"""The SWAR byte predicates of the kernels (grab_b200/csrc/swar.h are __host__ __device__): brute-force checked on the CPU."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))


def test_swar_primitives(tmp_path):
    exe = str(tmp_path / "swar_check")
    subprocess.run(["g++", "-O2", "-std=c++17", os.path.join(HERE, "swar_check.cc"), "-o", exe], check=True)
    p = subprocess.run([exe], stdout=subprocess.PIPE, timeout=300)
    assert p.returncode == 0, p.stdout.decode()
    assert b"swar ok" in p.stdout
