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


def test_hash_tables(tmp_path):
    """The hashed engine's tables (pattern.cc build_hash) against the kernel's lookup restated on the host."""
    root = os.path.dirname(HERE)
    exe = str(tmp_path / "hash_check")
    subprocess.run(["g++", "-O2", "-std=c++17", os.path.join(HERE, "hash_check.cc"), os.path.join(root, "grab_b200", "csrc", "pattern.cc"),
                    "-I", os.path.join(root, "include"), "-o", exe], check=True)
    p = subprocess.run([exe], stdout=subprocess.PIPE, timeout=600)
    assert p.returncode == 0, p.stdout.decode()
    assert b"hash ok" in p.stdout
