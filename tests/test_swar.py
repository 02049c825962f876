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


def test_pattern_compiler_fuzz(tmp_path):
    """Metacharacter soup through the pattern compiler and the oracle's parser under ASan/UBSan: no crash, no hang, every
    rejection has a message, MINLENGTH and nullability agree wherever both accept (tests/fuzz_compile.cc)."""
    root = os.path.dirname(HERE)
    exe = str(tmp_path / "fuzz_compile")
    obj = str(tmp_path / "oracle.o")
    subprocess.run(["gcc", "-O2", "-c", os.path.join(root, "oracle", "grab_oracle.c"), "-o", obj], check=True)
    subprocess.run(["g++", "-O2", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-std=c++17",
                    os.path.join(HERE, "fuzz_compile.cc"), os.path.join(root, "grab_b200", "csrc", "pattern.cc"), obj,
                    "-I", os.path.join(root, "include"), "-o", exe], check=True)
    p = subprocess.run([exe, "300000"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=600)
    assert p.returncode == 0 and b"fuzz ok" in p.stdout, p.stdout.decode()[-3000:]
