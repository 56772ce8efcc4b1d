"""The HQ packer's bit writer (four finished words per 16-byte store, one bit cursor; csrc/bitwriter.cuh) compiled as
plain C++ and fuzzed on the host against a bit-by-bit model of the reference's bounded MSB-first stream
(VLC.cpp:119-185) with the call sequence of hq_pack_kernel (Slices.cpp:478-530).  No GPU needed."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_wide_bit_writer_against_bit_model(tmp_path):
    exe = str(tmp_path / "bitwriter_fuzz")
    subprocess.run(["g++", "-O2", "-std=c++14", "-Wall", os.path.join(ROOT, "tests", "bitwriter_fuzz.cpp"), "-o", exe],
                   check=True, timeout=120)
    for seed in ("1", "2026"):
        r = subprocess.run([exe, "20000", seed], capture_output=True, text=True, timeout=120)
        assert r.returncode == 0, r.stdout + r.stderr
