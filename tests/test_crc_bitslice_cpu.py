"""The bit-sliced 32-frame CRC of csrc/crc_bitslice.cuh (the arithmetic of aos_scan_kernel / imtr_validate_kernel) is
plain C++17: build it for the host and check it against the bit-serial definition (ref CRC.h:806-834, :1519)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))


def test_bitsliced_crc_matches_bit_serial_definition(tmp_path):
    exe = str(tmp_path / "crc_bs_test")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-o", exe, os.path.join(HERE, "native", "crc_bitslice_host_test.cpp")])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "mismatches: 0" in out.stdout
