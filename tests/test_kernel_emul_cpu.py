"""Memory safety and warp convergence of the hot kernels without a GPU (K2, its prefetch / register-ring / hub-window forms, K2H, K2R): csrc/cb_spmm_kernel.cuh and cb_spmm_hub_kernel.cuh are
compiled UNMODIFIED for the host against a lock-step warp emulator (tests/emul/cuda_emul.h: 32 lanes = 32 threads, *_sync
intrinsics are rendezvous points; CTAs, clusters, shared memory and reads of a neighbour CTA's shared memory for the hub
variant) and run under AddressSanitizer + UBSan on small tiles with hub rows, empty rows, ragged widths, column slabs and the
accumulate mode, against a scalar loop over the same functors.  The hub variant (K2H) and the ring-pipelined variant (K2R,
whose cp.async copies the emulator DEFERS until the matching wait_group, so a slot read too early shows stale bytes) must
also reproduce plain K2 bit for bit.  (compute-sanitizer is closed on the GPU pool.)"""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_k2_and_fixup_on_the_warp_emulator_under_asan(tmp_path):
    exe = str(tmp_path / "kernel_emul")
    cmd = ["/usr/bin/g++", "-std=c++17", "-O1", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined",
           "-fno-omit-frame-pointer", "-Wno-int-in-bool-context", "-I/usr/local/cuda/include", f"-I{ROOT}/include",
           f"-I{ROOT}/combblas-spmm-test_b200/csrc", "-o", exe, f"{ROOT}/tests/emul/kernel_emul.cpp", "-lpthread"]
    subprocess.check_call(cmd, timeout=600)
    r = subprocess.run([exe, "1"], capture_output=True, text=True, timeout=600)      # a dead-locked warp = diverged *_sync = timeout
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "emulation ok" in r.stdout and "mismatches=0" in r.stdout and "runtime error" not in r.stderr
    assert r.stdout.count("mismatches=0") == 68 and "differs from K2" not in r.stdout     # 12 K2 + 6 narrow + 7 K2 with entry prefetch + 14 K2P + 4 K2P with L2 hints + 4 K2W + 4 persistent-warp K2 + 8 K2H + 9 K2R cases (argument 2 = a second seed)


def test_device_number_parser_against_strtod(tmp_path):
    """csrc/cb_mmparse.cuh (the decimal -> double conversion of the device Matrix Market reader) compiled as plain C++ under UBSan:
    every string it accepts must give strtod's bits - random doubles in the formats matrix files use, 19-digit mantissas over
    the whole exponent range, exact round-to-even ties."""
    exe = str(tmp_path / "mmparse_host")
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O2", "-fsanitize=undefined", "-fno-sanitize-recover=undefined", "-o", exe,
                           f"{ROOT}/tests/emul/mmparse_host.cpp"], timeout=600)
    r = subprocess.run([exe, "400000"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "mismatches 0 " in r.stdout and "runtime error" not in r.stderr, r.stdout[-2000:] + r.stderr[-2000:]
