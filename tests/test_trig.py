"""sincos_ref (one range reduction and one polynomial of each kind for both results; the kernels use it) must equal the
separate ports of compiler-rt's sinf / cosf (sin_ref / cos_ref, pinned by the wasm golden vectors) bit for bit.
Default: every 251st float bit pattern plus the neighbourhoods of all range boundaries (17 M arguments, ~1 s);
R2D_TRIG_EXHAUSTIVE=1: all 2^32 (~10 s on 8 cores; 0 mismatches when this was written)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))


def test_sincos_ref_equals_sin_ref_and_cos_ref():
    exe = os.path.join(HERE, "trig", "_build", "trig_check")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-pthread", "-o", exe,
                           os.path.join(HERE, "trig", "trig_check.cpp")])
    stride = "1" if os.environ.get("R2D_TRIG_EXHAUSTIVE") == "1" else "251"
    out = subprocess.run([exe, stride], capture_output=True, text=True)
    assert out.returncode == 0 and "mismatches: 0" in out.stdout, out.stdout[-2000:]
