"""CPU: the emulated CUDA sources under AddressSanitizer -- the no-GPU stand-in for
`compute-sanitizer --tool memcheck` (which profiles/ runs on the B200 box): every kernel of the
clip pipeline and of the operator level on exactly-sized buffers, in a child process that
preloads the ASan runtime."""
import os
import subprocess
import sys

import pytest

emu = pytest.importorskip("emu", reason="tests/emu")


def test_kernels_stay_inside_their_buffers():
    from emu import build
    asan = build.libasan()
    if asan is None:
        pytest.skip("no libasan for this gcc")
    build.build(asan=True)
    env = dict(os.environ, LD_PRELOAD=asan, ASAN_OPTIONS="detect_leaks=0:detect_stack_use_after_return=0")
    drive = os.path.join(os.path.dirname(os.path.abspath(build.__file__)), "asan_drive.py")
    proc = subprocess.run([sys.executable, drive], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                          text=True, timeout=900)
    assert proc.returncode == 0 and "ASAN DRIVE DONE" in proc.stdout, proc.stdout[-4000:]
    assert "ERROR: AddressSanitizer" not in proc.stdout
    # the sink case must have gone through the whole-tile fallback (bins) / the whole-batch fallback (direct index:
    # more excess pairs than the list holds)
    sink = {l.split()[1]: eval(l.split(None, 2)[2]) for l in proc.stdout.splitlines() if l.startswith("sink ")}
    assert sink["bins"]["full"] == 1
    assert sink["ldg"]["full"] == 0 and sink["ldg"]["excess"] > sink["ldg"]["excess_cap"]
