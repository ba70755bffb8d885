"""GPU tests of the exact multi-GPU mode (linked k-slabs, include/sdfb.h): the 16 sweeps keep the reference's serial
order across slab faces (cpu_lib/makelevelset3.cpp:143-149, :245-248) by handing boundary planes over column by column
inside the sweep kernel.  `slabs` needs ONE GPU (the slabs' sweep kernels are co-resident on it), the other two >= 2.
Bar: bit equality with one plan on the whole grid and with the compiled reference."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _run(mode, timeout=900):
    r = subprocess.run([sys.executable, os.path.join(HERE, "linked_gpu_worker.py"), mode], capture_output=True, text=True, timeout=timeout)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("LINKED_RESULT ")][-1]
    return json.loads(line[len("LINKED_RESULT "):])


def test_linked_slabs_on_one_gpu_equal_one_plan_and_the_reference():
    out = _run("slabs")
    print(out)
    assert len(out) == 4 and all(c["ok"] for c in out), out


def test_linked_slabs_one_per_device():
    out = _run("devices")
    if isinstance(out, dict) and "skipped" in out:
        pytest.skip(out["skipped"])
    print(out)
    assert all(c["ok"] for c in out), out


def test_multi_gpu_one_shot_call_equals_one_gpu():
    out = _run("oneshot")
    if isinstance(out, dict) and "skipped" in out:
        pytest.skip(out["skipped"])
    print(out)
    assert all(c["ok"] for c in out), out
