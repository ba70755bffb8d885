"""CPU test of the fused first-pass launch's scheduling logic (sdfgen_b200/csrc/sdfb_sweep_columns.cu:
k_sweep_columns_fused / wait_previous_sweep): oracle/experiments/sweep_overlap.c holds a literal port of the device's
prerequisite arithmetic and of its double-buffered progress words and runs the fused ticket order with W concurrent
"CTAs" and the most eager column choice.  Asserted: no deadlock, the buffer-reuse invariant (sweep q is complete when a
column of sweep q+2 runs), bit equality with the serial sweeps (cpu_lib/makelevelset3.cpp:104-151) -- on whole grids and
k-slabs with sizes that are no multiples of the column extent, and for launches that start at an odd sweep."""
import importlib.util
import os

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("fused_emulation", os.path.join(HERE, "..", "oracle", "experiments", "fused_emulation.py"))
fe = importlib.util.module_from_spec(spec)
spec.loader.exec_module(fe)


@pytest.fixture(scope="module")
def lib():
    return fe.build()


@pytest.mark.parametrize("case", fe.CASES, ids=lambda c: f"{c[0]}-{c[1]}-{c[2]}x{c[3]}-s{c[4]}+{c[5]}-W{c[6]}")
def test_fused_launch_order_is_exact(lib, case):
    dims, slab, EJ, EK, first, count, W = case
    rc, early = fe.run_case(lib, dims, slab, EJ, EK, first, count, W)
    declined = slab in ((63, 64), (0, 1))          # a one-plane slab on a grid face: half of the sweeps update nothing there
    assert rc == (-3 if declined else 0), rc
    if not declined and count >= 3:
        assert early > 0                            # consecutive sweeps really overlapped
