"""CPU test of the exact multi-GPU protocol (sdfgen_b200/csrc/sdfb_sweep_columns.cu, LINK = true): every slab runs its
fused 16-sweep ticket sequence with a few "CTA" slots, per-sweep inbound planes and per-column flags, interleaved by a
seeded random scheduler (oracle/experiments/sweep_overlap.c :: linked_emulation_run).  Asserted: no deadlock, no read of
a cell that was not handed over in this sweep (buffers start poisoned), bit equality with the serial sweeps of the whole
grid (cpu_lib/makelevelset3.cpp:104-151 in the order of :245-248)."""
import importlib.util
import os

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("linked_emulation", os.path.join(HERE, "..", "oracle", "experiments", "linked_emulation.py"))
le = importlib.util.module_from_spec(spec)
spec.loader.exec_module(le)


@pytest.fixture(scope="module")
def lib():
    return le.build()


@pytest.mark.parametrize("case", le.CASES, ids=lambda c: f"{c[0]}-{len(c[1]) - 1}slabs-{c[2]}x{c[3]}-W{c[5]}")
def test_linked_slabs_protocol_is_exact_and_deadlock_free(lib, case):
    dims, bounds, EJ, EK, count, W, seed = case
    rc, overlap = le.run_case(lib, dims, bounds, EJ, EK, count, W, seed)
    assert rc == 0, rc
    assert overlap > 0          # the slabs really were in different sweeps at the same time


def test_other_interleavings_and_ticket_orders(lib):
    for seed in range(20, 26):
        for w in (1, 2, 3, 64):                   # anti-diagonals, J-weighted keys, row by row
            rc, _ = le.run_case(lib, (18, 27, 48), (0, 12, 24, 36, 48), 8, 4, 16, 2, seed, order_w=w)
            assert rc == 0, (seed, w, rc)
