"""GPU parity: the ten dispatched quantizers vs the reference C functions (full_loop.c:29-516) over
the input classes of the reference's test/QuantAsmTest.cc + quantize_func_test.cc (zero, DC only,
large negative, random, all log_scales, with/without quantization matrices)."""
import numpy as np
import pytest

import quant_helpers as qh
from helpers import rng

pytestmark = pytest.mark.gpu


def test_quantizers_all_variants(b200, oracle):
    r = rng(31)
    n = 0
    for v, c, t, sc, qm, iqm, ls in qh.cases(r):
        if oracle.ref is not None:
            want = qh.call_ref(oracle.ref, v[1], c, t, sc, qh.ref_extra(v, qm, iqm, ls))
        else:
            want = qh.call_port(oracle.port, v[2], c, t, sc, qm, iqm, ls)
        got = b200.quantize(v[0], c, t, sc, qm, iqm, ls)
        assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1]) and got[2] == want[2], (
            v[0], c.size, ls, qm is not None)
        n += 1
    assert n > 1000
