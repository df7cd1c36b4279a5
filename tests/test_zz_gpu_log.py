"""cw_batch_log on the device (last in the order of the suite: the only test of this entry point)."""
import random

import numpy as np
import pytest

from circom_b200.circuit import CircuitDesc
from circom_b200 import circuits as C
from circom_b200.witness_calculator import Circuit, Batch
from oracle import ir_eval
from tests.util import flat_inputs

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("o0", [False, True])
def test_batch_log_prints_what_the_reference_prints(o0):
    """log() calls (LogBucket, log_bucket.rs:104-162): per instance the text of the reference calculator - here the
    evaluator's, which tests/test_oracle_c.py pins against the calculator's stdout - incl. a logged signal that the signal
    elimination removed from the witness"""
    d = CircuitDesc("bn128")
    d.set_main(C.logging(d))
    rng = random.Random(5)
    ins = [{"a": 3, "b": 5}, {"a": d.q - 1, "b": d.q - 2}] + [{"a": rng.randrange(d.q), "b": rng.randrange(d.q)} for _ in range(31)]
    c = Circuit(d, o0=o0)
    b = Batch(c, len(ins))
    b.set_inputs(flat_inputs(d, ins))
    b.run()
    assert not b.status().any()
    for i in (0, 1, 7, 32):
        ir_eval.LOG_SINK.clear()
        ir_eval.evaluate(d, ins[i])
        text = "".join(ir_eval.LOG_SINK)
        assert b.log(i) == text and text.count("\n") == 4
        assert c.format_log(b.witness()[i]) == text
