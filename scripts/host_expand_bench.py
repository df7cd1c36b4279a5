#!/usr/bin/env python
"""Host-only probe of the witness-row expansion (development): GB/s of 32-byte rows written, per pass, for the
pool configuration given by CW_UNPACK_THREADS / CW_UNPACK_PIN / CW_EXPAND_ISA."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from circom_b200.circuit import CircuitDesc
from circom_b200 import circuits as C
from circom_b200.witness_calculator import Circuit
from circom_b200.native import lib, check
n = int(sys.argv[1]) if len(sys.argv) > 1 else 160
d = CircuitDesc("bn128")
d.set_main(C.ecdsa_scale(d, 8, 132), "e")
c = Circuit(d, host_only=True)
for mode, name in ((0, "expand"), (1, "stream-fill"), (2, "memset")):
    g = np.zeros(4)
    check(lib.cw_host_expand_bench(c._h, n, 4, mode, g.ctypes.data))
    print("%-12s %s  GB/s per pass: %s" % (name, lib.cw_host_pool_info().decode(), " ".join("%.1f" % x for x in g)), flush=True)
