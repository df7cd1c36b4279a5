#!/usr/bin/env python
"""Counts, for the bench circuit's tape, the warp-level memory instructions of the interpreter and the distinct 128-byte
lines they touch per instance (operand loads, result stores incl. the cooperative bit-run stores) - the quantity
that bounds tape_exec_kernel.  CPU only.  Usage: python scripts/tape_lines.py"""
import sys, numpy as np, ctypes, collections, time
sys.path.insert(0,'.')
from circom_b200.circuit import CircuitDesc
from circom_b200 import circuits as C
from circom_b200.witness_calculator import Circuit
from circom_b200 import native
d=CircuitDesc("bn128"); d.set_main(C.ecdsa_scale(d,8,132))
c=Circuit(d, host_only=True)
st=c.stats
n_ops=st['n_tape_ops']; nl=st['n_levels']
ops=np.zeros((n_ops,4),dtype=np.uint32); ls=np.zeros(nl+1,dtype=np.uint32)
native.lib.cw_circuit_tape(c._h, ops.ctypes.data_as(ctypes.c_void_p), ls.ctypes.data_as(ctypes.c_void_p), None)
opc=ops[:,0]&0xFF; dst=ops[:,0]>>8
def lines_for(order_in_level=None, threads=64):
    tot_load_instr=0; tot_lines=0; tot_st_lines=0; tot_st_instr=0
    for l in range(nl):
        idx=np.arange(ls[l],ls[l+1])
        if order_in_level is not None: idx=order_in_level(idx)
        for w0 in range(0,len(idx),32):
            sel=idx[w0:w0+32]
            for col in (1,2):
                x=ops[sel,col]
                isslot=((x&0x80000000)==0)
                # BITS col2 unused (const); asserts have operands
                s=(x[isslot]&0xFFFFFF)
                if len(s)==0: continue
                tot_load_instr+=1
                tot_lines+=len(np.unique(s>>2))   # 4 slots of 32B per 128B line
            # stores (non-run, non-assert)
            o=opc[sel]; hasdst=~np.isin(o,[26,27,30,33]); run=(o==29)&((ops[sel,3]>>24)>0)
            s=dst[sel][hasdst&~run]
            if len(s): tot_st_instr+=1; tot_st_lines+=len(np.unique(s>>2))
            for r in sel[run]:
                n=(int(ops[r,3])>>24)+1; d0=int(dst[r])
                tot_st_instr+=1; tot_st_lines+=len(set(range(d0>>2,((d0+n-1)>>2)+1)))
    return tot_load_instr,tot_lines,tot_st_instr,tot_st_lines
t0=time.time()
print("current: load instr %d lines %d | store instr %d lines %d"%lines_for(), time.time()-t0)
def by_a(idx):
    a=ops[idx,1]; key=np.where(a&0x80000000, 0xFFFFFFFF, a&0xFFFFFF)
    o=opc[idx]
    return idx[np.lexsort((key,o))]
print("sorted by (opcode,a) [dst numbering unchanged]: load instr %d lines %d | store instr %d lines %d"%lines_for(by_a))
