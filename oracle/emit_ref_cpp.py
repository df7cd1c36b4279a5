"""Hand-lowering of a circuit description into the C++ the reference compiler would print
(TEST INFRASTRUCTURE), so that the REAL reference runtime — c_elements/common/{main,calcwit}.cpp +
the rendered generic/fr.cpp, built by oracle/build_ref.py — can execute the same program and
produce the `.wtns` our GPU path must match byte for byte.

The Rust compiler cannot run here, so this module prints, per template instance, what
compiler/src/circuit_design/template.rs:177-472 prints (`<T>_create` / `<T>_run`), with the bodies
the bucket emitters would give for straight-line code:

  Compute   Fr_<op>(&expaux[k], a, b)                          compute_bucket.rs:410-421
  Load      &signalValues[mySignalStart + i] | &circuitConstants[c] |
            &ctx->signalValues[ctx->componentMemory[mySubcomponents[s]].signalStart + i]
                                                               load_bucket.rs:325-447, value_bucket.rs:81-87
  Store     Fr_copy(dest, src); for a sub-component input: inputCounter -= 1 and run at zero
                                                               store_bucket.rs:607-734
  CreateCmp <Sub>_create(mySignalStart + off, ctx_index + coff + 1, ctx, name, myId)
                                                               create_component_bucket.rs:204-352
  `===`     Fr_eq + assert(Fr_isTrue)                          assert_bucket.rs:70-88
  c ? a : b if (Fr_isTrue(c)) .. else ..                       branch_bucket.rs:100-122
and the file prologue / epilogue of circuit.rs:420-563, plus the matching `.dat`
(c_code_generator.rs:575-679,818-865: input hash map, witness2signal list, constants in
Montgomery form with the short/long tag).
"""
from __future__ import annotations

import os
import struct
from typing import List

K_NONE, K_OWN, K_SUB, K_CONST, K_TMP, K_ONE = 0, 1, 2, 3, 4, 5

_FN = {1: "mul", 2: "div", 3: "add", 4: "sub", 5: "pow", 6: "idiv", 7: "mod", 8: "shl", 9: "shr",
       10: "leq", 11: "geq", 12: "lt", 13: "gt", 14: "eq", 15: "neq", 16: "lor", 17: "land", 19: "bor",
       20: "band", 21: "bxor"}
_FN1 = {18: "lnot", 22: "bnot", 23: "neg"}


def fnv1a(s: str) -> int:
    h = 0xCBF29CE484222325
    for ch in s.encode():
        h ^= ch
        h = (h * 0x100000001B3) & 0xFFFFFFFFFFFFFFFF
    return h


def _header(t) -> str:
    return "%s_%d" % (t.name, t.id)


def emit_cpp(desc) -> str:
    if desc.prime == "goldilocks":
        return emit_cpp_goldilocks(desc)
    out: List[str] = []
    w = out.append
    main = desc.main
    for inc in ("<stdio.h>", "<iostream>", "<assert.h>", '"circom.hpp"', '"calcwit.hpp"', '"fr.hpp"'):
        w("#include %s" % inc)
    for t in desc.templates:
        w("void %s_create(uint soffset,uint coffset,Circom_CalcWit* ctx,std::string componentName,uint componentFather);" % _header(t))
        w("void %s_run(uint ctx_index,Circom_CalcWit* ctx);" % _header(t))
    for f in getattr(desc, "functions", []):
        w("void %s_%d(Circom_CalcWit* ctx,FrElement* lvar,uint componentFather,FrElement* destination,int destination_size);"
          % (f.name, f.id))
    w("Circom_TemplateFunction _functionTable[%d] = { %s };" % (
        len(desc.templates), ",".join("%s_run" % _header(t) for t in desc.templates)))
    w("Circom_TemplateFunction _functionTableParallel[%d] = { %s };" % (
        len(desc.templates), ",".join("NULL" for _ in desc.templates)))
    w("uint get_main_input_signal_start() {return %d;}\n" % (main.n_out + 1))
    w("uint get_main_input_signal_no() {return %d;}\n" % main.n_in)
    w("uint get_total_signal_no() {return %d;}\n" % desc.total_signals)
    w("uint get_number_of_components() {return %d;}\n" % main.total_components)
    w("uint get_size_of_input_hashmap() {return %d;}\n" % hashmap_size(desc))
    w("uint get_size_of_witness() {return %d;}\n" % desc.total_signals)
    w("uint get_size_of_constants() {return %d;}\n" % len(desc.consts))
    w("uint get_size_of_io_map() {return %d;}\n" % len(getattr(desc, "io_map_templates", ())))
    w("uint get_size_of_bus_field_map() {return 0;}\n")
    w("void release_memory_component(Circom_CalcWit* ctx, uint pos) {{\n"
      "if (pos != 0){{\n"
      "if(ctx->componentMemory[pos].subcomponents)delete []ctx->componentMemory[pos].subcomponents;\n"
      "if(ctx->componentMemory[pos].subcomponentsParallel)delete []ctx->componentMemory[pos].subcomponentsParallel;\n"
      "if(ctx->componentMemory[pos].outputIsSet)delete []ctx->componentMemory[pos].outputIsSet;\n"
      "if(ctx->componentMemory[pos].mutexes)delete []ctx->componentMemory[pos].mutexes;\n"
      "if(ctx->componentMemory[pos].cvs)delete []ctx->componentMemory[pos].cvs;\n"
      "if(ctx->componentMemory[pos].sbct)delete []ctx->componentMemory[pos].sbct;\n"
      "}}\n}}\n")
    w("// function declarations")
    for f in getattr(desc, "functions", []):
        _emit_function(w, f)
    w("// template declarations")
    for t in desc.templates:
        _emit_template(w, t)
    w("void run(Circom_CalcWit* ctx){")
    w('%s_create(1,0,ctx,"main",0);' % _header(main))
    if main.n_in > 0:
        w("%s_run(0,ctx);" % _header(main))
    w("}\n")
    return "\n".join(out)


def _emit_function(w, f) -> None:
    """function.rs:91-126 prints `void f_k(Circom_CalcWit* ctx,FrElement* lvar,uint componentFather,
    FrElement* destination,int destination_size)`; variables are lvar[] slots, loops are
    `while(Fr_isTrue(..))` (loop_bucket.rs:76-91) - printed here with labels and gotos, which the same
    Fr_isTrue / Fr_toInt / Fr_* calls drive."""
    w("void %s_%d(Circom_CalcWit* ctx,FrElement* lvar,uint componentFather,FrElement* destination,int destination_size){"
      % (f.name, f.id))
    w("FrElement* circuitConstants = ctx->circuitConstants;")

    def addr(r):
        if r[0] == K_TMP:
            return "&lvar[%d]" % r[2]
        if r[0] == K_CONST:
            return "&circuitConstants[%d]" % r[2]
        raise ValueError(r)
    for pc, (op, d, a, b, c) in enumerate(f.code):
        w("L%d: ;" % pc)
        if op == 40:
            w("goto L%d;" % a[2])
        elif op == 41:
            w("if (!Fr_isTrue(%s)) goto L%d;" % (addr(a), b[2]))
        elif op == 42:
            if b[0] == 0 and b[2] > 1:     # array return (return_bucket.rs:70-120: Fr_copyn of destination_size elements)
                w("Fr_copyn(destination,&lvar[%d],destination_size); return;" % a[2])
            else:
                w("Fr_copy(destination,%s); return;" % addr(a))
        elif op == 43:
            w("Fr_copy(&lvar[%d],&lvar[%d + Fr_toInt(%s)]);" % (d[2], a[2], addr(b)))
        elif op == 44:
            w("Fr_copy(&lvar[%d + Fr_toInt(%s)],%s);" % (a[2], addr(b), addr(c)))
        elif op == 45:     # a nested call (call_bucket.rs:466-533): the callee's lvar array is filled with the arguments
            fn = f.desc.functions[a[2]]
            w("{")
            w("FrElement lvarcall[%d];" % fn.n_regs)
            for k in range(fn.n_params):
                w("Fr_copy(&lvarcall[%d],&lvar[%d]);" % (k, b[2] + k))
            w("%s_%d(ctx,lvarcall,componentFather,%s,%d);" % (fn.name, fn.id, addr(d), c[2] if c[0] == 0 and c[2] > 1 else 1))
            w("}")
        elif op == 24:
            w("Fr_copy(%s,%s);" % (addr(d), addr(a)))
        elif op in _FN:
            w("Fr_%s(%s,%s,%s);" % (_FN[op], addr(d), addr(a), addr(b)))
        elif op in _FN1:
            w("Fr_%s(%s,%s);" % (_FN1[op], addr(d), addr(a)))
        else:
            raise ValueError("cannot emit function op %d" % op)
    w("L%d: ;" % len(f.code))
    w("}\n")


def _emit_template(w, t) -> None:
    H = _header(t)
    w("void %s_create(uint soffset,uint coffset,Circom_CalcWit* ctx,std::string componentName,uint componentFather){" % H)
    w("ctx->componentMemory[coffset].templateId = %d;" % t.id)
    w('ctx->componentMemory[coffset].templateName = "%s";' % t.name)
    w("ctx->componentMemory[coffset].signalStart = soffset;")
    w("ctx->componentMemory[coffset].inputCounter = %d;" % t.n_in)
    w("ctx->componentMemory[coffset].componentName = componentName;")
    w("ctx->componentMemory[coffset].idFather = componentFather;")
    if t.subs:
        w("ctx->componentMemory[coffset].subcomponents = new uint[%d]{0};" % len(t.subs))
    else:
        w("ctx->componentMemory[coffset].subcomponents = new uint[0];")
    if t.n_in == 0:
        w("%s_run(coffset,ctx);" % H)
    w("}\n")

    w("void %s_run(uint ctx_index,Circom_CalcWit* ctx){" % H)
    w("FrElement* circuitConstants = ctx->circuitConstants;")
    w("FrElement* signalValues = ctx->signalValues;")
    # temporaries live in expaux (one slot per SSA temporary + one for assert comparisons)
    w("static thread_local FrElement expaux_store[%d];" % (t.n_tmp + 1) if t.n_tmp > 4000 else
      "FrElement expaux_store[%d];" % (t.n_tmp + 1))
    w("FrElement* expaux = expaux_store;")
    w("FrElement lvar[1];")
    w("u64 mySignalStart = ctx->componentMemory[ctx_index].signalStart;")
    w("std::string myTemplateName = ctx->componentMemory[ctx_index].templateName;")
    w("std::string myComponentName = ctx->componentMemory[ctx_index].componentName;")
    w("u64 myFather = ctx->componentMemory[ctx_index].idFather;")
    w("u64 myId = ctx_index;")
    w("u32* mySubcomponents = ctx->componentMemory[ctx_index].subcomponents;")
    w("bool* mySubcomponentsParallel = ctx->componentMemory[ctx_index].subcomponentsParallel;")
    w("uint sub_component_aux;")
    w("uint index_multiple_eq;")
    w("int cmp_index_ref_load = -1;")
    # CreateCmp buckets
    soff, coff = t.n_own, 0
    for i, s in enumerate(t.subs):
        w("{")
        w('std::string new_cmp_name = "%s";' % s.name.replace('"', ""))
        w("%s_create(mySignalStart+%d,%d+ctx_index+1,ctx,new_cmp_name,myId);" % (_header(s.tmpl), soff, coff))
        w("mySubcomponents[%d] = %d+ctx_index+1;" % (i, coff))
        w("}")
        soff += s.tmpl.total_signals
        coff += s.tmpl.total_components

    def addr(r):
        k = r[0]
        if k == K_OWN:
            return "&signalValues[mySignalStart + %d]" % r[2]
        if k == K_SUB:
            st = t.subs[r[1]].tmpl
            if st.id in getattr(t.desc, "io_map_templates", ()):
                # LocationRule::Mapped (load_bucket.rs:264-322, store_bucket.rs:498-566): the element of a component array of
                # mixed templates - the signal's offset comes from templateInsId2IOSignalInfo at run time
                code, idx, is_array = _io_code(st, r[2])
                d0 = "ctx->templateInsId2IOSignalInfo[ctx->componentMemory[mySubcomponents[%d]].templateId].defs[%d]" % (r[1], code)
                acc = "%s.offset" % d0
                if is_array:
                    acc = "%s+(%d)*%s.size" % (acc, idx, d0)
                return "&ctx->signalValues[ctx->componentMemory[mySubcomponents[%d]].signalStart + %s]" % (r[1], acc)
            return "&ctx->signalValues[ctx->componentMemory[mySubcomponents[%d]].signalStart + %d]" % (r[1], r[2])
        if k == K_CONST:
            return "&circuitConstants[%d]" % r[2]
        if k == K_TMP:
            return "&expaux[%d]" % r[2]
        if k == K_ONE:
            return "&signalValues[0]"
        raise ValueError(r)

    cmp_slot = "&expaux[%d]" % t.n_tmp
    pending_args = []
    for op, d, a, b, c in t.ops:
        if op == 46:  # ARG
            pending_args.append(a)
            continue
        if op == 45:  # CALL (call_bucket.rs:466-533): arguments are copied into the callee's lvar array
            fn = t.desc.functions[a[2]]
            n = b[2]
            args = pending_args[len(pending_args) - n:]
            del pending_args[len(pending_args) - n:]
            w("{")
            w("FrElement lvarcall[%d];" % fn.n_regs)
            for k, r in enumerate(args):
                w("Fr_copy(&lvarcall[%d],%s);" % (k, addr(r)))
            w("%s_%d(ctx,lvarcall,myId,%s,%d);" % (fn.name, fn.id, addr(d), c[2] if c[0] == 0 and c[2] > 1 else 1))
            w("}")
            continue
        if op == 27:  # ASSERT_EQ
            w("{")
            w("Fr_eq(%s,%s,%s);" % (cmp_slot, addr(a), addr(b)))
            w('if (!Fr_isTrue(%s)) std::cout << "Failed assert in template/function " << myTemplateName << std::endl;' % cmp_slot)
            w("assert(Fr_isTrue(%s));" % cmp_slot)
            w("}")
            continue
        if op == 26:  # ASSERT
            w("assert(Fr_isTrue(%s));" % addr(a))
            continue
        if op == 29:  # LOG: LogBucket::produce_c (log_bucket.rs:104-162)
            if a[0] == 0:
                w('{ printf("%s"); }' % t.desc.strings[b[2]])
            else:
                w("{ char* temp = Fr_element2str(%s); printf(\"%%s\",temp); delete [] temp; }" % addr(a))
            w('{ printf("%s"); }' % ("\\n" if c[2] else " "))
            continue
        if op == 48:  # LOADSIG: a load whose Indexed location is computed at run time (load_bucket.rs:325-447 with
            #           ComputeBucket ToAddress = Fr_toInt, compute_bucket.rs:361-363)
            w("{")
            w("Fr_copy(%s,&signalValues[mySignalStart + (%d + Fr_toInt(%s))]);" % (addr(d), a[2], addr(b)))
            w("}")
            continue
        dst = addr(d)
        w("{")
        if op == 24:
            w("Fr_copy(%s,%s);" % (dst, addr(a)))
        elif op == 25:
            w("if (Fr_isTrue(%s)) { Fr_copy(%s,%s); } else { Fr_copy(%s,%s); }" % (addr(c), dst, addr(a), dst, addr(b)))
        elif op in _FN:
            w("Fr_%s(%s,%s,%s);" % (_FN[op], dst, addr(a), addr(b)))
        elif op in _FN1:
            w("Fr_%s(%s,%s);" % (_FN1[op], dst, addr(a)))
        else:
            raise ValueError("cannot emit op %d" % op)
        if d[0] == K_SUB:
            st = t.subs[d[1]].tmpl
            if st.n_out <= d[2] < st.n_out + st.n_in:
                w("if(!(ctx->componentMemory[mySubcomponents[%d]].inputCounter -= 1)){" % d[1])
                w("%s_run(mySubcomponents[%d],ctx);" % (_header(st), d[1]))
                w("}")
        w("}")
    w("for (uint i = 0; i < %d; i++){" % len(t.subs))
    w("uint index_subc = ctx->componentMemory[ctx_index].subcomponents[i];")
    w("if (index_subc != 0){")
    w("assert(!(ctx->componentMemory[index_subc].inputCounter));")
    w("release_memory_component(ctx,index_subc);")
    w("}")
    w("}")
    w("}\n")


def _io_code(st, local_id: int):
    """(signal code, element index, is array) of input / output signal `local_id` of template instance st: codes number the
    declared outputs then inputs (TemplateDB::get_signal_id order = IODef order, build.rs:498-520)"""
    code, off = 0, 0
    for cat in ("out", "in"):
        for name, n in st.sigs[cat]:
            if off <= local_id < off + n:
                return code, local_id - off, st.sig_is_array[name]
            off += n
            code += 1
    raise ValueError("signal %d of %s is not an input or output" % (local_id, st.name))


def io_map_bytes(desc) -> bytes:
    """generate_dat_io_signals_info (c_code_generator.rs:681-735): the template ids, then per template the number of signals
    and per signal offset, #dimensions - 1, the dimensions but the first, element size, bus id - u32 little endian (the Rust
    writes to_be_bytes reversed).  generate_dat_bus_field_info (:737-794) adds nothing without buses."""
    from circom_b200.circuit import CircuitDesc
    tids = sorted(getattr(desc, "io_map_templates", ()))
    out = b"".join(struct.pack("<I", t) for t in tids)
    for t in tids:
        defs = CircuitDesc.io_defs(desc.templates[t])
        out += struct.pack("<I", len(defs))
        for off, lengths, size, bus in defs:
            out += struct.pack("<II", off, len(lengths) - 1 if lengths else 0)
            out += b"".join(struct.pack("<I", x) for x in lengths[1:])
            out += struct.pack("<II", size, bus)
    return out


def hashmap_size(desc) -> int:
    n = len(desc.main_inputs())
    s = 256
    while s < n:
        s <<= 1
    return s


def dat_bytes(desc) -> bytes:
    size = hashmap_size(desc)
    table = [(0, 0, 0)] * size
    for name, gid, n in desc.main_inputs():
        h = fnv1a(name)
        p = h % size
        while table[p][1] != 0:
            p = (p + 1) % size
        table[p] = (h, gid, n)
    out = b"".join(struct.pack("<QQQ", *e) for e in table)
    out += b"".join(struct.pack("<Q", i) for i in range(desc.total_signals))
    q = desc.q
    R = 1 << (((q.bit_length() + 63) // 64) * 64)
    for v in (() if desc.prime == "goldilocks" else desc.consts):   # (goldilocks: literals in the code, c_code_generator.rs:838-841)
        n = v % q
        nn = n - q if n > q // 2 else n
        if -2147483648 <= nn <= 2147483647:
            out += struct.pack("<iI", nn, 0x40000000)
        else:
            out += struct.pack("<iI", 0, 0xC0000000)
        out += ((n * R) % q).to_bytes(32, "little")
    return out + io_map_bytes(desc)


def emit_cpp_goldilocks(desc) -> str:
    """The goldilocks flavour of the generated code (every `prime_str != "goldilocks"` branch of the emitters taken the other
    way): values are plain `u64`, operators are expressions - `dest = Fr_mul(a, b);` (compute_bucket.rs:425-460,
    store_bucket.rs:641-645) -, constants are literals (`<n>ull`, value_bucket.rs:82-86), there is no constant table, the
    runtime is c_elements/common64/{main,calcwit}.cpp with c_elements/goldilocks/fr.hpp.  Component machinery as in the
    256-bit flavour (template.rs:177-472)."""
    if getattr(desc, "functions", []):
        raise ValueError("the goldilocks emitter covers templates only")
    out: List[str] = []
    w = out.append
    main = desc.main
    for inc in ("<stdio.h>", "<iostream>", "<assert.h>", '"circom.hpp"', '"calcwit.hpp"', '"fr.hpp"'):
        w("#include %s" % inc)
    for t in desc.templates:
        w("void %s_create(uint soffset,uint coffset,Circom_CalcWit* ctx,std::string componentName,uint componentFather);" % _header(t))
        w("void %s_run(uint ctx_index,Circom_CalcWit* ctx);" % _header(t))
    w("Circom_TemplateFunction _functionTable[%d] = { %s };" % (
        len(desc.templates), ",".join("%s_run" % _header(t) for t in desc.templates)))
    w("Circom_TemplateFunction _functionTableParallel[%d] = { %s };" % (
        len(desc.templates), ",".join("NULL" for _ in desc.templates)))
    w("uint get_main_input_signal_start() {return %d;}\n" % (main.n_out + 1))
    w("uint get_main_input_signal_no() {return %d;}\n" % main.n_in)
    w("uint get_total_signal_no() {return %d;}\n" % desc.total_signals)
    w("uint get_number_of_components() {return %d;}\n" % main.total_components)
    w("uint get_size_of_input_hashmap() {return %d;}\n" % hashmap_size(desc))
    w("uint get_size_of_witness() {return %d;}\n" % desc.total_signals)
    w("uint get_size_of_io_map() {return %d;}\n" % len(getattr(desc, "io_map_templates", ())))
    w("uint get_size_of_bus_field_map() {return 0;}\n")
    w("void release_memory_component(Circom_CalcWit* ctx, uint pos) {{\n"
      "if (pos != 0){{\n"
      "if(ctx->componentMemory[pos].subcomponents)delete []ctx->componentMemory[pos].subcomponents;\n"
      "}}\n}}\n")
    for t in desc.templates:
        H = _header(t)
        w("void %s_create(uint soffset,uint coffset,Circom_CalcWit* ctx,std::string componentName,uint componentFather){" % H)
        w("ctx->componentMemory[coffset].templateId = %d;" % t.id)
        w('ctx->componentMemory[coffset].templateName = "%s";' % t.name)
        w("ctx->componentMemory[coffset].signalStart = soffset;")
        w("ctx->componentMemory[coffset].inputCounter = %d;" % t.n_in)
        w("ctx->componentMemory[coffset].componentName = componentName;")
        w("ctx->componentMemory[coffset].idFather = componentFather;")
        w("ctx->componentMemory[coffset].subcomponents = new uint[%d]{0};" % max(len(t.subs), 1))
        if t.n_in == 0:
            w("%s_run(coffset,ctx);" % H)
        w("}\n")
        w("void %s_run(uint ctx_index,Circom_CalcWit* ctx){" % H)
        w("u64* signalValues = ctx->signalValues;")
        w("u64 expaux[%d];" % (t.n_tmp + 1))
        w("u64 lvar[1];")
        w("u64 mySignalStart = ctx->componentMemory[ctx_index].signalStart;")
        w("std::string myTemplateName = ctx->componentMemory[ctx_index].templateName;")
        w("u64 myId = ctx_index;")
        w("u32* mySubcomponents = ctx->componentMemory[ctx_index].subcomponents;")
        soff, coff = t.n_own, 0
        for i, sb in enumerate(t.subs):
            w("{")
            w('std::string new_cmp_name = "%s";' % sb.name.replace('"', ""))
            w("%s_create(mySignalStart+%d,%d+ctx_index+1,ctx,new_cmp_name,myId);" % (_header(sb.tmpl), soff, coff))
            w("mySubcomponents[%d] = %d+ctx_index+1;" % (i, coff))
            w("}")
            soff += sb.tmpl.total_signals
            coff += sb.tmpl.total_components

        def val(r, t=t):
            k = r[0]
            if k == K_OWN:
                return "signalValues[mySignalStart + %d]" % r[2]
            if k == K_SUB:
                st = t.subs[r[1]].tmpl
                if st.id in getattr(t.desc, "io_map_templates", ()):
                    code, idx, is_array = _io_code(st, r[2])
                    d0 = "ctx->templateInsId2IOSignalInfo[ctx->componentMemory[mySubcomponents[%d]].templateId].defs[%d]" % (r[1], code)
                    acc = "%s.offset" % d0
                    if is_array:
                        acc = "%s+(%d)*%s.size" % (acc, idx, d0)
                    return "ctx->signalValues[ctx->componentMemory[mySubcomponents[%d]].signalStart + %s]" % (r[1], acc)
                return "ctx->signalValues[ctx->componentMemory[mySubcomponents[%d]].signalStart + %d]" % (r[1], r[2])
            if k == K_CONST:
                return "%dull" % t.desc.consts[r[2]]
            if k == K_TMP:
                return "expaux[%d]" % r[2]
            if k == K_ONE:
                return "signalValues[0]"
            raise ValueError(r)
        for op, d, a, b, c in t.ops:
            if op == 27:
                w('if (!Fr_isTrue(Fr_eq(%s,%s))) std::cout << "Failed assert in template/function " << myTemplateName << std::endl;' % (val(a), val(b)))
                w("assert(Fr_isTrue(Fr_eq(%s,%s)));" % (val(a), val(b)))
                continue
            if op == 26:
                w("assert(Fr_isTrue(%s));" % val(a))
                continue
            if op == 48:
                w("%s = signalValues[mySignalStart + (%d + Fr_toInt(%s))];" % (val(d), a[2], val(b)))
                continue
            w("{")
            if op == 24:
                w("%s = %s;" % (val(d), val(a)))
            elif op == 25:
                w("if (Fr_isTrue(%s)) { %s = %s; } else { %s = %s; }" % (val(c), val(d), val(a), val(d), val(b)))
            elif op in _FN:
                w("%s = Fr_%s(%s,%s);" % (val(d), _FN[op], val(a), val(b)))
            elif op in _FN1:
                w("%s = Fr_%s(%s);" % (val(d), _FN1[op], val(a)))
            else:
                raise ValueError("cannot emit op %d" % op)
            if d[0] == K_SUB:
                st = t.subs[d[1]].tmpl
                if st.n_out <= d[2] < st.n_out + st.n_in:
                    w("if(!(ctx->componentMemory[mySubcomponents[%d]].inputCounter -= 1)){" % d[1])
                    w("%s_run(mySubcomponents[%d],ctx);" % (_header(st), d[1]))
                    w("}")
            w("}")
        w("for (uint i = 0; i < %d; i++){" % len(t.subs))
        w("uint index_subc = ctx->componentMemory[ctx_index].subcomponents[i];")
        w("if (index_subc != 0){")
        w("assert(!(ctx->componentMemory[index_subc].inputCounter));")
        w("release_memory_component(ctx,index_subc);")
        w("}")
        w("}")
        w("}\n")
    w("void run(Circom_CalcWit* ctx){")
    w('%s_create(1,0,ctx,"main",0);' % _header(main))
    if main.n_in > 0:
        w("%s_run(0,ctx);" % _header(main))
    w("}\n")
    return "\n".join(out)


def build_reference_calculator(desc, out_dir: str, name: str | None = None, opt: str = "-O3") -> str:
    """Writes <name>.cpp / <name>.dat and links the reference runtime; returns the binary path."""
    from . import build_ref
    name = name or desc.name
    os.makedirs(out_dir, exist_ok=True)
    cpp = os.path.join(out_dir, name + ".cpp")
    binp = os.path.join(out_dir, name)
    with open(cpp, "w") as f:
        f.write(emit_cpp(desc))
    with open(binp + ".dat", "wb") as f:
        f.write(dat_bytes(desc))
    build_ref.build_calculator(desc.prime, cpp, binp, opt=opt)
    return binp
