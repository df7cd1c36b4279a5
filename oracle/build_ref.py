"""Build the REAL reference runtime into oracle/_ref/ (TEST INFRASTRUCTURE ONLY).

Nothing here is product code.  The product (circom_b200/) never imports or
links anything from oracle/; only tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py do.

What is built (all outputs go to oracle/_ref/, which is git-ignored but is
shipped to the GPU box together with our own .so files):

  oracle/_ref/<prime>/fr.cpp, fr.hpp   the reference's portable (`--no_asm`) field
        library, rendered from its handlebars template
        /root/reference/code_producers/src/c_elements/generic/fr.{cpp,hpp}
        with the template variables the reference compiler computes at
        code_producers/src/c_elements/c_code_generator.rs:1086-1129.
  oracle/_ref/libfr_<prime>.so         that file compiled as a shared library
        (links the system libgmp.so.10 through the prototype shim in
        oracle/gmp_shim/gmp.h, because the image has no gmp.h).
  oracle/_ref/rt_<prime>_{main,calcwit}.o   the reference's runtime shell
        c_elements/common/{main,calcwit}.cpp compiled where they lie
        (never copied), for linking with a hand-lowered <circuit>.cpp
        (see oracle/emit_ref_cpp.py) into a `<bin> input.json out.wtns`
        calculator identical in shape to what `circom --c --no_asm` produces.

The asm field library (<prime>/fr.asm) is NOT buildable here (no nasm); the
reference states both variants behave identically (generic/fr.cpp:430).

/root/reference does not exist on the GPU box: this script is a no-op there
and the prebuilt files are used.
"""
from __future__ import annotations

import os
import re
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/code_producers/src/c_elements"
OUT = os.path.join(HERE, "_ref")
GMP_SO = "/usr/lib/x86_64-linux-gnu/libgmp.so.10"
JSON_INC_CANDIDATES = [
    "/opt/prime-rl/.venv/lib/python3.12/site-packages/include/cudnn_frontend/thirdparty",
    "/usr/include",
]

# program_structure/src/utils/constants.rs:3-6
PRIMES = {
    "bn128": 21888242871839275222246405745257275088548364400416034343698204186575808495617,
    "bls12381": 52435875175126190479447740508185965837690552500527637822603658699938581184513,
    "grumpkin": 21888242871839275222246405745257275088696311157297823662689037894645226208583,
    "pallas": 28948022309329048855892746252171976963363056481941560715954676764349967630337,
    "vesta": 28948022309329048855892746252171976963363056481941647379679742748393362948097,
    "secq256r1": 115792089210356248762697446949407573530086143415290314195533631308867097853951,
    "bls12377": 8444461749428370424248824938781546531375899335154063827935233455917409239041,
}


# --------------------------------------------------------------------------
# a small handlebars subset: {{x}}, {{helper arg}}, {{#if}}/{{else}}/{{/if}},
# {{#each}} with @index/@last/this, nested helper calls in parentheses.
# --------------------------------------------------------------------------
_TAG = re.compile(r"\{\{(.*?)\}\}", re.S)


def _parse(tokens, pos, stop):
    nodes = []
    while pos < len(tokens):
        kind, val = tokens[pos]
        if kind == "text":
            nodes.append(("text", val))
            pos += 1
            continue
        tag = val.strip()
        if tag in stop or (tag.startswith("/") and tag in stop):
            return nodes, pos
        if tag.startswith("#if"):
            cond = tag[3:].strip()
            then, pos = _parse(tokens, pos + 1, {"else", "/if"})
            els = []
            if tokens[pos][1].strip() == "else":
                els, pos = _parse(tokens, pos + 1, {"/if"})
            nodes.append(("if", cond, then, els))
            pos += 1
        elif tag.startswith("#each"):
            name = tag[5:].strip()
            body, pos = _parse(tokens, pos + 1, {"/each"})
            nodes.append(("each", name, body))
            pos += 1
        else:
            nodes.append(("expr", tag))
            pos += 1
    return nodes, pos


def _tokenise_expr(s):
    return re.findall(r"\(|\)|[^\s()]+", s)


def _eval_expr(s, scope):
    toks = _tokenise_expr(s)

    def atom(i):
        t = toks[i]
        if t == "(":
            v, i = call(i + 1)
            assert toks[i] == ")"
            return v, i + 1
        if t.isdigit():
            return int(t), i + 1
        for sc in reversed(scope):
            if t in sc:
                return sc[t], i + 1
        raise KeyError(t)

    def call(i):
        t = toks[i]
        if t in ("inc", "dec", "elements"):
            arg, j = atom(i + 1)
            if t == "inc":
                return arg + 1, j
            if t == "dec":
                return arg - 1, j
            return ",".join(str(x) for x in arg), j
        return atom(i)

    v, i = call(0)
    assert i == len(toks), (s, toks)
    return v


def _render(nodes, scope, out):
    for n in nodes:
        if n[0] == "text":
            out.append(n[1])
        elif n[0] == "expr":
            out.append(str(_eval_expr(n[1], scope)))
        elif n[0] == "if":
            cond = _eval_expr(n[1], scope)
            _render(n[2] if cond else n[3], scope, out)
        elif n[0] == "each":
            seq = _eval_expr(n[1], scope)
            for i, item in enumerate(seq):
                _render(n[2], scope + [{"@index": i, "this": item, "@last": i == len(seq) - 1}], out)


def render_handlebars(text: str, ctx: dict) -> str:
    tokens = []
    last = 0
    for m in _TAG.finditer(text):
        if m.start() > last:
            tokens.append(("text", text[last:m.start()]))
        tokens.append(("tag", m.group(1)))
        last = m.end()
    tokens.append(("text", text[last:]))
    nodes, _ = _parse(tokens, 0, set())
    out = []
    _render(nodes, [ctx], out)
    return "".join(out)


def prime_ctx(p: int) -> dict:
    """Template variables, computed as c_code_generator.rs:1086-1129 does."""
    pbits = p.bit_length()
    n64 = (pbits + 63) // 64
    nbits = n64 * 64
    half = p // 2
    inv = pow(p, -1, 1 << 64)
    np_ = (1 << 64) - inv
    lbo = ((1 << 64) >> (nbits - pbits)) - 1
    r2 = pow(2, nbits * 2, p)
    r3 = pow(2, nbits * 3, p)

    def limbs(v):
        return [str((v >> (64 * i)) & ((1 << 64) - 1)) + "U" for i in range(n64)]

    # the reference prints plain decimal limbs; a trailing U keeps g++ quiet
    # about >2^63 literals without changing the value
    return {
        "cannotOptimize": (p >> ((n64 - 1) * 64)) > ((((1 << 64) - 1) >> 1) - 1),
        "list0n64": list(range(n64)),
        "list0n64_1": list(range(n64 - 1)),
        "list1n64": list(range(1, n64)),
        "n64": n64,
        "fr_n64": n64,
        "qbits": pbits,
        "lboMask": hex(lbo),
        "fr_np": hex(np_),
        "fr_q_list": limbs(p),
        "fr_r2_list": limbs(r2),
        "fr_r3_list": limbs(r3),
        "half_list": limbs(half),
    }


def json_include_dir() -> str:
    for d in JSON_INC_CANDIDATES:
        if os.path.exists(os.path.join(d, "nlohmann", "json.hpp")):
            return d
    raise RuntimeError("nlohmann/json.hpp not found")


CXXFLAGS = ["-std=c++11", "-O3", "-fPIC", "-w", "-Wno-address-of-packed-member"]


def _run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("command failed: %s\n%s\n%s" % (" ".join(cmd), r.stdout[-4000:], r.stderr[-4000:]))


def have_reference() -> bool:
    return os.path.isdir(REF)


def build_prime(prime: str, force: bool = False) -> str:
    """Render + compile the reference field library and runtime shell for `prime`.
    Returns the path of libfr_<prime>.so."""
    so = os.path.join(OUT, "libfr_%s.so" % prime)
    pdir = os.path.join(OUT, prime)
    main_o = os.path.join(OUT, "rt_%s_main.o" % prime)
    calc_o = os.path.join(OUT, "rt_%s_calcwit.o" % prime)
    fr_o = os.path.join(OUT, "rt_%s_fr.o" % prime)
    if not have_reference():
        if os.path.exists(so):
            return so
        raise RuntimeError("reference tree absent and oracle/_ref not prebuilt")
    if not force and all(os.path.exists(x) for x in (so, main_o, calc_o, fr_o)):
        return so
    os.makedirs(pdir, exist_ok=True)
    ctx = prime_ctx(PRIMES[prime])
    for name in ("fr.cpp", "fr.hpp"):
        with open(os.path.join(REF, "generic", name)) as f:
            text = f.read()
        with open(os.path.join(pdir, name), "w") as f:
            f.write(render_handlebars(text, ctx))
    inc = ["-I", pdir, "-I", os.path.join(HERE, "gmp_shim"), "-I", json_include_dir()]
    _run(["g++"] + CXXFLAGS + inc + ["-c", os.path.join(pdir, "fr.cpp"), "-o", fr_o])
    _run(["g++", "-shared", "-o", so, fr_o, GMP_SO])
    # runtime shell, compiled from the reference tree in place
    _run(["g++"] + CXXFLAGS + inc + ["-c", os.path.join(REF, "common", "main.cpp"), "-o", main_o])
    _run(["g++"] + CXXFLAGS + inc + ["-c", os.path.join(REF, "common", "calcwit.cpp"), "-o", calc_o])
    return so


def build_goldilocks_runtime(force: bool = False):
    """The goldilocks runtime shell: c_elements/common64/main.cpp is a handlebars template over the prime
    (generate_main_cpp_file, c_code_generator.rs:945-962): rendered into oracle/_ref/goldilocks/main.cpp; calcwit.cpp compiled
    where it lies.  Returns the two objects."""
    gdir = os.path.join(OUT, "goldilocks")
    main_o, calc_o = os.path.join(OUT, "rt_goldilocks_main.o"), os.path.join(OUT, "rt_goldilocks_calcwit.o")
    if not force and os.path.exists(main_o) and os.path.exists(calc_o):
        return main_o, calc_o
    if not have_reference():
        raise RuntimeError("reference tree absent and oracle/_ref not prebuilt")
    os.makedirs(gdir, exist_ok=True)
    with open(os.path.join(REF, "common64", "main.cpp")) as f:
        text = f.read()
    with open(os.path.join(gdir, "main.cpp"), "w") as f:
        f.write(text.replace("{{prime}}", "18446744069414584321ull"))
    inc = ["-I", os.path.join(REF, "common64"), "-I", os.path.join(REF, "goldilocks"), "-I", os.path.join(HERE, "gmp_shim"),
           "-I", json_include_dir()]
    _run(["g++"] + CXXFLAGS + inc + ["-c", os.path.join(gdir, "main.cpp"), "-o", main_o])
    _run(["g++"] + CXXFLAGS + inc + ["-c", os.path.join(REF, "common64", "calcwit.cpp"), "-o", calc_o])
    return main_o, calc_o


def build_calculator(prime: str, circuit_cpp: str, out_bin: str, opt: str = "-O3") -> str:
    """Link a hand-lowered <circuit>.cpp with the reference runtime shell into
    `out_bin` (usage: out_bin input.json out.wtns; reads out_bin + '.dat')."""
    if prime == "goldilocks":
        main_o, calc_o = build_goldilocks_runtime()
        inc = ["-I", os.path.join(REF, "common64"), "-I", os.path.join(REF, "goldilocks"), "-I", os.path.join(HERE, "gmp_shim"),
               "-I", json_include_dir()]
        obj = out_bin + ".o"
        _run(["g++"] + [f if f != "-O3" else opt for f in CXXFLAGS] + inc + ["-c", circuit_cpp, "-o", obj])
        _run(["g++", "-o", out_bin, obj, main_o, calc_o, GMP_SO])
        os.remove(obj)
        return out_bin
    build_prime(prime)
    pdir = os.path.join(OUT, prime)
    inc = ["-I", pdir, "-I", os.path.join(HERE, "gmp_shim"), "-I", json_include_dir(),
           "-I", os.path.join(REF, "common")]
    obj = out_bin + ".o"
    flags = [f if f != "-O3" else opt for f in CXXFLAGS]
    _run(["g++"] + flags + inc + ["-c", circuit_cpp, "-o", obj])
    _run(["g++", "-o", out_bin, obj,
          os.path.join(OUT, "rt_%s_main.o" % prime),
          os.path.join(OUT, "rt_%s_calcwit.o" % prime),
          os.path.join(OUT, "rt_%s_fr.o" % prime), GMP_SO])
    os.remove(obj)
    return out_bin


def build_goldilocks(force: bool = False) -> str:
    """The reference's goldilocks field library is one header (c_elements/goldilocks/fr.hpp) of inline functions on
    uint64_t: compile it, where it lies, behind the C entry point of oracle/goldilocks_wrap.cpp."""
    so = os.path.join(OUT, "libfr_goldilocks.so")
    if not have_reference():
        if os.path.exists(so):
            return so
        raise RuntimeError("reference tree absent and oracle/_ref not prebuilt")
    if not force and os.path.exists(so):
        return so
    os.makedirs(OUT, exist_ok=True)
    _run(["g++"] + CXXFLAGS + ["-I", os.path.join(REF, "goldilocks"), "-I", os.path.join(HERE, "gmp_shim"), "-shared",
                               "-o", so, os.path.join(HERE, "goldilocks_wrap.cpp"), GMP_SO])
    return so


def build_all(force: bool = False):
    os.makedirs(OUT, exist_ok=True)
    for p in PRIMES:
        build_prime(p, force)
    build_goldilocks(force)


if __name__ == "__main__":
    if "--clean" in sys.argv:
        shutil.rmtree(OUT, ignore_errors=True)
    build_all(force="--force" in sys.argv)
    print("oracle/_ref built:", sorted(os.listdir(OUT)))
