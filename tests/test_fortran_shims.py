"""Static check of the source-only Fortran shims (no Fortran compiler in the image): every bind(C) interface in
fortran/dccm_b200_c.f90 must agree with the prototype of the same name in include/dccm_b200.h -- number of arguments,
by-value vs by-reference, and the C type each dummy is declared with."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def c_prototypes():
    text = open(os.path.join(ROOT, "include", "dccm_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    protos, returns = {}, {}
    for m in re.finditer(r"([A-Za-z_][\w\s\*]*?)\b(dccm_\w+)\s*\(([^;{}]*?)\)\s*;", text, flags=re.S):
        name, args = m.group(2), " ".join(m.group(3).split())
        params = [] if args in ("", "void") else [a.strip() for a in args.split(",")]
        protos[name] = params
        returns[name] = " ".join(m.group(1).split())
    return protos, returns


def fortran_interfaces():
    lines = open(os.path.join(ROOT, "fortran", "dccm_b200_c.f90")).read().splitlines()
    joined, cur = [], ""
    for ln in lines:                                   # join free-form continuation lines
        ln = ln.split("!")[0].rstrip()
        if not ln.strip():
            continue
        piece = ln.strip()
        if piece.startswith("&"):
            piece = piece[1:].lstrip()
        if piece.endswith("&"):
            cur += piece[:-1]
            continue
        joined.append(cur + piece)
        cur = ""
    out, k = {}, 0
    while k < len(joined):
        m = re.match(r'(function|subroutine)\s+(\w+)\s*\(([^)]*)\)\s*bind\(C,\s*name="(\w+)"\)(?:\s*result\((\w+)\))?',
                     joined[k], flags=re.I)
        if not m:
            k += 1
            continue
        fname, dummies, cname = m.group(2), [d.strip().lower() for d in m.group(3).split(",") if d.strip()], m.group(4)
        res = m.group(5).lower() if m.group(5) else None         # None: a subroutine, i.e. a void C function
        assert (m.group(1).lower() == "function") == (res is not None), joined[k]
        decl = {}
        k += 1
        while not re.match(r"end\s+(function|subroutine)", joined[k], flags=re.I):
            for stmt in joined[k].split(";"):
                if "::" not in stmt:
                    continue
                spec, names = stmt.split("::")
                for nm in re.split(r",(?![^()]*\))", names):
                    decl[re.sub(r"\(.*\)", "", nm).strip().lower()] = (spec.strip().lower(), "(" in nm)
            k += 1
        out[cname] = (fname, dummies, decl, res)
    return out


KIND = {"int": "c_int", "int32_t": "c_int32_t", "int64_t": "c_int64_t", "double": "c_double", "char": "c_char"}


def test_every_fortran_interface_matches_its_c_prototype():
    (protos, returns), ifaces = c_prototypes(), fortran_interfaces()
    assert len(ifaces) >= 19
    for cname, (fname, dummies, decl, res) in ifaces.items():
        assert fname == cname and cname in protos, f"{cname}: not declared in include/dccm_b200.h"
        params = protos[cname]
        assert len(params) == len(dummies), f"{cname}: {len(dummies)} Fortran dummies vs {len(params)} C parameters"
        ret = returns[cname].replace("const", "").strip()
        if res is None:
            assert ret == "void", f"{cname}: a subroutine binds a C function returning '{ret}'"
        else:
            assert res in decl, f"{cname}: result variable undeclared"
            want = "type(c_ptr)" if "*" in ret else KIND[ret]
            assert want in decl[res][0], f"{cname}: result declared '{decl[res][0]}' for C return type '{ret}'"
        for p, d in zip(params, dummies):
            assert d in decl, f"{cname}: dummy {d} has no declaration"
            spec, is_array = decl[d]
            by_value = "value" in [s.strip() for s in spec.split(",")]
            c_pointer = "*" in p
            base = p.replace("const", "").replace("*", " ").split()[0]
            if not c_pointer:                          # C scalar: Fortran must pass by value with the matching kind
                assert by_value and not is_array, f"{cname}: {d} must have the VALUE attribute ({p})"
                assert KIND[base] in spec, f"{cname}: {d} declared '{spec}' for C '{p}'"
            elif "type(c_ptr)" in spec:                # opaque handle: by value for T*, by reference for T**
                assert by_value == (p.count("*") == 1), f"{cname}: handle {d} ({p})"
            else:                                      # array / scalar by reference
                assert not by_value, f"{cname}: {d} is a pointer in C ({p}) but VALUE in Fortran"
                assert KIND[base] in spec, f"{cname}: {d} declared '{spec}' for C '{p}'"
                if "const" in p:
                    assert "intent(in)" in spec, f"{cname}: {d} is const in C, must be intent(in)"


def test_shims_only_call_declared_interfaces():
    """every dccm_* name the shim sources call is declared in the interface module (and hence in the header)"""
    ifaces = set(fortran_interfaces()) | {"dccm_check", "dccm_b200_c"}
    assert {"grid_mapping_util.f90", "grid_mapping_util_jones99.f90"} <= set(os.listdir(os.path.join(ROOT, "fortran")))
    for f in os.listdir(os.path.join(ROOT, "fortran")):
        text = open(os.path.join(ROOT, "fortran", f)).read()
        text = "\n".join(l.split("!")[0] for l in text.splitlines())
        for name in set(re.findall(r"\b(dccm_\w+)\s*\(", text)):
            if name.lower() in ("dccm_check",) or name in ifaces:
                continue
            # module-level names of the shim itself (handles, helper procedures) are declared in the same file
            assert re.search(rf"(subroutine|function)\s+{name}\b", text, flags=re.I), f"{f}: {name} is not bound"
