"""The ISO_C_BINDING shim (fortran/fen_gpu_mod.f90) cannot be compiled in this image (no Fortran compiler), so its
agreement with the C ABI is checked textually: every bind(C) interface names a function that include/fen_gpu.h
declares with the same number of arguments, passes scalars by value exactly where the C prototype takes them by
value, the bind(C) derived types mirror the C structs member by member, and the enum constants carry the header's
values.  A typo in the shim would otherwise only show up on the first machine that has mpif90."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HDR = open(os.path.join(ROOT, "include", "fen_gpu.h")).read()
F90 = open(os.path.join(ROOT, "fortran", "fen_gpu_mod.f90")).read()


def _strip_c_comments(s):
    return re.sub(r"/\*.*?\*/", " ", s, flags=re.S)


def c_prototypes():
    """name -> list of (is_pointer, base type) per parameter."""
    src = _strip_c_comments(HDR)
    out = {}
    for m in re.finditer(r"\b([A-Za-z_][\w\s\*]*?)\b(fen_gpu_\w+)\s*\(([^;{]*?)\)\s*;", src):
        name, args = m.group(2), m.group(3).strip()
        params = []
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                is_ptr = "*" in a or "[" in a or re.search(r"\bfen_(forcing|distance)_fn\b", a) is not None
                params.append((is_ptr, a))
        out[name] = params
    return out


def f90_interfaces():
    """name -> (argument names, set of names passed by value)."""
    # join continuation lines
    src = re.sub(r"&\s*\n\s*", " ", F90)
    out = {}
    for m in re.finditer(r"function\s+(\w+)\s*\(([^)]*)\)\s*bind\(C,\s*name='(\w+)'\)\s*result\((\w+)\)(.*?)end function",
                         src, flags=re.S | re.I):
        fname, args, cname, res, body = m.groups()
        assert fname == cname, (fname, cname)
        names = [a.strip() for a in args.split(",") if a.strip()]
        by_value = set()
        for line in body.splitlines():
            if "::" not in line:
                continue
            decl, vars_ = line.split("::", 1)
            vs = [v.strip().split("(")[0] for v in vars_.split(",")]
            if re.search(r"\bvalue\b", decl, flags=re.I):
                by_value.update(vs)
        out[cname] = (names, by_value)
    return out


def test_every_shim_interface_matches_a_c_prototype():
    protos, ifaces = c_prototypes(), f90_interfaces()
    assert len(ifaces) >= 30
    for name, (args, by_value) in ifaces.items():
        assert name in protos, "%s is bound by the shim but not declared in fen_gpu.h" % name
        params = protos[name]
        assert len(params) == len(args), (name, params, args)
        for (is_ptr, ctext), a in zip(params, args):
            if is_ptr:
                # pointers: either a c_ptr / c_funptr passed by value, or a Fortran variable passed by reference
                continue
            assert a in by_value, "%s: C takes `%s` by value, the shim must declare %s with VALUE" % (name, ctext, a)
        # and nothing is passed by value that C takes through a non-void pointer to a scalar it writes
        for (is_ptr, ctext), a in zip(params, args):
            if is_ptr and re.search(r"\b(double|int)\s*\*", ctext) and "const" not in ctext and a in by_value:
                # allowed only for raw buffers handed over as c_ptr (host arrays)
                assert re.search(r"\b(host|plane|handle)", ctext) or a in ("host", "plane"), (name, ctext, a)


def _c_struct(name):
    src = _strip_c_comments(HDR)
    body = re.search(r"typedef struct %s\s*\{(.*?)\}\s*%s\s*;" % (name, name), src, flags=re.S).group(1)
    members = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        ctype, rest = decl.split(None, 1)
        for v in rest.split(","):
            v = v.strip()
            m = re.match(r"(\w+)(?:\[(\d+)\])?$", v)
            members.append((ctype, m.group(1).lower(), int(m.group(2) or 1)))
    return members


def _f_type(name):
    body = re.search(r"type,\s*bind\(C\)\s*::\s*%s\s*\n(.*?)end type" % name, F90, flags=re.S | re.I).group(1)
    kinds = {"integer(c_int)": "int", "real(c_double)": "double"}
    members = []
    for line in body.splitlines():
        line = line.split("!")[0].strip()
        if "::" not in line:
            continue
        decl, vars_ = [x.strip() for x in line.split("::", 1)]
        for v in re.findall(r"\w+(?:\(\d+\))?", vars_):
            m = re.match(r"(\w+)(?:\((\d+)\))?$", v)
            members.append((kinds[decl.lower().replace(" ", "")], m.group(1).lower(), int(m.group(2) or 1)))
    return members


def test_the_shim_binds_the_whole_header():
    protos, ifaces = c_prototypes(), f90_interfaces()
    assert set(protos) == set(ifaces), set(protos) ^ set(ifaces)


def test_bind_c_types_mirror_the_c_structs():
    for name in ("fen_grid_desc", "fen_ns_params", "fen_mf_params"):
        assert _f_type(name) == _c_struct(name), name


def test_enum_constants_carry_the_headers_values():
    src = _strip_c_comments(HDR)
    cvals = {k: int(v) for k, v in re.findall(r"\b(FEN_[A-Z0-9_]+)\s*=\s*(-?\d+)", src)}
    fvals = {}
    for line in F90.splitlines():
        if "parameter" in line.lower() and "FEN_" in line:
            for k, v in re.findall(r"\b(FEN_[A-Z0-9_]+)\s*=\s*(-?\d+)", line):
                fvals[k] = int(v)
    assert len(fvals) >= 15
    for k, v in fvals.items():
        assert cvals.get(k) == v, (k, v, cvals.get(k))
