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


def _code_lines():
    """Source lines without comments and strings, continuation lines joined, lower case."""
    out, cur = [], ""
    for raw in F90.splitlines():
        line = re.sub(r"'[^']*'|\"[^\"]*\"", "''", raw)          # strings first: they may hold '!'
        line = line.split("!")[0].rstrip()
        if not line.strip():
            continue
        if cur:
            line = cur + " " + line.strip().lstrip("&")
            cur = ""
        if line.rstrip().endswith("&"):
            cur = line.rstrip()[:-1]
            continue
        out.append(line.strip().lower())
    assert not cur
    return out


def test_shim_block_structure_is_balanced():
    """What a compiler's parser would reject first in a file that has never been compiled: every module / interface /
    type / function / subroutine / if-then / do / select / block construct is closed by the matching END, in order, and no
    line exceeds the 132 columns of free-form Fortran."""
    assert max(len(line) for line in F90.splitlines()) <= 132
    stack = []
    opener = re.compile(r"^(?:(?:pure|elemental|recursive)\s+)*(?:[\w\(\)=, ]*?\s)?(function|subroutine)\s+\w+")
    for n, line in enumerate(_code_lines(), 1):
        m_end = re.match(r"^end\s*(module|interface|type|function|subroutine|if|do|select|block|program)?\b", line)
        if m_end:
            kind = m_end.group(1)
            assert stack, "END without an open construct: %r" % line
            top = stack.pop()
            assert kind is None or kind == top, "%r closes %r" % (line, top)
            continue
        if re.match(r"^module\s+(?!procedure)\w+", line):
            stack.append("module")
        elif re.match(r"^(abstract\s+)?interface\b", line):
            stack.append("interface")
        elif re.match(r"^type\s*(,[^:]*)?::\s*\w+", line) or re.match(r"^type\s+\w+\s*$", line):
            stack.append("type")
        elif opener.match(line) and not line.startswith(("procedure", "call ", "use ")) and "::" not in line.split("function")[0].split("subroutine")[0]:
            stack.append(opener.match(line).group(1))
        elif re.match(r"^(\w+\s*:\s*)?if\s*\(.*\)\s*then$", line):
            stack.append("if")
        elif re.match(r"^(\w+\s*:\s*)?do\b", line):
            stack.append("do")
        elif re.match(r"^(\w+\s*:\s*)?select\s+(case|type)\b", line):
            stack.append("select")
        elif re.match(r"^(\w+\s*:\s*)?block$", line):
            stack.append("block")
    assert stack == [], "unclosed constructs: %r" % stack


def test_shim_dummy_arguments_are_declared():
    """Every dummy argument of every procedure in the shim has a declaration in that procedure (implicit none is in
    force: an undeclared dummy is a compile error), and every result variable too."""
    src = "\n".join(_code_lines())
    pat = re.compile(r"^(?:[\w\(\)=\*, ]*?\s)?(function|subroutine)\s+(\w+)\s*\(([^)]*)\)([^\n]*)\n(.*?)^end\s*(?:function|subroutine)",
                     re.S | re.M)
    checked = 0
    for m in pat.finditer(src):
        kind, name, args, tail, body = m.group(1), m.group(2), m.group(3), m.group(4), m.group(5)
        names = [a.strip() for a in args.split(",") if a.strip()]
        res = re.search(r"result\s*\((\w+)\)", tail)
        if res:
            names.append(res.group(1))
        elif kind == "function" and not re.match(r"^\s*(real|integer|logical|type|character)", m.group(0)):
            names.append(name)
        decl = " ".join(line for line in body.splitlines() if "::" in line or line.strip().startswith("procedure"))
        for a in names:
            assert re.search(r"\b%s\b" % re.escape(a), decl), "%s: dummy %r is not declared" % (name, a)
        checked += 1
    assert checked > 100


def test_shim_calls_pass_the_declared_number_of_arguments():
    """Part 2 of the shim calls the bind(C) functions of part 1: every call site passes as many actual arguments as the
    interface (and so the C prototype) declares."""
    ifaces = f90_interfaces()
    lines = _code_lines()
    end_iface = max(i for i, l in enumerate(lines) if re.match(r"^end\s*interface", l))
    body = "\n".join(lines[end_iface + 1:])
    calls = 0
    for m in re.finditer(r"\b(fen_gpu_\w+)\s*\(", body):
        name = m.group(1)
        depth, i, nargs, seen = 1, m.end(), 0, False
        while depth:
            ch = body[i]
            if ch == "(":
                depth += 1
            elif ch == ")":
                depth -= 1
            elif ch == "," and depth == 1:
                nargs += 1
            elif not ch.isspace():
                seen = True
            i += 1
        nargs = nargs + 1 if seen else 0
        lowered = {k.lower(): v for k, v in ifaces.items()}
        assert name in lowered, "%s is called but has no interface" % name
        assert nargs == len(lowered[name][0]), "%s called with %d arguments, interface has %d" % (
            name, nargs, len(lowered[name][0]))
        calls += 1
    assert calls >= 50                       # most entry points are wrapped at least once (the rest are bound only)
