"""The Julia glue (julia/B200Interpolations) cannot run in this image (no Julia), so it is checked statically against
the C ABI it binds: every `ccall` names an exported symbol of include/b200itp.h with the right arity, return type and
argument kinds, and the Julia mirrors of the two C structs list the same fields in the same order with the same widths.
CPU only."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
JL = os.path.join(ROOT, "julia", "B200Interpolations", "src", "B200Interpolations.jl")
HDR = os.path.join(ROOT, "include", "b200itp.h")


def _split_top(s):
    """Split on commas that are not nested in (), {} or []."""
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "({[":
            depth += 1
        elif ch in ")}]":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


def header_prototypes():
    text = re.sub(r"/\*.*?\*/", "", open(HDR).read(), flags=re.S)
    protos = {}
    for m in re.finditer(r"B200ITP_API\s+([\w\s\*]+?)\s*\b(b200itp_\w+)\s*\(([^;]*?)\)\s*;", text, flags=re.S):
        ret, name, args = m.group(1).strip(), m.group(2), " ".join(m.group(3).split())
        argl = [] if args in ("void", "") else _split_top(args)
        protos[name] = (ret, argl)
    return protos


def c_kind(decl):
    d = decl.replace("const", " ").strip()
    if "*" in d:
        return "ptr"
    base = d.split()[0] if len(d.split()) == 1 else " ".join(d.split()[:-1])
    return {"int": "i32", "int32_t": "i32", "uint32_t": "u32", "int64_t": "i64", "uint64_t": "u64", "size_t": "usize",
            "double": "f64", "void": "void"}[base]


def jl_kind(t):
    t = t.strip()
    if t.startswith(("Ptr{", "Ref{")):
        return "ptr"
    return {"Cint": "i32", "Int32": "i32", "UInt32": "u32", "Int64": "i64", "UInt64": "u64", "Csize_t": "usize",
            "Cdouble": "f64", "Float64": "f64", "Cvoid": "void"}[t]


def julia_ccalls():
    text = open(JL).read()
    calls = []
    for m in re.finditer(r"ccall\(\(:(\w+),\s*libb200itp\),\s*([\w{}]+),\s*\(", text):
        start = m.end()
        depth, i = 1, start
        while depth:
            depth += {"(": 1, ")": -1}.get(text[i], 0)
            i += 1
        types = _split_top(text[start:i - 1])
        types = [t for t in types if t]
        # the call arguments follow the type tuple up to the closing paren of ccall
        j, depth = i, 1
        while depth:
            depth += {"(": 1, ")": -1}.get(text[j], 0)
            j += 1
        nargs = len(_split_top(text[i:j - 1].lstrip(", \n")))
        calls.append((m.group(1), m.group(2), types, nargs))
    return calls


def test_every_ccall_matches_the_header():
    protos = header_prototypes()
    calls = julia_ccalls()
    assert len(calls) >= 15
    seen = set()
    for name, ret, types, nargs in calls:
        assert name in protos, f"{name} is not declared in include/b200itp.h"
        cret, cargs = protos[name]
        assert jl_kind(ret) == c_kind(cret + " x"), (name, ret, cret)
        assert len(types) == len(cargs), (name, types, cargs)
        assert nargs == len(types), (name, "ccall passes", nargs, "values for", len(types), "types")
        for jt, ca in zip(types, cargs):
            assert jl_kind(jt) == c_kind(ca), (name, jt, ca)
        seen.add(name)
    # the seams of the hot path are all bound
    for need in ("b200itp_malloc", "b200itp_free", "b200itp_memcpy_h2d", "b200itp_memcpy_d2h", "b200itp_sync",
                 "b200itp_copy_with_padding", "b200itp_prefilter_axis", "b200itp_eval", "b200itp_eval_sorted",
                 "b200itp_eval_sorted_scratch_bytes", "b200itp_gradient_sorted", "b200itp_hessian", "b200itp_eval_grid",
                 "b200itp_eval_grid_scratch_bytes", "b200itp_out_dtype", "b200itp_last_error",
                 "b200itp_comm_init", "b200itp_comm_destroy", "b200itp_bcast_coefs", "b200itp_prefilter_sharded"):
        assert need in seen, need


def _c_struct_fields(name):
    text = re.sub(r"/\*.*?\*/", "", open(HDR).read(), flags=re.S)
    body = re.search(r"typedef struct " + name + r"\s*\{(.*?)\}\s*" + name + r"\s*;", text, flags=re.S).group(1)
    fields = []
    for stmt in body.split(";"):
        stmt = " ".join(stmt.split())
        if not stmt:
            continue
        m = re.match(r"(const\s+)?(\w+)\s*(\*?)\s*(.*)", stmt)
        ctype, star, rest = m.group(2), m.group(3), m.group(4)
        for item in rest.split(","):
            item = item.strip()
            ptr = bool(star) or item.startswith("*")
            fm = re.match(r"\*?\s*(\w+)(\[(\w+)\])?", item)
            fields.append((fm.group(1), "ptr" if ptr else ctype, bool(fm.group(2))))
    return fields


def _jl_struct_fields(name):
    text = open(JL).read()
    body = re.search(r"\nstruct " + name + r"\n(.*?)\nend\n", text, flags=re.S).group(1)
    out = []
    for line in body.splitlines():
        line = line.strip()
        if not line or line.startswith("#"):
            continue
        fname, ftype = [x.strip() for x in line.split("::")]
        out.append((fname, ftype))
    return out


def test_julia_structs_mirror_the_c_structs():
    width = {"int32_t": "Int32", "int64_t": "Int64", "double": "Float64"}
    for cname, jname in (("b200itp_desc", "Desc"), ("b200itp_axis_system", "AxisSystem")):
        cf, jf = _c_struct_fields(cname), _jl_struct_fields(jname)
        assert [f[0] for f in cf] == [f[0] for f in jf], (cname, [f[0] for f in cf], [f[0] for f in jf])
        for (fname, ctype, is_array), (_, jtype) in zip(cf, jf):
            if ctype == "ptr":
                assert jtype.startswith("Ptr{"), (fname, jtype)
            elif is_array:
                assert jtype == "NTuple{4," + width[ctype] + "}", (fname, ctype, jtype)
            else:
                assert jtype == width[ctype], (fname, ctype, jtype)


def test_reference_gpu_cases_are_covered():
    """test/runtests.jl exercises the cases of the reference's array-type-generic testset (test/gpu_support.jl): 1-D and 2-D
    B-splines, bare / scaled / Flat / filled, values, gradients, hessians, host and device coordinates, with `==`."""
    t = open(os.path.join(ROOT, "julia", "B200Interpolations", "test", "runtests.jl")).read()
    for needle in ("BSpline(Cubic(Line(OnGrid())))", "BSpline(Linear())", "extrapolate(sitp, Flat())", "extrapolate(sitp, 0.0)",
                   "Interpolations.gradient", "Interpolations.hessian", "cpu == collect(on_host_args) == collect(on_dev_args)",
                   "adapt(Array{Float32}, obj)", "b200itp_prefilter_set_strict", "evaluate_sharded(replicate(comm, one), xs...)",
                   "interpolate_sharded(comm, A, it)"):
        assert needle in t, needle
