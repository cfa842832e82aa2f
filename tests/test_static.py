"""Static checks that protect the code that only runs on a GPU box (what cannot run here is at least
checked here): no undefined names anywhere, and the three places that must agree about every native kernel entry point
-- the ``extern "C"`` definition in ``ops/csrc/*.cu``, the ctypes signature table in ``ops/lib.py`` and the call sites
in the package -- do agree on the number of arguments."""
import ast
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


def test_no_undefined_names_in_the_repo():
    import lint_names

    paths = [os.path.join(ROOT, p) for p in ("trainingjob_operator_b200", "tools", "tests", "baseline", "bench.py",
                                             "__graft_entry__.py")]
    problems = []
    for p in paths:
        if os.path.isdir(p):
            for d, _dirs, names in os.walk(p):
                if "__pycache__" in d or "/scratch" in d:
                    continue
                for n in names:
                    if n.endswith(".py"):
                        problems += lint_names.check_file(os.path.join(d, n))
        else:
            problems += lint_names.check_file(p)
    assert not problems, "\n".join(problems)


def _c_entry_points(with_types=False):
    """name -> number of parameters (or the parameter declarations) of every ``aitj_*`` definition in the .cu sources."""
    out = {}
    csrc = os.path.join(ROOT, "trainingjob_operator_b200", "ops", "csrc")
    for fn in sorted(os.listdir(csrc)):
        if not fn.endswith(".cu"):
            continue
        src = open(os.path.join(csrc, fn)).read()
        src = re.sub(r"//[^\n]*", "", src)
        for m in re.finditer(r"\b(?:int|void|long long|float)\s+(aitj_\w+)\s*\(([^)]*)\)\s*\{", src):
            params = m.group(2).strip()
            plist = [] if params in ("", "void") else [p.strip() for p in params.split(",") if p.strip()]
            out[m.group(1)] = plist if with_types else len(plist)
    return out


def test_ctypes_argument_types_match_the_c_declarations():
    """A pointer passed as a 32-bit int or an element count passed as ``int`` where the kernel takes ``long long`` is a
    bug that only shows on large tensors on the GPU box."""
    import ctypes

    from trainingjob_operator_b200.ops import lib

    for name, params in _c_entry_points(with_types=True).items():
        if name not in lib._SIGS:
            continue
        for i, (p, ct) in enumerate(zip(params, lib._SIGS[name])):
            if "*" in p:
                want = ctypes.c_void_p
            elif re.match(r"(const\s+)?(long long|int64_t|size_t|uint64_t|unsigned long long)\b", p):
                want = ctypes.c_longlong
            elif re.match(r"(const\s+)?float\b", p):
                want = ctypes.c_float
            elif re.match(r"(const\s+)?(int|unsigned|unsigned int|uint32_t|bool)\b", p):
                want = ctypes.c_int
            else:
                raise AssertionError(f"{name} argument {i}: unrecognised C type in '{p}'")
            assert ct is want, f"{name} argument {i}: C declares '{p}', ctypes table says {ct.__name__}"


def test_kernel_entry_points_agree_between_cuda_sources_ctypes_table_and_call_sites():
    from trainingjob_operator_b200.ops import lib

    defined = _c_entry_points()
    assert len(defined) > 20, defined
    for name, sig in lib._SIGS.items():
        assert name in defined, f"{name} is declared in ops/lib.py but not defined in ops/csrc"
        assert len(sig) == defined[name], f"{name}: {len(sig)} ctypes arguments, {defined[name]} parameters in the .cu"
    # call sites: lib.call("aitj_x", ...) and lib.load().aitj_x(...)
    pkg = os.path.join(ROOT, "trainingjob_operator_b200")
    seen = 0
    for d, _dirs, names in os.walk(pkg):
        for n in names:
            if not n.endswith(".py"):
                continue
            path = os.path.join(d, n)
            tree = ast.parse(open(path).read(), path)
            for node in ast.walk(tree):
                if not isinstance(node, ast.Call):
                    continue
                fname, nargs = None, None
                f = node.func
                if isinstance(f, ast.Attribute) and f.attr == "call" and node.args and \
                        isinstance(node.args[0], ast.Constant) and str(node.args[0].value).startswith("aitj_"):
                    fname, nargs = node.args[0].value, len(node.args) - 1
                elif isinstance(f, ast.Attribute) and f.attr.startswith("aitj_"):
                    fname, nargs = f.attr, len(node.args)
                if fname is None or any(isinstance(a, ast.Starred) for a in node.args):
                    continue
                assert fname in lib._SIGS, f"{path}:{node.lineno}: {fname} has no ctypes signature"
                assert nargs == len(lib._SIGS[fname]), \
                    f"{path}:{node.lineno}: {fname} called with {nargs} arguments, signature has {len(lib._SIGS[fname])}"
                seen += 1
    assert seen >= 20, seen


def test_every_environment_knob_is_documented():
    """docs/ENVIRONMENT.md lists every ``AITJ_*`` variable the package reads or sets (compile-time macros excepted)."""
    doc = open(os.path.join(ROOT, "docs", "ENVIRONMENT.md")).read()
    macros = {"AITJ_DISPATCH", "AITJ_PAIR", "AITJ_MBAR_DEBUG", "AITJ_ATTN_DEBUG"}
    used = set()
    pkg = os.path.join(ROOT, "trainingjob_operator_b200")
    for d, _dirs, names in os.walk(pkg):
        for n in names:
            if n.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                used |= set(re.findall(r"\bAITJ_[A-Z0-9_]+\b", open(os.path.join(d, n), errors="replace").read()))
    missing = sorted(v for v in used - macros if v not in doc)
    assert not missing, missing
