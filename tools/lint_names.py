"""Undefined-name check over the repo's Python sources (no linter is installed in the image).

Most of the worker / kernel-binding code only runs on a GPU box, so a misspelt variable in one of those branches would
first be seen there.  This walks every module's AST with proper scoping (module / class / function / comprehension,
``global`` / ``nonlocal``, imports, assignment targets, ``with`` / ``for`` / ``except`` / walrus targets, ``match``
captures) and reports loads of names that are bound nowhere visible and are not builtins.  It is deliberately
conservative: a name bound anywhere in a scope counts as bound everywhere in it (no flow analysis), so it reports
typos, not use-before-assignment.

    python tools/lint_names.py [paths...]        # default: the package, tools/, tests/, bench.py, __graft_entry__.py
"""
from __future__ import annotations

import ast
import builtins
import os
import sys
from typing import Dict, List, Optional, Set

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILTINS = set(dir(builtins)) | {"__file__", "__name__", "__doc__", "__package__", "__spec__", "__path__", "__class__",
                                 "__builtins__", "__debug__", "__loader__", "__annotations__", "__dict__"}


class Scope:
    def __init__(self, kind: str, parent: Optional["Scope"]):
        self.kind, self.parent = kind, parent
        self.bound: Set[str] = set()
        self.globals: Set[str] = set()
        self.nonlocals: Set[str] = set()
        self.star_import = False


def targets(node: ast.AST, out: Set[str]) -> None:
    if isinstance(node, ast.Name):
        out.add(node.id)
    elif isinstance(node, (ast.Tuple, ast.List)):
        for e in node.elts:
            targets(e, out)
    elif isinstance(node, ast.Starred):
        targets(node.value, out)


class Binder(ast.NodeVisitor):
    """First pass over ONE scope's body: which names does it bind (nested scopes are not entered)."""

    def __init__(self, scope: Scope):
        self.s = scope

    def bind(self, name: str) -> None:
        self.s.bound.add(name)

    def visit_FunctionDef(self, n):
        self.bind(n.name)
        for d in n.decorator_list:
            self.visit(d)

    visit_AsyncFunctionDef = visit_FunctionDef

    def visit_ClassDef(self, n):
        self.bind(n.name)

    def visit_Lambda(self, n):
        pass

    def _comp(self, n):
        # the first iterable is evaluated in the enclosing scope; walrus targets inside bind there too
        for sub in ast.walk(n):
            if isinstance(sub, ast.NamedExpr):
                targets(sub.target, self.s.bound)

    visit_ListComp = visit_SetComp = visit_DictComp = visit_GeneratorExp = _comp

    def visit_Import(self, n):
        for a in n.names:
            self.bind((a.asname or a.name).split(".")[0])

    def visit_ImportFrom(self, n):
        for a in n.names:
            if a.name == "*":
                self.s.star_import = True
            else:
                self.bind(a.asname or a.name)

    def visit_Global(self, n):
        self.s.globals.update(n.names)

    def visit_Nonlocal(self, n):
        self.s.nonlocals.update(n.names)

    def visit_Assign(self, n):
        for t in n.targets:
            targets(t, self.s.bound)
        self.generic_visit(n)

    def visit_AugAssign(self, n):
        targets(n.target, self.s.bound)
        self.generic_visit(n)

    def visit_AnnAssign(self, n):
        targets(n.target, self.s.bound)
        self.generic_visit(n)

    def visit_NamedExpr(self, n):
        targets(n.target, self.s.bound)
        self.generic_visit(n)

    def visit_For(self, n):
        targets(n.target, self.s.bound)
        self.generic_visit(n)

    visit_AsyncFor = visit_For

    def visit_With(self, n):
        for it in n.items:
            if it.optional_vars is not None:
                targets(it.optional_vars, self.s.bound)
        self.generic_visit(n)

    visit_AsyncWith = visit_With

    def visit_ExceptHandler(self, n):
        if n.name:
            self.bind(n.name)
        self.generic_visit(n)

    def visit_MatchAs(self, n):
        if n.name:
            self.bind(n.name)
        self.generic_visit(n)

    def visit_MatchStar(self, n):
        if n.name:
            self.bind(n.name)

    def visit_MatchMapping(self, n):
        if n.rest:
            self.bind(n.rest)
        self.generic_visit(n)

    def visit_Delete(self, n):
        self.generic_visit(n)


class Checker(ast.NodeVisitor):
    def __init__(self, path: str):
        self.path = path
        self.problems: List[str] = []
        self.scope: Optional[Scope] = None
        self.module: Optional[Scope] = None

    # ------------------------------------------------------------------ scope handling
    def resolve(self, name: str) -> bool:
        s = self.scope
        first = True
        while s is not None:
            if name in s.globals:
                return name in self.module.bound or name in BUILTINS or self.module.star_import
            # class bodies are not visible from nested functions
            if (first or s.kind != "class") and name in s.bound:
                return True
            if s.star_import:
                return True
            first = False
            s = s.parent
        return name in BUILTINS

    def enter(self, kind: str, body_nodes, args: Optional[ast.arguments] = None) -> Scope:
        s = Scope(kind, self.scope)
        if args is not None:
            for a in list(args.posonlyargs) + list(args.args) + list(args.kwonlyargs):
                s.bound.add(a.arg)
            if args.vararg:
                s.bound.add(args.vararg.arg)
            if args.kwarg:
                s.bound.add(args.kwarg.arg)
        b = Binder(s)
        for n in body_nodes:
            b.visit(n)
        return s

    def run_scope(self, s: Scope, body_nodes) -> None:
        prev, self.scope = self.scope, s
        for n in body_nodes:
            self.visit(n)
        self.scope = prev

    # ------------------------------------------------------------------ visitors
    def visit_Module(self, n):
        self.module = self.enter("module", n.body)
        self.run_scope(self.module, n.body)

    def _function(self, n):
        for d in n.decorator_list:
            self.visit(d)
        for d in list(n.args.defaults) + [d for d in n.args.kw_defaults if d is not None]:
            self.visit(d)
        s = self.enter("function", n.body, n.args)
        self.run_scope(s, n.body)

    visit_FunctionDef = visit_AsyncFunctionDef = _function

    def visit_Lambda(self, n):
        for d in list(n.args.defaults) + [d for d in n.args.kw_defaults if d is not None]:
            self.visit(d)
        s = self.enter("function", [], n.args)
        b = Binder(s)
        b.visit(n.body)
        self.run_scope(s, [n.body])

    def visit_ClassDef(self, n):
        for d in n.decorator_list + n.bases + [k.value for k in n.keywords]:
            self.visit(d)
        s = self.enter("class", n.body)
        self.run_scope(s, n.body)

    def _comp(self, n):
        gens = n.generators
        self.visit(gens[0].iter)                       # evaluated in the enclosing scope
        s = Scope("function", self.scope)
        for g in gens:
            targets(g.target, s.bound)
        prev, self.scope = self.scope, s
        for i, g in enumerate(gens):
            if i:
                self.visit(g.iter)
            for c in g.ifs:
                self.visit(c)
        if isinstance(n, ast.DictComp):
            self.visit(n.key)
            self.visit(n.value)
        else:
            self.visit(n.elt)
        self.scope = prev

    visit_ListComp = visit_SetComp = visit_DictComp = visit_GeneratorExp = _comp

    def visit_Name(self, n):
        if isinstance(n.ctx, ast.Load) and not self.resolve(n.id):
            self.problems.append(f"{self.path}:{n.lineno}: undefined name '{n.id}'")

    def visit_AnnAssign(self, n):
        # annotations are strings under `from __future__ import annotations`; check target / value only
        if n.value is not None:
            self.visit(n.value)
        if not isinstance(n.target, ast.Name):
            self.visit(n.target)

    def visit_arg(self, n):
        pass                                            # parameter annotations: not evaluated

    def visit_Constant(self, n):
        pass


def check_file(path: str) -> List[str]:
    src = open(path, encoding="utf-8").read()
    try:
        tree = ast.parse(src, path)
    except SyntaxError as e:
        return [f"{path}:{e.lineno}: syntax error: {e.msg}"]
    # return annotations are not evaluated either
    for node in ast.walk(tree):
        if isinstance(node, (ast.FunctionDef, ast.AsyncFunctionDef)):
            node.returns = None
    c = Checker(os.path.relpath(path, ROOT))
    c.visit(tree)
    return c.problems


def main(argv: List[str]) -> int:
    paths = argv or [os.path.join(ROOT, p) for p in ("trainingjob_operator_b200", "tools", "tests", "baseline",
                                                      "bench.py", "__graft_entry__.py")]
    files: List[str] = []
    for p in paths:
        if os.path.isdir(p):
            for d, _dirs, names in os.walk(p):
                if "__pycache__" in d or "/scratch" in d:
                    continue
                files += [os.path.join(d, n) for n in names if n.endswith(".py")]
        elif os.path.exists(p):
            files.append(p)
    problems: List[str] = []
    for f in sorted(files):
        problems += check_file(f)
    for p in problems:
        print(p)
    print(f"{len(files)} files checked, {len(problems)} undefined name(s)")
    return 1 if problems else 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
