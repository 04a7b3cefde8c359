"""Static guard for the branches the CPU suite cannot execute (CUDA-only paths): every global name a function of the
package loads must exist in its module (or in builtins).  Catches the missing-import class of bug before a GPU box does."""
import builtins
import dis
import importlib
import os
import pkgutil
import sys
import types

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SKIP = {"bagua_b200._C", "bagua_b200._C_torch", "bagua_b200.contrib.lightning"}  # binaries; lightning needs an optional package


def _code_objects(code):
    yield code
    for c in code.co_consts:
        if isinstance(c, types.CodeType):
            yield from _code_objects(c)


def _module_sources():
    for pkg_name in ("bagua_b200", "bagua", "bagua_core"):
        pkg = importlib.import_module(pkg_name)
        yield pkg
        for info in pkgutil.walk_packages(pkg.__path__, pkg_name + "."):
            if info.name in SKIP or info.name.rsplit(".", 1)[-1].startswith("_C"):
                continue
            try:
                yield importlib.import_module(info.name)
            except ImportError:
                continue  # optional third-party dependency missing (redis, lightning, ...)
    for rel in ("bench.py", "__graft_entry__.py"):
        spec = importlib.util.spec_from_file_location("_static_" + rel.replace(".", "_"), os.path.join(REPO, rel))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        yield mod


def test_every_global_name_resolves():
    import importlib.util  # noqa: F401

    problems = []
    for mod in _module_sources():
        path = getattr(mod, "__file__", None)
        if not path or not path.endswith(".py"):
            continue
        with open(path) as f:
            top = compile(f.read(), path, "exec")
        defined = set(vars(mod)) | set(dir(builtins))
        # names bound anywhere at module level (including under ``if TYPE_CHECKING`` / try-except) count as defined
        defined |= set(top.co_names) & {i.argval for i in dis.get_instructions(top) if i.opname in ("STORE_NAME", "STORE_GLOBAL", "IMPORT_NAME", "IMPORT_FROM")}
        for code in _code_objects(top):
            stores = {i.argval for i in dis.get_instructions(code) if i.opname == "STORE_GLOBAL"}
            for ins in dis.get_instructions(code):
                if ins.opname == "LOAD_GLOBAL" and ins.argval not in defined and ins.argval not in stores:
                    problems.append(f"{os.path.relpath(path, REPO)}:{ins.positions.lineno if ins.positions else '?'} {code.co_name}: {ins.argval}")
    assert not problems, "\n".join(sorted(set(problems)))


def test_scripts_examples_and_benchmarks_resolve_their_globals():
    """Same guard without importing (these files parse arguments / start training at import time under ``__main__``): a name
    loaded as a global inside a function must be bound somewhere at module level."""
    import glob

    files = []
    for pat in ("examples/**/*.py", "benchmarks/*.py", "scripts/*.py", "tests/*.py", "setup.py"):
        files += glob.glob(os.path.join(REPO, pat), recursive=True)
    problems = []
    for path in sorted(files):
        with open(path) as f:
            top = compile(f.read(), path, "exec")
        defined = set(dir(builtins)) | {"__file__", "__name__", "__doc__"}
        defined |= {i.argval for i in dis.get_instructions(top) if i.opname in ("STORE_NAME", "STORE_GLOBAL")}
        for code in _code_objects(top):
            defined |= {i.argval for i in dis.get_instructions(code) if i.opname == "STORE_GLOBAL"}
        for code in _code_objects(top):
            for ins in dis.get_instructions(code):
                if ins.opname == "LOAD_GLOBAL" and ins.argval not in defined:
                    problems.append(f"{os.path.relpath(path, REPO)}:{ins.positions.lineno if ins.positions else '?'} {code.co_name}: {ins.argval}")
    assert not problems, "\n".join(sorted(set(problems)))


def test_native_call_sites_match_the_bindings():
    """Every ``native().f(...)`` / ``C.Cls(...)`` call in the repository is checked against the pybind11 signature in the built
    module: attribute exists, positional count fits, keywords exist, required parameters are given.  Most of these calls sit
    in CUDA-only branches that no CPU test executes."""
    import ast
    import re

    from bagua_b200.core import native

    C = native()

    def sigs(obj):
        doc=(obj.__doc__ or "")
        out=[]
        for line in doc.splitlines():
            m=re.match(r'\s*(?:\d+\.\s*)?(\w+)\((.*)\)\s*(->.*)?$', line)
            if m:
                params=[p.strip() for p in split_top(m.group(2)) if p.strip()]
                out.append(params)
        return out

    def split_top(s):
        parts=[];depth=0;cur=''
        for ch in s:
            if ch in '[(': depth+=1
            if ch in '])': depth-=1
            if ch==',' and depth==0: parts.append(cur);cur=''
            else: cur+=ch
        parts.append(cur); return parts

    def check(params, npos, kws, is_method):
        if is_method: params=params[1:]
        names=[p.split(':')[0].strip() for p in params]
        required=[n for n,p in zip(names,params) if '=' not in p]
        if npos>len(names): return f"too many positional ({npos} > {len(names)})"
        for k in kws:
            if k not in names: return f"unknown kw {k}"
        given=set(names[:npos])|set(kws)
        miss=[r for r in required if r not in given]
        if miss: return f"missing {miss}"
        return None

    problems=[]
    checked=0
    for root,_,files in os.walk(REPO):
        if any(x in root for x in ('/.git','baseline','gpurun_out','/build')): continue
        for f in files:
            if not f.endswith('.py'): continue
            p=os.path.join(root,f); src=open(p).read()
            try: tree=ast.parse(src)
            except SyntaxError: continue
            has_C=bool(re.search(r'\bC\s*=\s*_?native\(\)', src))
            for node in ast.walk(tree):
                if not isinstance(node, ast.Call) or not isinstance(node.func, ast.Attribute): continue
                fn=node.func; base=fn.value; chain=[fn.attr]
                # native().X(...) or C.X(...)  or C.Cls.static(...)
                target=None
                def is_native(b):
                    return (isinstance(b, ast.Call) and isinstance(b.func, ast.Name) and b.func.id in ('native','_native')) or (has_C and isinstance(b, ast.Name) and b.id=='C')
                if is_native(base): target=getattr(C, fn.attr, None); nm=fn.attr
                elif isinstance(base, ast.Attribute) and is_native(base.value):
                    cls=getattr(C, base.attr, None); target=getattr(cls, fn.attr, None) if cls else None; nm=base.attr+'.'+fn.attr
                else: continue
                if target is None:
                    problems.append((os.path.relpath(p,REPO), node.lineno, nm, 'no such attribute')); continue
                if any(isinstance(a, ast.Starred) for a in node.args) or any(k.arg is None for k in node.keywords): continue
                npos=len(node.args); kws=[k.arg for k in node.keywords]
                if isinstance(target, type):
                    ss=sigs(target.__init__); is_method=True
                else:
                    ss=sigs(target); is_method=False
                if not ss: continue
                checked+=1
                errs=[check(s, npos, kws, is_method) for s in ss]
                if all(errs):
                    problems.append((os.path.relpath(p,REPO), node.lineno, nm, errs[0]))
    assert checked > 60, checked
    assert not problems, "\n".join(str(p) for p in sorted(set(problems)))


def test_public_api_is_documented_and_the_index_is_current():
    """Every public class / function of the package carries a docstring, and docs/api.md is what scripts/gen_api_docs.py
    generates from them now (regenerate with ``python scripts/gen_api_docs.py > docs/api.md``)."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("gen_api_docs", os.path.join(REPO, "scripts", "gen_api_docs.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    _, missing = gen.collect()
    assert not missing, "public symbols without a docstring: " + ", ".join(missing)
    with open(os.path.join(REPO, "docs", "api.md")) as f:
        assert f.read() == gen.render(), "docs/api.md is stale: python scripts/gen_api_docs.py > docs/api.md"
