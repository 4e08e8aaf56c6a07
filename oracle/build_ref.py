#!/usr/bin/env python
"""Build oracle/_ref: the UNMODIFIED reference, byte-compiled from the sources where they lie -- TEST / BENCH INFRASTRUCTURE.

The reference is pure Python, so "compiling its path from its own few source files" (what a C reference gets from gcc)
is byte compilation: every module the update loop needs (src/bss/{ilrma,iva,mnmf}.py, src/algorithm/nmf.py and whatever
they import from the same tree) is compiled from /root/reference/src straight into oracle/_ref/<package>/<module>.rbc
(a marshalled code object; not `.pyc`, which repository snapshots commonly filter out as cache files).
No source file is copied; oracle/_ref/ is git-ignored but travels to the GPU box with the repository snapshot like any
other built artefact, so `bench.py` can time the reference's own classes there (`cpu_baseline.kind = "reference"`).
Marshalled code loads on the same CPython minor version only (build container and GPU box share one image);
`load()` returns None when the files are absent or stale and the callers fall back to the oracle port (kind "port").

    python oracle/build_ref.py          # in the build container (needs /root/reference)
"""
import ast
import importlib
import os
import marshal
import types
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = '/root/reference/src'
OUT = os.path.join(HERE, '_ref')
ENTRY = ['bss.ilrma', 'bss.iva', 'bss.mnmf', 'algorithm.nmf', 'algorithm.projection_back', 'utils.utils_linalg', 'transform.stft']


def _local_imports(path):
    with open(path) as fh:
        tree = ast.parse(fh.read(), path)
    found = set()
    for node in ast.walk(tree):
        names = []
        if isinstance(node, ast.Import):
            names = [a.name for a in node.names]
        elif isinstance(node, ast.ImportFrom) and node.module and node.level == 0:
            names = [node.module]
        for name in names:
            if os.path.exists(os.path.join(REF_SRC, *name.split('.')) + '.py'):
                found.add(name)
    return found


def build(verbose=True):
    if not os.path.isdir(REF_SRC):
        return False
    todo, done = list(ENTRY), set()
    while todo:
        name = todo.pop()
        if name in done:
            continue
        done.add(name)
        src = os.path.join(REF_SRC, *name.split('.')) + '.py'
        dst = os.path.join(OUT, *name.split('.')) + '.rbc'
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        with open(src) as fh:
            code = compile(fh.read(), 'reference/src/' + name.replace('.', '/') + '.py', 'exec')
        with open(dst, 'wb') as fh:
            marshal.dump(code, fh)
        todo.extend(_local_imports(src) - done)
    with open(os.path.join(OUT, 'BUILT_FROM'), 'w') as fh:
        fh.write("{} (python {}.{}): {}\n".format(REF_SRC, sys.version_info[0], sys.version_info[1], ' '.join(sorted(done))))
    if verbose:
        print("oracle/_ref: byte-compiled {} reference modules".format(len(done)))
    return True


class _RefFinder:
    """Import hook for the byte-compiled reference: `bss.ilrma` -> oracle/_ref/bss/ilrma.rbc; `bss`, `algorithm`, ... are
    plain namespace modules (the reference tree has no __init__.py either)."""

    def find_spec(self, name, path=None, target=None):
        from importlib.machinery import ModuleSpec
        base = os.path.join(OUT, *name.split('.'))
        if os.path.isfile(base + '.rbc'):
            return ModuleSpec(name, self, origin=base + '.rbc')
        if os.path.isdir(base):
            return ModuleSpec(name, self, origin=base, is_package=True)
        return None

    def create_module(self, spec):
        return None

    def exec_module(self, module):
        origin = module.__spec__.origin
        if os.path.isdir(origin):
            module.__path__ = [origin]
            return
        with open(origin, 'rb') as fh:
            code = marshal.load(fh)
        if not isinstance(code, types.CodeType):
            raise ImportError("stale byte code: " + origin)
        exec(code, module.__dict__)


def load():
    """Import the byte-compiled reference behind the NumPy-1.x `linalg.solve` shim (oracle/pin/np1shim.py: without it the
    reference raises under NumPy >= 2).  Returns a dict of its classes, or None when oracle/_ref is absent / unusable."""
    if not os.path.exists(os.path.join(OUT, 'bss', 'ilrma.rbc')):
        return None
    sys.path.insert(0, os.path.join(HERE, 'pin'))
    import np1shim  # noqa: F401
    if not any(isinstance(f, _RefFinder) for f in sys.meta_path):
        sys.meta_path.insert(0, _RefFinder())
    try:
        ilrma = importlib.import_module('bss.ilrma')
        iva = importlib.import_module('bss.iva')
        mnmf = importlib.import_module('bss.mnmf')
        nmf = importlib.import_module('algorithm.nmf')
    except Exception:   # stale byte code (another CPython) or a missing dependency
        return None
    return {'GaussILRMA': ilrma.GaussILRMA, 'AuxLaplaceIVA': iva.AuxLaplaceIVA, 'FastMultichannelISNMF': mnmf.FastMultichannelISNMF,
            'EUCNMF': nmf.EUCNMF}


if __name__ == '__main__':
    sys.exit(0 if build() else 1)
