#!/usr/bin/env python
"""Build oracle/_ref: the UNMODIFIED reference, byte-compiled from the sources where they lie -- TEST / BENCH INFRASTRUCTURE.

The reference is pure Python, so "compiling its path from its own few source files" (what a C reference gets from gcc)
is `py_compile`: every module the update loop needs (src/bss/{ilrma,iva,mnmf}.py, src/algorithm/nmf.py and whatever
they import from the same tree) is compiled from /root/reference/src straight into oracle/_ref/<package>/<module>.pyc.
No source file is copied; oracle/_ref/ is git-ignored but travels to the GPU box with the repository snapshot like any
other built artefact, so `bench.py` can time the reference's own classes there (`cpu_baseline.kind = "reference"`).
Sourceless .pyc files import on the same CPython minor version only (build container and GPU box share one image);
`load()` returns None when they are absent or stale and the callers fall back to the oracle port (kind "port").

    python oracle/build_ref.py          # in the build container (needs /root/reference)
"""
import ast
import importlib
import os
import py_compile
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
        dst = os.path.join(OUT, *name.split('.')) + '.pyc'
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        py_compile.compile(src, cfile=dst, dfile='reference/src/' + name.replace('.', '/') + '.py', doraise=True)
        todo.extend(_local_imports(src) - done)
    with open(os.path.join(OUT, 'BUILT_FROM'), 'w') as fh:
        fh.write("{} (python {}.{}): {}\n".format(REF_SRC, sys.version_info[0], sys.version_info[1], ' '.join(sorted(done))))
    if verbose:
        print("oracle/_ref: byte-compiled {} reference modules".format(len(done)))
    return True


def load():
    """Import the byte-compiled reference behind the NumPy-1.x `linalg.solve` shim (oracle/pin/np1shim.py: without it the
    reference raises under NumPy >= 2).  Returns a dict of its classes, or None when oracle/_ref is absent / unusable."""
    if not os.path.exists(os.path.join(OUT, 'bss', 'ilrma.pyc')):
        return None
    sys.path.insert(0, os.path.join(HERE, 'pin'))
    import np1shim  # noqa: F401
    if OUT not in sys.path:
        sys.path.append(OUT)
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
