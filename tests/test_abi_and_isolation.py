"""CPU: the C-ABI library loads and exports exactly what include/annb.h declares; the product package
never touches oracle/; and the Python prototypes table matches the header."""
import ast
import os
import re
import subprocess

from conftest import ROOT


def header_functions():
    src = open(os.path.join(ROOT, 'include', 'annb.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'ANNB_API\s+(?:const\s+char\s*\*|int)\s*(annb_\w+)\s*\(', src)))


def test_header_cites_reference_and_declares_functions():
    src = open(os.path.join(ROOT, 'include', 'annb.h')).read()
    fns = header_functions()
    assert len(fns) >= 30
    for cite in ['hnsw_bindings.cpp:303', ':393-516', 'pq_bindings.pyx:149', 'pq_bindings.pyx:52',
                 'hnswalg.h:708', 'container.py:130', 'pq.py:158']:
        assert cite in src, cite


def test_library_exports_every_declared_symbol():
    from annlite_b200 import _lib
    lib = _lib.load()
    fns = header_functions()
    for f in fns:
        assert hasattr(lib, f), f'{f} declared in include/annb.h but not exported'
    out = subprocess.check_output(['nm', '-D', '--defined-only', _lib.LIB_PATH], text=True)
    exported = sorted(l.split()[-1] for l in out.splitlines() if ' T ' in l)
    assert exported == fns, set(exported) ^ set(fns)        # nothing else leaks out of the .so
    assert sorted(_lib.PROTOTYPES) == fns
    assert lib.annb_version() == 100


def test_no_gpu_means_loud_failure():
    from annlite_b200 import _lib
    from annlite_b200.engine import Engine
    import pytest
    if _lib.load().annb_device_count() > 0:
        pytest.skip('GPU present')
    with pytest.raises(_lib.AnnbError) as ei:
        Engine(128, 8, 256)
    assert ei.value.code == _lib.ENODEVICE and 'no CPU fallback' in str(ei.value)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, 'annlite_b200')
    bad = []
    for dp, _, files in os.walk(pkg):
        for f in files:
            p = os.path.join(dp, f)
            if f.endswith('.py'):
                tree = ast.parse(open(p).read())
                for node in ast.walk(tree):
                    names = []
                    if isinstance(node, ast.Import):
                        names = [a.name for a in node.names]
                    elif isinstance(node, ast.ImportFrom) and node.module:
                        names = [node.module]
                    bad += [(p, n) for n in names if n.split('.')[0] == 'oracle']
            elif f.endswith(('.cu', '.cpp', '.h', '.cuh', '.sh')):
                txt = open(p).read()
                if re.search(r'oracle/|liborc|pq_oracle', txt):
                    bad.append((p, 'mentions oracle'))
    assert not bad, bad


def test_library_does_not_link_the_oracle():
    from annlite_b200 import _lib
    out = subprocess.check_output(['ldd', _lib.LIB_PATH], text=True)
    assert 'liborc' not in out
