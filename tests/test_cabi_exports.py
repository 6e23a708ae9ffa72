"""CPU-side checks of the boundary: the C-ABI library loads and exports every symbol include/gpmc.h
declares; the product path refuses to run without a CUDA device (no CPU fallback)."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, 'include', 'gpmc.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(gpmc_[a-z0-9_]+)\s*\(', text)))


def test_header_symbols_are_exported(gp):
    lib = gp._lib.load()
    names = _declared_symbols()
    assert len(names) >= 10
    for name in names:
        assert hasattr(lib, name), 'libgpmc.so does not export %s' % name
        assert name in gp._lib.SIGNATURES, 'no ctypes prototype for %s' % name
    assert sorted(gp._lib.SIGNATURES) == names
    assert lib.gpmc_version() == 100


def test_header_is_plain_c_and_links_against_the_library(gp, tmp_path):
    """include/gpmc.h must be consumable by a C compiler (no C++-isms, no torch types), and a C program that calls the
    host-only entry points must link against libgpmc.so and run without a GPU."""
    import shutil
    import subprocess
    gcc = shutil.which('gcc')
    if gcc is None:
        pytest.skip('no gcc')
    lib_dir = os.path.join(ROOT, 'gaussianprocess-mcmc_b200')
    src = tmp_path / 'abi.c'
    src.write_text(
        '#include <stdio.h>\n#include "gpmc.h"\n'
        'int main(void) {\n'
        '  size_t w = gpmc_workspace_bytes(GPMC_OP_LOGLIK, 4096, 1, 8);\n'
        '  printf("%d %d %zu %zu\\n", gpmc_version(), gpmc_panel_width(), w, gpmc_sds_workspace_bytes(512, 3, 4));\n'
        '  return gpmc_set_tuning(99, 0) == GPMC_EINVAL ? 0 : 1;\n}\n')
    exe = tmp_path / 'abi'
    subprocess.run([gcc, '-std=c99', '-Wall', '-Werror', '-pedantic', '-I', os.path.join(ROOT, 'include'), str(src), '-o', str(exe),
                    '-L', lib_dir, '-lgpmc', '-Wl,-rpath,' + lib_dir], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    assert out[0] == '100' and out[1] == '128'
    assert int(out[2]) > 8 * 4096 * 4096 and int(out[3]) > 0


def test_workspace_query_is_pure_host(gp):
    lib = gp._lib.load()
    one = lib.gpmc_workspace_bytes(gp._lib.OP_LOGLIK, 4096, 1, 1)
    assert one >= 4096 * 4096 * 8 + 128 * 128 * 8
    assert lib.gpmc_workspace_bytes(gp._lib.OP_LOGLIK, 4096, 1, 1024) <= (48 << 30) + (1 << 20)
    assert lib.gpmc_workspace_bytes(gp._lib.OP_LOGLIK, 0, 1, 1) == 0


def test_no_cpu_fallback(gp):
    import torch
    if torch.cuda.is_available():
        pytest.skip('a GPU is present')
    with pytest.raises(gp.GpmcError):
        gp.ops.loglik_host(np.zeros((4, 1)), np.zeros((1, 4)), np.ones((1, 3)))
    with pytest.raises(gp.GpmcError):
        gp.ops.cov_assemble(np.zeros((4, 1)), np.ones((1, 3)))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, 'gaussianprocess-mcmc_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                src = open(os.path.join(dirpath, f)).read()
                assert 'import oracle' not in src and 'from oracle' not in src, f
