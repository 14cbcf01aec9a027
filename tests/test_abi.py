"""CPU: the C-ABI library builds, loads, and exports every symbol include/osd_b200.h declares
(no compute calls: there is no GPU here and no CPU fallback in the product)."""
import ctypes
import os
import re

import pytest


def test_library_loads_and_exports_header_symbols():
    from osu_dreamer_b200 import lib
    assert os.path.exists(lib.LIB_PATH), 'run `python -c "import __graft_entry__ as g; g.build()"` first'
    h = open(lib.HEADER_PATH).read()
    names = re.findall(r'OSD_API\s+[\w\s\*]+?\b(osd_\w+)\s*\(', h)
    assert len(names) >= 18 and len(set(names)) == len(names)
    dll = ctypes.CDLL(lib.LIB_PATH)
    for n in names:
        assert hasattr(dll, n), f'{n} declared in osd_b200.h but not exported'
    dll.osd_abi_version.restype = ctypes.c_int
    assert dll.osd_abi_version() == 1
    # size queries are pure host functions
    dll.osd_packed_bytes.restype = ctypes.c_size_t
    assert dll.osd_packed_bytes(0) > 60e6
    dll.osd_workspace_bytes.restype = ctypes.c_size_t
    infer = dll.osd_workspace_bytes(16, 8192, 16, 0, 0)
    train = dll.osd_workspace_bytes(16, 8192, 16, 0, 1)
    assert infer < 8e9 and 25e9 < train < 45e9  # ~33 KB/token/layer of saved activations


def test_no_cpu_fallback():
    import torch
    from osu_dreamer_b200 import lib
    from osu_dreamer_b200.denoiser import DiffusionModel, default_args
    m = DiffusionModel(6, 128, 32, default_args())
    with pytest.raises(lib.OsdError):
        m(torch.randn(1, 128, 64), torch.randn(1, 32), torch.randn(1, 6, 64))
    with pytest.raises(lib.OsdError):
        lib.gemm(torch.zeros(128, 64, dtype=torch.bfloat16), torch.zeros(128, 64, dtype=torch.bfloat16),
                 torch.zeros(128, 128))


def test_product_never_imports_oracle():
    """only tests/, __graft_entry__.smoke() and bench.py may touch oracle/ (and only as checker/baseline)."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, 'osu-dreamer_b200')
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                src = open(os.path.join(dp, f)).read()
                assert 'oracle' not in src.replace('# oracle', ''), f'{f} mentions oracle'


def test_header_is_plain_c99(tmp_path):
    """the drop-in boundary is a C ABI: include/osd_b200.h must compile as strict C99 (no C++ / torch types)"""
    import shutil
    import subprocess
    if shutil.which('gcc') is None:
        pytest.skip('gcc not available')
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / 'hdr.c'
    src.write_text('#include "osd_b200.h"\nint main(void) { return osd_abi_version() == 0; }\n')
    r = subprocess.run(['gcc', '-std=c99', '-Wall', '-Wextra', '-pedantic', '-Werror', '-fsyntax-only', '-I', os.path.join(root, 'include'), str(src)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
