"""CPU-side checks of the drop-in boundary: libojdf.so loads without a GPU, exports every
symbol include/ojdf.h declares, and rejects bad arguments before touching CUDA."""
import ctypes as C
import os
import re

import pytest
import torch

from online_joint_depthfusion_and_semantic_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def L():
    build.build()
    return _lib.lib()


def declared_symbols():
    src = open(os.path.join(ROOT, 'include', 'ojdf.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(ojdf_[a-z_0-9]+)\s*\(', src)))


def test_header_symbols_all_exported(L):
    names = declared_symbols()
    assert len(names) >= 9
    raw = C.CDLL(_lib.SO_PATH)
    for n in names:
        assert hasattr(raw, n), n
    assert set(names) == set(_lib.EXPORTS)


def test_version_and_error_strings(L):
    assert L.ojdf_version() == 100
    assert b'bad argument' in L.ojdf_error_string(-1)
    assert b'workspace' in L.ojdf_error_string(-2)
    assert L.ojdf_error_string(0) == b'success'


def test_workspace_bytes_monotonic(L):
    prev = 0
    for e in (1, 8, 2048, 2049, 19200 * 56, 76800 * 56, 307200 * 56):
        b = L.ojdf_integrate_workspace_bytes(e)
        assert b >= prev and b > 0
        prev = b
    assert L.ojdf_integrate_workspace_bytes(0) == 0
    assert L.ojdf_integrate_workspace_bytes(2 ** 31) == 0      # does not fit the 32-bit entry index


def test_bad_arguments_rejected_without_cuda(L):
    assert L.ojdf_unproject(None, 4, 4, None, None, None, None) == -1
    assert L.ojdf_extract(None, None, 4, 4, None, None, None, 0.1, None, None, 8, 8, 8, 9,
                          None, None, None, None, None, None, None, None) == -1
    assert L.ojdf_integrate(None, None, None, 16, 9, 7, 0.1, None, None, 8, 8, 8, None, None, None, None, 0,
                            None, 0, None) == -1
    assert L.ojdf_integrate_updates(None, None, None, 16, None, None, 8, 8, 8, None, None, None, None, 0,
                                    None, 0, None) == -1
    assert L.ojdf_integrate_workspace_init(None, 0, None) == -2
    # even P is not a valid ray length (centre sample must exist)
    dummy = (C.c_char * 64)()
    p = C.addressof(dummy)
    assert L.ojdf_extract(p, None, 4, 4, p, p, p, 0.1, p, p, 8, 8, 8, 8, p, p, p, p, None, None, None, None) == -1


def test_host_modules_refuse_cpu_tensors():
    from online_joint_depthfusion_and_semantic_b200.config import fusion_config
    from online_joint_depthfusion_and_semantic_b200.modules import Extractor, Integrator
    cfg = fusion_config(8, 8, device='cpu')
    ex = Extractor(cfg)
    vol = torch.zeros(4, 4, 4, dtype=torch.float16)
    with pytest.raises(_lib.OjdfError):
        ex.forward(torch.ones(1, 8, 8), torch.eye(4)[None], torch.eye(3)[None].double(), vol, vol,
                   torch.zeros(3, dtype=torch.float64), 0.1)
    integ = Integrator(cfg)
    with pytest.raises(_lib.OjdfError):
        integ.forward({}, vol, vol, vol, vol.to(torch.uint8))
