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


def test_conv_tc_weight_packing_layout(L):
    """ojdf_conv_tc_pack_weights against a numpy restatement of its documented layout:
    [group][tap][K chunk of 32][hi|lo][npad][32], 16-byte chunk c of row r stored at chunk c ^ (r & 7),
    hi = w with the low 13 mantissa bits cleared, lo = w - hi."""
    import numpy as np
    rng = np.random.default_rng(7)
    for cin, cout, taps, npad_req in ((19, 19, 9, 0), (114, 95, 1, 0), (70, 300, 1, 0), (40, 256, 9, 32)):
        npad, groups = C.c_int(), C.c_int()
        assert L.ojdf_conv_tc_layout(cout, npad_req, C.byref(npad), C.byref(groups)) == 0
        npad, groups = npad.value, groups.value
        assert npad % 16 == 0 and npad <= 128 and npad * groups >= cout
        w = rng.standard_normal((cout, cin, taps)).astype(np.float32)
        n = L.ojdf_conv_tc_weight_floats(cin, cout, taps, npad_req)
        nkc = (cin + 31) // 32
        assert n == groups * taps * nkc * 2 * npad * 32
        packed = np.full(n, np.nan, np.float32)
        assert L.ojdf_conv_tc_pack_weights(w.ctypes.data, cin, cout, taps, npad_req, packed.ctypes.data) == 0
        img = packed.reshape(groups, taps, nkc, 2, npad, 8, 4)
        hi = (w.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)
        lo = w - hi
        assert np.array_equal(hi + lo, w)
        for _ in range(200):
            co, ci, t = int(rng.integers(cout)), int(rng.integers(cin)), int(rng.integers(taps))
            g, r, kc, k = co // npad, co % npad, ci // 32, ci % 32
            chunk = (k >> 2) ^ (r & 7)
            assert img[g, t, kc, 0, r, chunk, k & 3] == hi[co, ci, t]
            assert img[g, t, kc, 1, r, chunk, k & 3] == lo[co, ci, t]
        assert np.count_nonzero(packed) <= 2 * w.size and not np.isnan(packed).any()
    assert L.ojdf_conv_tc_batched(None, 1, 32, 32, 8, 16, 1, 0, 0.0, 1.0, 0, 0, None, 0, None) == -1


def test_conv_problem_struct_matches_header(tmp_path):
    """The ctypes mirror of ojdf_conv_problem (modules/fusion_engine.py) has the size and field offsets the C
    compiler gives include/ojdf.h."""
    import subprocess
    from online_joint_depthfusion_and_semantic_b200.modules.fusion_engine import ConvProblem
    fields = [f[0] for f in ConvProblem._fields_]
    src = tmp_path / 'layout.c'
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "ojdf.h"\nint main(void) {\n'
                   '  printf("%zu", sizeof(ojdf_conv_problem));\n' +
                   ''.join('  printf(" %%zu", offsetof(ojdf_conv_problem, %s));\n' % f for f in fields) +
                   '  return 0;\n}\n')
    exe = tmp_path / 'layout'
    subprocess.check_call(['gcc', '-I', os.path.join(ROOT, 'include'), str(src), '-o', str(exe)])
    out = [int(v) for v in subprocess.check_output([str(exe)]).split()]
    assert out[0] == C.sizeof(ConvProblem)
    assert out[1:] == [getattr(ConvProblem, f).offset for f in fields]


@pytest.mark.parametrize('ctype,cname', [('ChainInput', 'ojdf_chain_input'), ('ChainStep', 'ojdf_chain_step'),
                                         ('PoolProblem', 'ojdf_pool_problem')])
def test_other_problem_structs_match_header(tmp_path, ctype, cname):
    """The ctypes mirrors of the chain / pool structs (modules/fusion_engine.py) against the C compiler's layout of include/ojdf.h."""
    import subprocess
    from online_joint_depthfusion_and_semantic_b200.modules import fusion_engine
    T = getattr(fusion_engine, ctype)
    fields = [f[0] for f in T._fields_]
    src = tmp_path / 'layout.c'
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "ojdf.h"\nint main(void) {\n'
                   '  printf("%%zu", sizeof(%s));\n' % cname +
                   ''.join('  printf(" %%zu", offsetof(%s, %s));\n' % (cname, f) for f in fields) +
                   '  return 0;\n}\n')
    exe = tmp_path / 'layout'
    subprocess.check_call(['gcc', '-I', os.path.join(ROOT, 'include'), str(src), '-o', str(exe)])
    out = [int(v) for v in subprocess.check_output([str(exe)]).split()]
    assert out[0] == C.sizeof(T)
    assert out[1:] == [getattr(T, f).offset for f in fields]


def test_chain_argument_checks_need_no_gpu():
    """ojdf_conv_chain rejects malformed chains before touching the device."""
    from online_joint_depthfusion_and_semantic_b200.modules.fusion_engine import ChainInput, ChainStep
    L = _lib.lib()
    assert L.ojdf_conv_chain(None, 1, None, 1, 1, 8, 16, None, None, 32, 1.0, 0, None) == -1
    ia, sa = (ChainInput * 1)(), (ChainStep * 13)()
    op, oc = (C.c_void_p * 1)(), (C.c_int * 1)()
    assert L.ojdf_conv_chain(ia, 1, sa, 13, 1, 8, 16, op, oc, 32, 1.0, 0, None) == -1           # more than 12 steps
    assert L.ojdf_conv_chain(ia, 1, sa, 1, 3, 8, 16, op, oc, 32, 1.0, 0, None) == -1            # more than 2 problems


def test_frame_stream_needs_cuda():
    from online_joint_depthfusion_and_semantic_b200.stream import FrameStream
    with pytest.raises(ValueError):
        FrameStream(None, None, 'cpu')
