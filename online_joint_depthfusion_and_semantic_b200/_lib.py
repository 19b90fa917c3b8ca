"""ctypes binding of csrc/libojdf.so -- the only way the host side reaches the GPU kernels.

There is deliberately no fallback: if the library is missing or a call fails, this module
raises.  Signatures mirror include/ojdf.h one to one.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, 'csrc', 'libojdf.so')
_lib = None

_vp, _i, _i64, _f, _d, _sz = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_double, C.c_size_t

_SIGNATURES = {
    'ojdf_version': (C.c_int, []),
    'ojdf_error_string': (C.c_char_p, [_i]),
    'ojdf_launch_count': (C.c_uint64, []),
    'ojdf_unproject': (_i, [_vp, _i, _i, _vp, _vp, _vp, _vp]),
    'ojdf_extract': (_i, [_vp, _vp, _i, _i, _vp, _vp, _vp, _d, _vp, _vp, _i, _i, _i, _i,
                          _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    'ojdf_rays': (_i, [_vp, _vp, _i, _i, _vp, _vp, _vp, _d, _vp, _vp, _vp]),
    'ojdf_gather': (_i, [_vp, _i, _i, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp]),
    'ojdf_integrate_workspace_bytes': (_sz, [_i64]),
    'ojdf_integrate_workspace_init': (_i, [_vp, _sz, _vp]),
    'ojdf_integrate_workspace_idle_bytes': (_sz, [_sz]),
    'ojdf_integrate': (_i, [_vp, _vp, _vp, _i64, _i, _i, _f, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _i,
                            _vp, _sz, _vp]),
    'ojdf_integrate_plan': (_i, [_vp, _vp, _i64, _i, _i, _i, _i, _i, _vp, _sz, _vp]),
    'ojdf_integrate_apply': (_i, [_vp, _i64, _i, _i, _f, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _i, _vp, _sz, _vp]),
    'ojdf_integrate_updates': (_i, [_vp, _vp, _vp, _i64, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _i,
                                    _vp, _sz, _vp]),
    'ojdf_conv_tc_layout': (_i, [_i, _i, _vp, _vp]),
    'ojdf_conv_tc_weight_floats': (_sz, [_i, _i, _i, _i]),
    'ojdf_conv_tc_pack_weights': (_i, [_vp, _i, _i, _i, _i, _vp]),
    'ojdf_conv_tc_batched': (_i, [_vp, _i, _i, _i, _i, _i, _i, _i, _f, _f, _i, _i, _vp, _sz, _vp]),
    'ojdf_avgpool3_nhwc': (_i, [_vp, _i, _i, _i, _i, _vp, _i, _vp]),
    'ojdf_avgpool3_batched': (_i, [_vp, _i, _i, _i, _i, _i, _vp]),
    'ojdf_vortex_bias': (_i, [_vp, _i, _i, _i, _vp, _vp, _vp, _i, _vp, _vp, _vp, _i, _vp, _i, _vp, _vp]),
    'ojdf_gap_bias': (_i, [_vp, _i, _i, _i, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _i, _vp, _i, _vp, _vp]),
    'ojdf_conv_chain': (_i, [_vp, _i, _vp, _i, _i, _i, _i, _vp, _vp, _i, _f, _i, _vp]),
    'ojdf_adapnet_skip_join': (_i, [_vp, _i, _i, _i, _vp, _vp, _i, _vp, _i, _vp, _i, _vp, _i, _vp]),
    'ojdf_nchw_to_nhwc': (_i, [_vp, _i, _i, _vp, _i, _i, _vp]),
    'ojdf_nhwc_to_nchw': (_i, [_vp, _i, _i, _i, _i, _vp, _vp]),
    'ojdf_adapnet_stem': (_i, [_vp, _i, _i, _i, _vp]),
    'ojdf_softmax_max': (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    'ojdf_pack_fusion_input': (_i, [_vp, _vp, _vp, _vp, _i, _i, _vp, _vp, _i, _vp]),
}

EXPORTS = tuple(_SIGNATURES)


class OjdfError(RuntimeError):
    pass


def lib():
    """Load libojdf.so (once).  Raises if it has not been built -- never falls back."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise OjdfError('%s is missing: build it with `python -m online_joint_depthfusion_and_semantic_b200.build` '
                            '(there is no CPU or PyTorch fallback for the fusion hot path)' % SO_PATH)
        l = C.CDLL(SO_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype, fn.argtypes = res, args
        _lib = l
    return _lib


def check(code):
    if code != 0:
        raise OjdfError('libojdf: %s (code %d)' % (lib().ojdf_error_string(code).decode(), code))


def ptr(t):
    """Device (or host) pointer of a contiguous tensor, None -> NULL."""
    if t is None:
        return None
    assert t.is_contiguous(), 'ojdf: tensor must be contiguous'
    return t.data_ptr()


def stream_ptr(device=None):
    return torch.cuda.current_stream(device).cuda_stream


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise OjdfError('ojdf: the fusion hot path runs on CUDA tensors only (got a %s tensor); '
                            'there is no CPU fallback' % t.device)


_REPLAYED = 0


def note_replayed(n):
    """A CUDA-graph replay re-launched `n` libojdf kernels that the C-side counter only saw at capture time."""
    global _REPLAYED
    _REPLAYED += int(n)


def launch_count():
    """libojdf kernels launched by this process: direct launches (C-side counter) + graph replays."""
    return int(lib().ojdf_launch_count()) + _REPLAYED


class KernelTimers:
    """Optional CUDA-event brackets around the libojdf calls (bench.py's roofline leg).
    Disabled (None) by default: no events are created on the normal path."""

    def __init__(self):
        self.events = {}

    def bracket(self, name, device):
        return _Bracket(self, name, device)

    def mean_ms(self, name):
        ev = self.events.get(name, [])
        return sum(a.elapsed_time(b) for a, b in ev) / len(ev) if ev else None

    def count(self, name):
        return len(self.events.get(name, []))


class _Bracket:
    def __init__(self, timers, name, device):
        self.t, self.name, self.device = timers, name, device

    def __enter__(self):
        self.a = torch.cuda.Event(enable_timing=True)
        self.b = torch.cuda.Event(enable_timing=True)
        self.a.record(torch.cuda.current_stream(self.device))

    def __exit__(self, *exc):
        self.b.record(torch.cuda.current_stream(self.device))
        self.t.events.setdefault(self.name, []).append((self.a, self.b))
        return False


class _NoBracket:
    def __enter__(self):
        return None

    def __exit__(self, *exc):
        return False


TIMERS = None
_NO = _NoBracket()


def timed(name, device):
    return TIMERS.bracket(name, device) if TIMERS is not None else _NO
