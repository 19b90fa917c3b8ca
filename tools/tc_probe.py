"""Bring-up / timing probe for the tcgen05 tap GEMM (csrc/ojdf_conv_tc.cu): each case runs in its own process
(a trapped kernel must not take the others down), compares with an fp64 torch convolution and prints one line.
Usage on the GPU box:  python tools/tc_probe.py [--time] [--only PREFIX]
Case tuple: (name, H, W, cin, cout, taps, dil, in_stride, out_stride, out_coff, act, residual, n_problems, flags);
bits 16-23 of flags = npad_req."""
import ctypes as C
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = [
    # correctness (also in tests/test_gpu_conv_tc.py)
    ('1x1 one tile one chunk', 8, 16, 32, 32, 1, 1, 32, 32, 0, 0, False, 1, 0),
    ('1x1 cin 19 of stride 116', 8, 16, 19, 19, 1, 1, 116, 20, 0, 2, False, 1, 0),
    ('3x3 one tile', 8, 16, 32, 32, 9, 1, 32, 32, 0, 0, False, 1, 0),
    ('3x3 dense block', 48, 64, 95, 19, 9, 1, 116, 116, 95, 2, False, 2, 0),
    ('3x3 dil 3 ragged image', 37, 53, 19, 19, 9, 3, 20, 20, 0, 1, False, 4, 0),
    ('3x3 dil 27', 48, 64, 19, 19, 9, 27, 20, 20, 0, 1, False, 8, 0),
    ('1x1 -> 9 tanh', 48, 64, 19, 9, 1, 1, 116, 9, 0, 3, False, 1, 0),
    ('1x1 256 out (2 groups) residual sigmoid', 30, 40, 256, 256, 1, 1, 256, 256, 0, 4, True, 1, 0),
    ('3x3 30x40 64->64 dil 2 halo', 30, 40, 64, 64, 9, 2, 64, 64, 0, 1, False, 1, 0),
    ('3x3 dense block, per-tap boxes', 48, 64, 95, 19, 9, 1, 116, 116, 95, 2, False, 2, 4),
    ('3x3 dense block, MT=1', 48, 64, 95, 19, 9, 1, 116, 116, 95, 2, False, 2, 2),
    ('3x3 dense block, per-thread stores', 48, 64, 95, 19, 9, 1, 116, 116, 95, 2, False, 2, 8),
    # FusionNet layer shapes at 240x320 (flag 1: padded channel groups -> TMA-store epilogue)
    ('F: 3x3 114->19 x2 (dense block)', 240, 320, 114, 19, 9, 1, 120, 120, 100, 2, False, 2, 1),
    ('F: 3x3 19->19 x2', 240, 320, 19, 19, 9, 1, 20, 20, 0, 2, False, 2, 1),
    ('F: 1x1 456->114 x2 (vortex final)', 240, 320, 456, 114, 1, 1, 456, 116, 0, 0, False, 2, 1),
    ('F: 1x1 114->95 (pred)', 240, 320, 114, 95, 1, 1, 116, 116, 0, 2, False, 1, 1),
    ('F: 1x1 19->114 into 464 x8 (vortex branch out)', 240, 320, 19, 114, 1, 1, 20, 464, 116, 1, False, 8, 1),
    ('F: 3x3 19->19 x8 dilations 27/26', 240, 320, 19, 19, 9, 27, 20, 20, 0, 1, False, 8, 1),
    # AdapNet++ layer shapes
    ('A: 1x1 15x20 1024->256 x2 (split K)', 15, 20, 1024, 256, 1, 1, 1024, 512, 0, 1, False, 2, 0),
    ('A: 1x1 15x20 1024->256 x2, no split', 15, 20, 1024, 256, 1, 1, 1024, 512, 0, 1, False, 2, 4096),
    ('A: 3x3 15x20 512->256 x4 dil 4/3 (split K)', 15, 20, 512, 256, 9, 4, 512, 512, 0, 1, False, 4, 0),
    ('A: 1x1 15x20 512->2048 x2 residual', 15, 20, 512, 2048, 1, 1, 512, 2048, 0, 1, True, 2, 0),
    ('A: 3x3 60x80 280->256 (decoder stage 3)', 60, 80, 280, 256, 9, 1, 280, 256, 0, 1, False, 1, 0),
    ('A: 3x3 60x80 64->64 x2 (layer1)', 60, 80, 64, 64, 9, 1, 64, 64, 0, 1, False, 2, 0),
    ('A: 1x1 60x80 64->256 x2 residual', 60, 80, 64, 256, 1, 1, 64, 256, 0, 1, True, 2, 0),
    ('A: 3x3 30x40 48->4 relu (SSMA)', 30, 40, 48, 4, 9, 1, 48, 4, 0, 1, False, 1, 1),
    ('A: 1x1 30x40 128->512 x2 residual', 30, 40, 128, 512, 1, 1, 128, 512, 0, 1, True, 2, 0),
    ('A: 1x1 30x40 512->128 x2', 30, 40, 512, 128, 1, 1, 512, 128, 0, 1, False, 2, 0),
    ('A: 3x3 30x40 280->256 (decoder stage 2)', 30, 40, 280, 256, 9, 1, 280, 256, 0, 1, False, 1, 0),
    ('A: 3x3 30x40 256->128 x4 (layer3[0] conv2a/b)', 30, 40, 256, 128, 9, 2, 256, 256, 0, 1, False, 4, 0),
    ('A: 1x1 60x80 256->256 x2 (wide 1x1 at 60x80)', 60, 80, 256, 256, 1, 1, 256, 256, 0, 1, False, 2, 0),
    # role profile / timing experiments (flag 128 = per-role wait cycles of block 0; 16 = no MMAs, 32 = no split
    # work, 64 = 1xTF32, 1024 = busy-poll the A ring, 2048 = single accumulator set)
    ('P: dense', 240, 320, 114, 19, 9, 1, 120, 120, 100, 2, False, 2, 128 | 1),
    ('P: dense no MMA', 240, 320, 114, 19, 9, 1, 120, 120, 100, 2, False, 2, 128 | 16 | 1),
    ('P: dense no lo pass', 240, 320, 114, 19, 9, 1, 120, 120, 100, 2, False, 2, 128 | 32 | 1),
    ('P: dense neither', 240, 320, 114, 19, 9, 1, 120, 120, 100, 2, False, 2, 128 | 48 | 1),
    ('Q: dense', 240, 320, 114, 19, 9, 1, 120, 120, 100, 2, False, 2, 1),
    ('Q: dense no MMA', 240, 320, 114, 19, 9, 1, 120, 120, 100, 2, False, 2, 16 | 1),
    ('Q: dense neither', 240, 320, 114, 19, 9, 1, 120, 120, 100, 2, False, 2, 48 | 1),
    ('Q: dense neither no B loads', 240, 320, 114, 19, 9, 1, 120, 120, 100, 2, False, 2, 48 | 256 | 1),
    ('Q: dense neither no A loads', 240, 320, 114, 19, 9, 1, 120, 120, 100, 2, False, 2, 48 | 512 | 1),
    ('Q: dense neither no loads', 240, 320, 114, 19, 9, 1, 120, 120, 100, 2, False, 2, 48 | 768 | 1),
    ('Q: dense neither no loads per-thread stores', 240, 320, 114, 19, 9, 1, 120, 120, 100, 2, False, 2, 48 | 768 | 8 | 1),
    ('Q: dense no B loads', 240, 320, 114, 19, 9, 1, 120, 120, 100, 2, False, 2, 256 | 1),
    ('Q: dense no loads', 240, 320, 114, 19, 9, 1, 120, 120, 100, 2, False, 2, 768 | 1),
    ('Q: 19->19', 240, 320, 19, 19, 9, 1, 20, 20, 0, 2, False, 2, 1),
    ('Q: 19->19 no loads', 240, 320, 19, 19, 9, 1, 20, 20, 0, 2, False, 2, 768 | 1),
    ('Q: 19->19 neither no loads', 240, 320, 19, 19, 9, 1, 20, 20, 0, 2, False, 2, 48 | 768 | 1),
    ('P: dense MT=1', 240, 320, 114, 19, 9, 1, 120, 120, 100, 2, False, 2, 128 | 2 | 1),
    ('P: dense per-tap boxes', 240, 320, 114, 19, 9, 1, 120, 120, 100, 2, False, 2, 128 | 4 | 1),
    ('P: 19->19', 240, 320, 19, 19, 9, 1, 20, 20, 0, 2, False, 2, 128 | 1),
    ('P: 19->19 no MMA', 240, 320, 19, 19, 9, 1, 20, 20, 0, 2, False, 2, 128 | 16 | 1),
    ('P: 19->19 no lo pass', 240, 320, 19, 19, 9, 1, 20, 20, 0, 2, False, 2, 128 | 32 | 1),
    ('P: 19->19 neither', 240, 320, 19, 19, 9, 1, 20, 20, 0, 2, False, 2, 128 | 48 | 1),
    ('P: 19->19 per-thread stores', 240, 320, 19, 19, 9, 1, 20, 20, 0, 2, False, 2, 128 | 8 | 1),
    ('P: 456->114', 240, 320, 456, 114, 1, 1, 456, 116, 0, 0, False, 2, 128 | 1),
    ('P: 456->114 no MMA', 240, 320, 456, 114, 1, 1, 456, 116, 0, 0, False, 2, 128 | 16 | 1),
    ('P: 456->114 neither', 240, 320, 456, 114, 1, 1, 456, 116, 0, 0, False, 2, 128 | 48 | 1),
    ('P: 456->114 nacc 2', 240, 320, 456, 114, 1, 1, 456, 116, 0, 0, False, 2, 128 | 2048 | 1),
    ('P: 19->114 x8', 240, 320, 19, 114, 1, 1, 20, 464, 116, 1, False, 8, 128 | 1),
    ('P: 19->114 x8 no MMA', 240, 320, 19, 114, 1, 1, 20, 464, 116, 1, False, 8, 128 | 16 | 1),
    ('P: 114->95', 240, 320, 114, 95, 1, 1, 116, 116, 0, 2, False, 1, 128 | 1),
    ('P: dense (TS kernel)', 240, 320, 114, 19, 9, 1, 120, 120, 100, 2, False, 2, 65536 | 1),
]
FN = 'ojdf_conv_tc_batched'


def run_case(idx, timing):
    import numpy as np
    import torch
    from online_joint_depthfusion_and_semantic_b200 import _lib
    from online_joint_depthfusion_and_semantic_b200.modules.fusion_engine import ConvProblem
    name, H, W, cin, cout, taps, dil, istr, ostr, ocoff, act, use_res, nprob, flags = CASES[idx]
    dev = torch.device('cuda:0')
    L = _lib.lib()
    g = torch.Generator().manual_seed(100 + idx)
    k = 3 if taps == 9 else 1
    keep, probs, refs, outs = [], [], [], []
    for i in range(nprob):
        x = torch.randn(H * W, istr, generator=g)                           # channels >= cin are garbage on purpose
        w = torch.randn(cout, cin, k, k, generator=g) / (cin * taps) ** 0.5
        sc, sh = 0.5 + torch.rand(cout, generator=g), 0.1 * torch.randn(cout, generator=g)
        res = torch.randn(H * W, cout, generator=g) if use_res else None
        npr = (flags >> 16) & 255
        n = L.ojdf_conv_tc_weight_floats(cin, cout, taps, npr)
        packed = np.zeros(n, np.float32)
        wc = np.ascontiguousarray(w.numpy().reshape(cout, cin, taps))
        _lib.check(L.ojdf_conv_tc_pack_weights(wc.ctypes.data, cin, cout, taps, npr, packed.ctypes.data))
        d = dil if nprob == 1 else max(1, dil - i % 2) if taps == 9 else 1
        xin = x[:, :cin].double().t().reshape(1, cin, H, W)
        y = torch.nn.functional.conv2d(xin, w.double(), padding=d * (k // 2), dilation=d)[0].reshape(cout, H * W).t()
        y = y * sc.double() + sh.double()
        if use_res:
            y = y + res.double()
        y = {0: y, 1: y.clamp(min=0), 2: torch.where(y > 0, y, 0.01 * y), 3: torch.tanh(y), 4: torch.sigmoid(y)}[act]
        out = torch.full((H * W, ostr), 7.0, device=dev)
        t = [x.to(dev), torch.from_numpy(packed).to(dev), sc.to(dev), sh.to(dev), out, res.to(dev) if use_res else None]
        keep.append(t)
        probs.append(ConvProblem(t[0].data_ptr(), t[1].data_ptr(), t[2].data_ptr(), t[3].data_ptr(), out.data_ptr(),
                                 t[5].data_ptr() if use_res else None, istr, ostr, ocoff, d, cout if use_res else 0, 0, 0, 0, 0, 0))
        refs.append(y)
        outs.append(out)
    arr = (ConvProblem * nprob)(*probs)
    st = torch.cuda.current_stream().cuda_stream
    scratch = torch.zeros((64 << 20) // 4, dtype=torch.float32, device=dev)
    scratch_ptr, scratch_bytes = scratch.data_ptr(), scratch.numel() * 4
    _lib.check(getattr(L, FN)(arr, nprob, cin, cout, H, W, taps, act, 0.01, 1.0, (flags >> 16) & 255, flags & 0x1ffff, scratch_ptr, scratch_bytes, st))
    torch.cuda.synchronize()
    worst = 0.0
    for y, out in zip(refs, outs):
        o = out.cpu().double()
        got = o[:, ocoff:ocoff + cout]
        err = float((got - y).abs().max() / y.abs().max())
        worst = max(worst, err)
        untouched = torch.cat([o[:, :ocoff], o[:, ocoff + cout:]], 1)
        if untouched.numel() and not bool((untouched == 7.0).all()) and not (flags & 1):
            worst = float('inf')
    line = '%-44s rel.err %.3e  %s' % (name, worst, 'OK' if worst < 5e-5 else 'FAIL')
    if timing:
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(3):
            getattr(L, FN)(arr, nprob, cin, cout, H, W, taps, act, 0.01, 1.0, (flags >> 16) & 255, flags & 0x1ffff, scratch_ptr, scratch_bytes, st)
        a.record()
        for _ in range(20):
            getattr(L, FN)(arr, nprob, cin, cout, H, W, taps, act, 0.01, 1.0, (flags >> 16) & 255, flags & 0x1ffff, scratch_ptr, scratch_bytes, st)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 20
        line += '  %.1f us  %.1f TFLOP/s (fp32-equivalent)' % (ms * 1e3, 2.0 * H * W * cin * cout * taps * nprob / ms / 1e9)
    print(line, flush=True)
    if flags & 128:
        prof = (C.c_longlong * 32)()
        pf = L.ojdf_conv_tc_profile if (flags & 65536) else L.ojdf_conv_ss_profile
        pf.argtypes = [C.c_void_p]
        pf(prof)                                            # discard what the earlier launches accumulated
        getattr(L, FN)(arr, nprob, cin, cout, H, W, taps, act, 0.01, 1.0, (flags >> 16) & 255, flags & 0x1ffff, scratch_ptr, scratch_bytes, st)
        torch.cuda.synchronize()
        pf(prof)
        p = list(prof)
        print('   block 0 cycles: producer total %d (wait src_empty %d, b_empty %d) | mma total %d (acc_empty %d, b_full %d, a_full %d) | '
              'split0 total %d (src_full %d, a_empty %d) | split1 total %d (src_full %d, a_empty %d) | epilogue total %d (acc_full %d)'
              % (p[2], p[0], p[1], p[6], p[3], p[4], p[5], p[9], p[7], p[8], p[12], p[10], p[11], p[14], p[13]), flush=True)
        print('   raw role profile: %s' % ' '.join('%d:%d' % (i, v) for i, v in enumerate(p) if v), flush=True)
        if not (flags & 65536):
            print('   ss issuer: in the MMA issue blocks %d, in the weight-stage commits %d' % (p[8], p[10]), flush=True)


if __name__ == '__main__':
    timing = '--time' in sys.argv
    if '--case' in sys.argv:
        run_case(int(sys.argv[sys.argv.index('--case') + 1]), timing)
    else:
        only = sys.argv[sys.argv.index('--only') + 1] if '--only' in sys.argv else ''
        for i in range(len(CASES)):
            if not CASES[i][0].startswith(only):
                continue
            try:
                r = subprocess.run([sys.executable, os.path.abspath(__file__), '--case', str(i)] + [a for a in ('--time', '--v1') if a in sys.argv],
                                   capture_output=True, text=True, timeout=180)
                tail = (r.stdout + r.stderr).strip().splitlines()
                print('\n'.join(tail[-4:]) if r.returncode else r.stdout.strip(), flush=True)
            except subprocess.TimeoutExpired:
                print('%-44s TIMEOUT' % CASES[i][0], flush=True)
