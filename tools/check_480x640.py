import sys, torch
sys.path.insert(0, '/root/repo')
sys.path.insert(0, '/root/repo/tests')
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
import test_gpu_fusionnet as tf, test_gpu_adapnet as ta
from online_joint_depthfusion_and_semantic_b200.modules.model import FusionNet_v3
h, w = 480, 640
net = tf._net(FusionNet_v3, h, w, True)
x = tf._inputs(h, w)
with torch.no_grad():
    net.use_engine = False; ref = net(x)
    net.use_engine = True; out = net(x)
torch.cuda.synchronize()
print('FusionNet 480x640 rel err %.2e' % (float((out - ref).abs().max()) / float(ref.abs().max())))
a = ta._net(2)
g = torch.Generator().manual_seed(3)
x1, x2 = torch.randn(1, 3, h, w, generator=g).cuda(), torch.randn(1, 3, h, w, generator=g).cuda()
with torch.no_grad():
    a.use_engine = False; r = a(x1, x2)
    a.use_engine = True; o = a(x1, x2)
torch.cuda.synchronize()
print('AdapNet 480x640 rel err %.2e, label agreement %.5f' % (float((o[0] - r[0]).abs().max()) / float(r[0].abs().max()),
      float((o[0].argmax(1) == r[0].argmax(1)).float().mean())))
import time
for name, fn in (('FusionNet', lambda: net(x)), ('AdapNet', lambda: a(x1, x2))):
    with torch.no_grad():
        for _ in range(3): fn()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(10): fn()
        torch.cuda.synchronize()
    print('%s 480x640 eager %.2f ms' % (name, (time.perf_counter() - t0) * 100))
