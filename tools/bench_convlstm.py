"""Time one ConvLSTM step (K2) at workload c3 size: C = F = 256, 64^3 voxels: tensor-core path vs the fp32 CUDA-core kernel."""
import sys
import torch
sys.path.insert(0, '.')
import mulit_view_object_detection_b200 as m

X = int(sys.argv[1]) if len(sys.argv) > 1 else 64
C = F = int(sys.argv[2]) if len(sys.argv) > 2 else 256
run_fp32 = len(sys.argv) > 3 and sys.argv[3] == 'fp32'
dev = torch.device('cuda:0')
g = torch.Generator(device=dev); g.manual_seed(0)
W = torch.randn((3, 3, 3, C + F, 4 * F), device=dev, generator=g) * (2.0 / (27 * (C + F) + 4 * F)) ** 0.5
b = torch.randn(4 * F, device=dev, generator=g) * 0.1
x = torch.randn((1, X, X, X, C), device=dev, generator=g).relu_()
cell = m.ConvLSTMTensorCore(W, b, 1.0)
h, c = cell.step(x, None, None)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for has_h in (False, True):
    for _ in range(2):
        h2, c2 = cell.step(x, h if has_h else None, c if has_h else None)
    torch.cuda.synchronize()
    n = 3
    e0.record()
    for _ in range(n):
        h2, c2 = cell.step(x, h if has_h else None, c if has_h else None)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    K = 27 * (C + (F if has_h else 0))
    flop = 2.0 * X ** 3 * K * 4 * F
    print("tc   has_h=%d  %.2f ms/step  %.1f TFLOP/s useful (x3 MMA work: %.1f TF/s of f16 MMA)" % (has_h, ms, flop / ms / 1e9, 3 * flop / ms / 1e9))
if run_fp32:
    hf, cf = m.convlstm_step(x, h, c, W, b)
    torch.cuda.synchronize()
    e0.record(); hf, cf = m.convlstm_step(x, h, c, W, b); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print("fp32 has_h=1  %.2f ms/step  %.1f TFLOP/s" % (ms, 2.0 * X ** 3 * 27 * (C + F) * 4 * F / ms / 1e9))
    print("max |h_tc - h_fp32| =", float((h2 - hf).abs().max()), " max |c| diff =", float((c2 - cf).abs().max()))
