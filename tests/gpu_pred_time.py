"""Tensor-core pred() throughput by kernel size (one head block: ks <= 11; several: ks >= 13)."""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import aadff_b200
def timeit(fn, iters=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
for ks in (11, 13, 21, 31):
    torch.manual_seed(0)
    lens = aadff_b200.PSFNet(kernel_size=ks, device="cuda")
    inp = torch.rand(1 << 18, 4, device="cuda")
    ref = lens.pred(inp[:4096])
    got = lens.pred(inp[:4096], mode="parity")
    ms = timeit(lambda: lens.pred(inp, mode="parity"))
    print(f"pred tensor-core parity ks={ks}: {ms:.3f} ms {inp.shape[0] / ms / 1e3:.1f} Mprobes/s  max|tc - fp32| {float((got - ref).abs().max()):.2e}  row sums {float(got.sum((-1, -2)).min()):.6f}..{float(got.sum((-1, -2)).max()):.6f}")
