"""Does a concurrent H2D DMA slow the coverage kernel down? (debug aid)"""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import coverage_b200 as cov
e = cov.CoverageEngine(0)
bits, n = cov.synth.fire_grid(256); d = 500 / 256
e.set_grid_bits(bits, 256, 256, d, d); e.set_params(5, np.full(5, 30 * cov.TAN_HALF_FOV_DEFAULT))
B = 139808
dX = e.device_alloc(B * 120); dobj = e.device_alloc(B * 8); dc = e.device_alloc(B * 8); df = e.device_alloc(B)
e.generate_candidates(dX, B, 5, seed=1); e.sync()
def kernels(n):
    ms0, l0 = e.kernel_time_total()
    for _ in range(n): e.eval_batch_device(dX, B, dobj, dc, df)
    e.sync(); ms1, l1 = e.kernel_time_total()
    return (ms1 - ms0) / (l1 - l0)
kernels(5)
print("kernel alone: %.3f ms" % kernels(50))
h = torch.empty(16_777_216, dtype=torch.uint8).pin_memory(); g = torch.empty(16_777_216, dtype=torch.uint8, device="cuda")
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    for _ in range(400): g.copy_(h, non_blocking=True)   # ~0.3 ms each: 120 ms of back-to-back DMA
time.sleep(0.005)
print("kernel with concurrent H2D: %.3f ms" % kernels(50))
torch.cuda.synchronize()
o = torch.empty(2_500_000, dtype=torch.uint8, device="cuda"); ho = torch.empty(2_500_000, dtype=torch.uint8).pin_memory()
with torch.cuda.stream(s):
    for _ in range(3000): ho.copy_(o, non_blocking=True)
time.sleep(0.005)
print("kernel with concurrent D2H: %.3f ms" % kernels(50))
torch.cuda.synchronize()
# pure-compute control: the brute kernel touches almost no global memory per test
e.set_option(cov.OPT_KERNEL, cov.KERNEL_BRUTE)
B = 4096
def kb(n):
    ms0, l0 = e.kernel_time_total()
    for _ in range(n): e.eval_batch_device(dX, B, dobj, dc, df)
    e.sync(); ms1, l1 = e.kernel_time_total()
    return (ms1 - ms0) / (l1 - l0)
kb(2)
print("brute alone: %.3f ms" % kb(10))
with torch.cuda.stream(s):
    for _ in range(400): g.copy_(h, non_blocking=True)
time.sleep(0.005)
print("brute with concurrent H2D: %.3f ms" % kb(10))
torch.cuda.synchronize()
import subprocess
print(subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.mem,power.draw", "--format=csv,noheader"], capture_output=True, text=True).stdout)
