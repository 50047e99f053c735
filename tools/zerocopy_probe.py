"""Probe: can kernels stream the per-query tensors straight from / to PINNED HOST memory (UVA zero-copy)
at the rate the copy engines reach?  Uses the value-epilogue kernel (a pure streaming pass) through the C ABI
with host pointers.  Developer tool (decides whether a staging-free e2e path is worth building)."""
import ctypes as C
import importlib
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
capi = importlib.import_module("tce-rvos_b200.capi")
lib = capi.load()

rows, width = 10 * 4820 * 2, 256                      # 98.7 MB fp32
nbytes = rows * width * 4
h_src = torch.randn(rows, width).pin_memory()
h_dst = torch.empty(rows, width).pin_memory()
d_src = h_src.cuda()
d_dst = torch.empty_like(d_src)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)


def run(src, dst, n=5):
    def once():
        rc = lib.msda_value_epilogue(rows, width, src.data_ptr(), capi.MSDA_F32, None, dst.data_ptr(), capi.MSDA_F32, st)
        assert rc == 0, rc
    once(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        once()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n


for name, s, d in (("device -> device", d_src, d_dst), ("pinned host -> device (kernel reads over PCIe)", h_src, d_dst),
                   ("device -> pinned host (kernel writes over PCIe)", d_src, h_dst), ("pinned host -> pinned host", h_src, h_dst)):
    t = run(s, d)
    print(f"{name:52s} {t * 1e3:8.3f} ms   {nbytes / t / 1e9:7.1f} GB/s per direction")
assert torch.equal(h_dst, h_src)
t = run(d_src, d_dst)
a = time.perf_counter(); d_dst.copy_(h_src, non_blocking=True); torch.cuda.synchronize(); b = time.perf_counter()
print(f"copy engine H2D for comparison: {nbytes / (b - a) / 1e9:.1f} GB/s")
