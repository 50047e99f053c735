"""PCIe ceiling and host-pipeline timing probe (developer tool)."""
import importlib, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
capi = importlib.import_module("tce-rvos_b200.capi")
from tools import synthetic as O

nbytes = 173_000_000
h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
h2 = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
d = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
d2 = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, n=10):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n
a = t(lambda: d.copy_(h, non_blocking=True))
b = t(lambda: h2.copy_(d2, non_blocking=True))
def both():
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
c = t(both)
print(f"H2D {nbytes/a/1e9:.1f} GB/s  D2H {nbytes/b/1e9:.1f} GB/s  concurrent {nbytes/c/1e9:.1f} GB/s each way ({c*1e3:.2f} ms for 173 MB each way)")

inp = O.make_inputs("model", 10, O.LEVELS_A, seed=0)
pin = {k: inp[k].pin_memory() for k in ("value", "loc", "aw", "grad_out")}
res = [torch.empty(10, 4820, 256).pin_memory(), torch.empty_like(inp["value"]).pin_memory(),
       torch.empty_like(inp["loc"]).pin_memory(), torch.empty_like(inp["aw"]).pin_memory()]
sh = capi.shape_of(pin["value"], pin["loc"], 4)
ctx = capi.HostContext(0)
f = lambda: ctx.forward_backward(sh, pin["value"], inp["shapes"], inp["lsi"], pin["loc"], pin["aw"], pin["grad_out"], *res)
print(f"MSDA_HOST_CHUNKS={os.environ.get('MSDA_HOST_CHUNKS', '8(default)')}: fwd+bwd host call {t(f, 20)*1e3:.2f} ms")
