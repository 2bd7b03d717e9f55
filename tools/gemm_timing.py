import ctypes, os, sys
import numpy as np, torch
ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(ROOT, "stable-diffusion-on-device_b200"))
from sdod import ops, _cabi as C
M, N, K = (int(v) for v in sys.argv[1:4])
mode = sys.argv[4]
a = torch.randn(M, K, device="cuda").to(torch.bfloat16); w = (torch.randn(N, K, device="cuda") / K ** 0.5).to(torch.bfloat16)
bias = torch.randn(N, device="cuda"); res = torch.randn(M, N, device="cuda")
if mode == "geglu":
    wp, bp = ops.pack_geglu_weight(w, bias, 256)
def run():
    if mode == "geglu":
        return torch.ops.sdod.linear(a, wp, bp, None, C.ACT_GEGLU)
    return torch.ops.sdod.linear(a, w, bias, res, 0, 1.0, True) if mode == "res32" else torch.ops.sdod.linear(a, w)
for _ in range(5): run()
torch.cuda.synchronize()
run(); torch.cuda.synchronize()
lib = C.lib(); lib.sdod_debug_gemm_times.argtypes = [ctypes.c_void_p, ctypes.c_int]
buf = np.zeros(1024 * 8, dtype=np.uint64)
lib.sdod_debug_gemm_times(buf.ctypes.data, 1024 * 8)
bn = 256 if mode == "geglu" else 160
t = buf.reshape(1024, 8)[: min(1024, ((M + 127) // 128) * ((N + bn - 1) // bn))].astype(np.int64)
t0 = t[:, 0].min()
rel = t[:, :8] - t0
print("M%d N%d K%d %s  CTAs %d   (ns since first CTA start; median over CTAs | max)" % (M, N, K, mode, len(t)))
for i, name in enumerate(["start", "setup done", "mainloop done", "phase1 done", "epilogue done", "exit", "p2 bias loaded", "p2 first group done"]):
    print("  %-14s median %7d   max %7d" % (name, np.median(rel[:, i]), rel[:, i].max()))
