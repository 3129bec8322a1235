"""Cycle timeline of gemm_tcgen05_kernel for the PPI-shaped node transform T = H.[W0|W1|W2] (M=2245, K=256, N=768).
Builds / loads lib/librgnn_trace.so (same sources, -DRGNN_GEMM_TRACE: clock64 stamps per CTA, see gemm_tcgen05.cu) -- run
`python tools/gemm_trace.py --build-only` where nvcc is available, then `python tools/gemm_trace.py` on the GPU."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tf_gnn_samples_b200 import _build

path = os.path.join(_build.LIB_DIR, "librgnn_trace.so")
if "--build-only" in sys.argv or not os.path.exists(path):
    path = _build.build_variant("trace", ["-DRGNN_GEMM_TRACE"])
    if "--build-only" in sys.argv:
        print("built", path); sys.exit(0)
_build.LIB_PATH = path                      # point the package at the instrumented library
import numpy as np, torch
import tf_gnn_samples_b200 as G
from tf_gnn_samples_b200 import ops
from tf_gnn_samples_b200.engine import load_library

dev = torch.device("cuda", 0)
G.set_weight_cache(True)
V, K, N = 2245, 256, 768
h = torch.randn(V, K, device=dev)
w = torch.randn(K, N, device=dev) * 0.05
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(5):
    ops.dense(h, w)
flush.zero_(); torch.cuda.synchronize()
ops.dense(h, w); torch.cuda.synchronize()          # the traced launch (cold L2, cached weight images)
lib = load_library()
lib.rgnn_debug_gemm_trace.restype = ctypes.c_int
buf = (ctypes.c_longlong * (160 * 64))()
assert lib.rgnn_debug_gemm_trace(buf, 160 * 64) == 0
t = np.frombuffer(buf, dtype=np.int64).reshape(160, 64)
names = {0: "entry", 1: "setup done", 50: "accumulator complete", 51: "epilogue warp0 done", 52: "helper (warp 4) starts",
         53: "helper done", 54: "teardown", 55: "first B copy issued", 56: "last B copy issued"}
for cta in (0, 71, 143):
    r = t[cta]; t0 = r[0]
    print("CTA %d" % cta)
    print("  " + "  ".join("%s %d" % (names[k], r[k] - t0) for k in (1, 55, 56)))
    print("  chunk:        " + " ".join("%6d" % q for q in range(8)))
    print("  A published:  " + " ".join("%6d" % (r[2 + q] - t0) for q in range(8)))
    print("  full seen:    " + " ".join("%6d" % (r[18 + q] - t0) for q in range(8)))
    print("  MMA committed:" + " ".join("%6d" % (r[34 + q] - t0) for q in range(8)))
    print("  " + "  ".join("%s %d" % (names[k], r[k] - t0) for k in (50, 52, 51, 53, 54)))
