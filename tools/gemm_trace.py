"""Measurement driver (not a pytest file): timeline of CTA 0 inside one decoder-sized GEMM launch.  Needs an instrumented
build of the library (gemm_tcgen05.cu compiled with -DSST_GEMM_TRACE, everything else as in csrc/build.sh):

    cd <pkg>/csrc && mkdir -p /tmp/tr && for f in *.cu; do nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 \
        -Xcompiler -fPIC --expt-relaxed-constexpr -DSST_GEMM_TRACE -c $f -o /tmp/tr/${f%.cu}.o; done && \
        nvcc -gencode arch=compute_100a,code=sm_100a -shared -o <repo>/tools/libsst_trace.so /tmp/tr/*.o -lcudart
    SST_LIB=<repo>/tools/libsst_trace.so python tools/gemm_trace.py

The stamps (%globaltimer by one lane + a global store) cost ~0.1 us each: read differences of several stamps, not single gaps."""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sst_b200  # noqa
from sst_b200 import lib as L

M = 7744
for (N, K) in [(768, 64), (768, 768), (2304, 768)]:
    x = (torch.randn(M, K, device="cuda") * 0.5).bfloat16()
    w = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
    y = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    for _ in range(3):
        L.gemm(x, w, y, M, N, K, K, K, N)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); L.gemm(x, w, y, M, N, K, K, K, N); e1.record()
    torch.cuda.synchronize()
    buf = (C.c_ulonglong * 64)()
    assert L.lib().sst_debug_gemm_trace(buf) == 0
    t = list(buf)
    t0 = t[0]
    names = {0: "entry", 1: "setup done", 2: "first TMA issued", 3: "first operands landed", 4: "epilogue warp at final barrier", 5: "after final barrier"}
    print("GEMM %dx%dx%d: event-timed %.1f us" % (M, N, K, e0.elapsed_time(e1) * 1e3))
    for k in sorted(names):
        print("   %-32s +%6.2f us" % (names[k], (t[k] - t0) / 1e3))
    for i in range(8, 20, 2):
        if t[i] > t0 and t[i] - t0 < 10**9:
            k = (i - 8) // 2
            m = 20 + 3 * k
            print("   tile %d: issuer arrives +%6.2f, buffer released +%6.2f" % (k, (t[40 + 2 * k] - t0) / 1e3, (t[41 + 2 * k] - t0) / 1e3))
            print("   tile %d: MMA warp: buffer free +%6.2f, first operands +%6.2f, last operands +%6.2f | epilogue: accumulator ready +%6.2f us, done +%6.2f us"
                  % (k, (t[m] - t0) / 1e3, (t[m + 1] - t0) / 1e3, (t[m + 2] - t0) / 1e3, (t[i] - t0) / 1e3, (t[i + 1] - t0) / 1e3))
