"""Development aid: build a trace-instrumented copy of the library (-DMPPI_STEP_TRACE), run the
one-kernel step at the bench shape and print when each role of CTA 0..3 finished (us from kernel start)."""
import ctypes as C
import os
import subprocess
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import numpy as np

SRC = os.path.join(ROOT, "mppi_gpu_b200", "csrc")
OUT = os.path.join(ROOT, "tools", "_build", "libmppi_trace.so")
if not os.path.exists(OUT):
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17",
                           "-Xcompiler", "-fPIC", "-DMPPI_STEP_TRACE", "-shared", "-o", OUT,
                           SRC + "/kernels.cu", SRC + "/step.cu", SRC + "/controller.cu",
                           SRC + "/comm.cpp", "-ldl"])
from mppi_gpu_b200 import capi
capi.LIB_PATH = OUT
import mppi_gpu_b200 as m

K, T, A = (int(sys.argv[1]) if len(sys.argv) > 1 else 1000000), 200, 3
ctl = m.PointMassModel(K, T, 0.1, 2 * A, A, flags=capi.FLAG_STEP_KERNEL)
ctl.memcpy_set_data(np.zeros(6), np.zeros(T * A), [1, .5, .75, 0, 0, 0], [1, 1, 1, 5, 5, 5])
for _ in range(4):
    ctl.get_act()
buf = (C.c_uint64 * (4 * 256))()
assert capi.load().mppi_debug_read_step_trace(buf) == 0
tr = np.array(buf[:], dtype=np.uint64).reshape(4, 256).astype(np.int64)
t0 = tr[:, 0].min()
for c in range(4):
    r = tr[c]
    print(f"CTA {c}: start {(r[0]-t0)/1e3:.1f} us")
    for w in range(15):
        rounds = [f"{(r[1+w*8+i]-r[0])/1e3:7.1f}" for i in range(8) if r[1+w*8+i] >= r[0]]
        print(f"  warp {w:2d} rounds done at us:", " ".join(rounds))
    print(f"  producer last box issued {(r[200]-r[0])/1e3:.1f}  consumers done {(r[201]-r[0])/1e3:.1f}  "
          f"record {(r[202]-r[0])/1e3:.1f}")
print(f"finalize done {(tr[0][203]-t0)/1e3:.1f} us")
