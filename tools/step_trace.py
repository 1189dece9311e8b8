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
OUT = os.path.join(ROOT, "tools", "_build", os.environ.get("MPPI_TRACE_LIB", "libmppi_trace.so"))
if not os.path.exists(OUT):
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17",
                           "-Xcompiler", "-fPIC", "-DMPPI_STEP_TRACE", "-shared", "-o", OUT,
                           SRC + "/kernels.cu", SRC + "/step.cu", SRC + "/tile.cu", SRC + "/controller.cu",
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
    print(f"  producer last box issued {(r[200]-r[0])/1e3:.1f}  consumers done {(r[201]-r[0])/1e3:.1f}  "
          f"record {(r[202]-r[0])/1e3:.1f}")
print(f"finalize done {(tr[0][203]-t0)/1e3:.1f} us")
buf2 = (C.c_uint64 * (4 * 3 * 128))()
assert capi.load().mppi_debug_read_step_trace_li(buf2) == 0
tl = np.array(buf2[:], dtype=np.uint64).reshape(4, 3, 128).astype(np.int64)
c = 0
print("CTA 0 list entries: li  rollout_done  producer_start  producer_issued_all  lag(us)")
for li in range(128):
    if tl[c, 0, li] == 0:
        break
    d, ps, pe = [(tl[c, j, li] - tr[c, 0]) / 1e3 for j in range(3)]
    print(f"  {li:3d} {d:8.1f} {ps:8.1f} {pe:8.1f}   lag {ps - d:7.1f}  issue {pe - ps:6.1f}")
# all CTAs: clear, run one more step, read
buf3 = (C.c_uint64 * (160 * 8))()
capi.load().mppi_debug_read_step_trace_all(buf3, 1)
ctl.get_act()
assert capi.load().mppi_debug_read_step_trace_all(buf3, 0) == 0
ta = np.array(buf3[:], dtype=np.uint64).reshape(160, 8).astype(np.int64)[:148]
t0 = ta[:, 0].min()
names = ["start", "rollout_done", "producer_done", "consumers_done", "ticket"]
for j, nm in enumerate(names):
    v = (ta[:, j] - t0) / 1e3
    print(f"all CTAs {nm:15s}: min {v.min():7.1f} p10 {np.percentile(v,10):7.1f} median {np.median(v):7.1f} "
          f"p90 {np.percentile(v,90):7.1f} max {v.max():7.1f} (argmax CTA {int(v.argmax())})")
print(f"merge begins {(ta[0,5]-t0)/1e3:.1f}  finalize done {(ta[0,6]-t0)/1e3:.1f}  last CTA {ta[0,7]}")
late = np.argsort(ta[:, 4])[-8:]
print("latest CTAs:", [(int(c), round((ta[c,1]-t0)/1e3,1), round((ta[c,4]-t0)/1e3,1)) for c in late])
