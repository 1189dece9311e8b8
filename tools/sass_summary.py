"""Development aid: per-kernel SASS mnemonic counts of libmppi_b200.so (cuobjdump -sass), the
evidence for TMA (UTMALDG / UBLKCP + SYNCS mbarriers), packed FP32x2 (FFMA2/FADD2/FMUL2),
register reallocation (USETMAXREG) and local-memory traffic (STL/LDL) in the hot kernels.
usage: python tools/sass_summary.py > profiles/r01_sass_mnemonics.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sass = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "mppi_gpu_b200", "libmppi_b200.so")],
                      capture_output=True, text=True, check=True).stdout
pat = re.compile(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)")
counts, cur = collections.defaultdict(collections.Counter), None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    m = pat.match(line)
    if m and cur:
        counts[cur][m.group(1)] += 1
WANT = ["UTMALDG", "UBLKCP", "UTMAPF", "SYNCS", "USETMAXREG", "FFMA2", "FADD2", "FMUL2", "FFMA", "MUFU",
        "IMAD", "LOP3", "LDG", "STG", "LDS", "STS", "SHFL", "ATOMG", "REDG", "BAR", "STL", "LDL"]
KEEP = re.compile(r"step_kernel<3, mppi::Model<false, mppi::DoubleIntegrator>|average_kernel<true, true>|"
                  r"sample_kernel<3>|rollout_kernel<3, mppi::Model<false, mppi::DoubleIntegrator>, (true|false), 4>|"
                  r"rollout_tma_kernel<[23], mppi::Model<false, mppi::DoubleIntegrator>, 256>|xchg_merge_finalize")
print("SASS mnemonic counts (static, per kernel) of the sm_100a kernels at the bench shapes; cuobjdump -sass")
for fn, c in sorted(counts.items()):
    d = subprocess.run(["c++filt", fn], capture_output=True, text=True).stdout.strip()
    if not KEEP.search(d):
        continue
    short = re.sub(r"\(.*", "", d).replace("mppi::", "").replace("void ", "")
    print(f"\n{short}: {sum(c.values())} instructions")
    print("  " + "  ".join(f"{k}={c[k]}" for k in WANT if c[k]))
