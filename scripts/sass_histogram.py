"""Opcode histogram of the hot kernels of libfdcm_b200.so from `cuobjdump -sass` (profiles/r02_sass_hot_kernels.txt).
usage: python scripts/sass_histogram.py [lib] > profiles/r02_sass_hot_kernels.txt"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "openfdcm_b200/libfdcm_b200.so"
HOT = ["raster_kernel", "dt_col_band_kernel", "dt_row_band_kernelILb0", "dt_resolve_kernel", "dt_fill_propagate_kernel", "dt_l1_propagate_kernel",
       "integral_tma_kernel", "search_key_kernel", "search_warp_kernel", "topk_level1_kernel", "topk_level2_kernel", "topk_merge_kernel"]
NOTABLE = re.compile(r"^(UTMA|UBLK|SYNCS|MUFU|REDUX|CREDUX|SHFL|VOTE|BAR|FENCE|HMMA|UTCMMA|ATOMS|RED|LDGSTS|MATCH)")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
print(f"# cuobjdump -sass {lib} (sm_100a): opcode histogram of the hot kernels")
print("# things to look for: UTMALDG.3D (TMA tensor loads) + SYNCS.* (mbarrier) in integral_tma_kernel; REDUX.OR / SHFL / MUFU.RSQ in the fused fill;")
print("# MUFU.RCP in the envelope kernel (biased-reciprocal quotient); SHFL.BFLY butterfly in dt_col_band_kernel; no HMMA / UTCMMA anywhere")
print("# (nothing on these paths is a contraction)")
fn, ops = None, collections.Counter()


def flush():
    if fn and any(h in fn for h in HOT):
        n = sum(ops.values())
        print(f"\n## {fn}  ({n} instructions)")
        print("  " + ", ".join(f"{k} x{v}" for k, v in ops.most_common(28)))
        note = sorted((k, v) for k, v in ops.items() if NOTABLE.match(k))
        print("  notable: " + ", ".join(f"{k} x{v}" for k, v in note))


for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        flush()
        fn, ops = m.group(1), collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m:
        ops[m.group(1)] += 1
flush()
