"""Build kernel variants of libgravhopper_b200.so for A/B measurements on the GPU box:
gravhopper_b200/variants/lib_<name>.so = the normal objects with tree.cu recompiled under extra -D
flags.  Select one at run time with GH_B200_LIB=<path>.  usage: python scripts/build_variants.py"""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gravhopper_b200 import build as B

VARIANTS = {
    # group walk register caps (measured: no gain, profiles/r01_walk_group_regcap_ab.txt)
    # "w24": ["-DGH_GW_WARPS_PER_SM=24"], "w20": ["-DGH_GW_WARPS_PER_SM=20"], "w16": ["-DGH_GW_WARPS_PER_SM=16"],
    # radix-sort scatter kernel: loop structure x register budget (measured: noise level,
    # profiles/r01_sort_scatter_ab.txt)
    # "rs1": ["-DGH_RS_VARIANT=1"], "rs2": ["-DGH_RS_VARIANT=2"], "rs2b3": ["-DGH_RS_VARIANT=2", "-DGH_RS_MINBLOCKS=3"],
    # emit kernel: occupancy against registers
    # "emit8": ["-DGH_EMIT_MINBLOCKS=8"], ... "emit16": no gain (profiles/r01_emit_occupancy_ab.txt)
    # radix sort: keys per thread (tile size)
    # measured (profiles/r02_sort_ab.txt): 12 / 10 / 8 / 6 keys per thread -> build 1.72 / 1.59 / 1.46 / 1.40 ms
    # "rs6"/"rs8", "rk6"/"rk8"/"rk10" (GH_RS_RANK=1): see profiles/r02_sort_ab.txt
    # "place" form of the splitter sort: bits ranked first inside a bucket, tile size
    # measured with the global-atomic form (gpurun_out/ab_*_place_warp.json): sb16 -0.024 ms, sb32 +0.027 ms,
    # rs8 -0.005 ms of build
    # "sb16": ["-DGH_BP_SUBBITS=16"], "sb32": ["-DGH_BP_SUBBITS=32"], "rs8sb16": [...]
    "rs8": ["-DGH_RS_ROUNDS=8"],
    # ranking with MATCH.ANY instead of ballots
    "match0": ["-DGH_RS_MATCH=0"],
}
KERNEL = "bp_bucket_kernel"
B.build()
out = os.path.join(B.HERE, "variants")
os.makedirs(out, exist_ok=True)
flags = [f for f in B.NVCC_FLAGS if not f.startswith("--use_fast_math")]
procs = []
for name, defs in VARIANTS.items():
    o = os.path.join(out, "tree_%s.o" % name)
    procs.append((name, o, subprocess.Popen([B._nvcc()] + flags + defs + ["-Xptxas", "-v", "-c", os.path.join(B.CSRC, "tree.cu"), "-o", o],
                                            stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
for name, o, p in procs:
    log, _ = p.communicate()
    if p.returncode:
        raise SystemExit(log)
    lines = log.splitlines()
    for i, l in enumerate(lines):
        if KERNEL in l and "Compiling" in l:
            print(name, lines[i + 1].strip(), "|", lines[i + 2].strip())
    objs = [os.path.join(B.HERE, "build", s.replace(".cu", ".o")) for s in B.SOURCES if s != "tree.cu"] + [o]
    lib = os.path.join(out, "lib_%s.so" % name)
    subprocess.run([B._nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib] + objs, check=True)
    print("built", lib)
