"""Write the SASS of the pixel-pair loops of the headline kernels (profiles/r2_sass_hot_loops.txt).  No GPU needed.

    python scripts/sass_hot_loop_listing.py > profiles/r2_sass_hot_loops.txt
"""
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(REPO, "dphtools_b200", "csrc", "build", "lm_inst_gauss2d.o")
KERNELS = [
    ("StageB<Gauss2D, MLE>, G = 2 (lm() loop, the kernel bench.py's pipelined steps run)", "StageBINS_7Gauss2DELb1EEES2_Li2ENS_8IoDeviceIS3_S2_Li2ELb1ELb1ELb0"),
    ("StageA<Gauss2D>, G = 2 (least-squares pre-fit)", "StageAINS_7Gauss2DEEES2_Li2ENS_8IoDeviceIS3_S2_Li2ELb0ELb0ELb0"),
    ("StageB<Gauss2D, MLE>, G = 4", "StageBINS_7Gauss2DELb1EEES2_Li4ENS_8IoDeviceIS3_S2_Li4ELb1ELb1ELb0"),
]
out = subprocess.run(["cuobjdump", "-sass", OBJ], capture_output=True, text=True).stdout
funcs = out.split("Function : ")
print("SASS of the hot pixel-pair loops (cuobjdump -sass dphtools_b200/csrc/build/lm_inst_gauss2d.o, sm_100a, nvcc 12.9, -O3 -lineinfo).")
print("The loop body holds TWO pixel pairs per lane (Gauss2D::kUnroll = 2): four pixels per trip of the loop.  FFMA2 / FMUL2 / FADD2 are")
print("Blackwell's packed fp32 instructions; MUFU.EX2 x2 / MUFU.RCP x2 per pixel (the MUFU.LG2 sit in rarely taken branches: steps that")
print("change a pixel by more than 12 %, and the chi-square of the first pass); no LDL/STL (no spills).")
for title, pat in KERNELS:
    for f in funcs[1:]:
        name = f.split("\n", 1)[0]
        if pat not in name:
            continue
        ins = []
        for line in f.split("\n"):
            m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
            if m:
                ins.append((int(m.group(1), 16), m.group(2).strip()))
        index = {a: i for i, (a, _) in enumerate(ins)}
        best = None
        for i, (a, t) in enumerate(ins):
            m = re.search(r"BRA.*?(0x[0-9a-f]+)", t)
            if m and int(m.group(1), 16) < a and int(m.group(1), 16) in index:
                j = index[int(m.group(1), 16)]
                body = ins[j:i + 1]
                n2 = sum(1 for _, x in body if "FFMA2" in x)
                if n2 and (best is None or n2 > best[0]) and len(body) < 600:
                    best = (n2, body)
        body = best[1]
        ops = {}
        for _, b in body:
            tk = b.split()
            op = (tk[1] if tk[0].startswith("@") else tk[0]).split(".")[0]
            ops[op] = ops.get(op, 0) + 1
        print("\n" + "=" * 110)
        print("%s: %d SASS instructions in the loop (whole kernel %d)" % (title, len(body), len(ins)))
        print("  mix: " + ", ".join("%s %d" % kv for kv in sorted(ops.items(), key=lambda x: -x[1])))
        for a, t in body:
            print("    /*%04x*/  %s" % (a, t))
        break
