"""List backward-branch loops in the SASS of one kernel (static instruction counts, no GPU needed).

    python scripts/sass_loops.py dphtools_b200/csrc/build/lm_inst_gauss2d.o 'StageBINS_7Gauss2DELb1EEES2_Li4E'
"""
import re
import subprocess
import sys

obj, pat = sys.argv[1], sys.argv[2]
out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
funcs = out.split("Function : ")
for f in funcs[1:]:
    name = f.split("\n", 1)[0]
    if pat not in name:
        continue
    ins = []
    for line in f.split("\n"):
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2)))
    print(name[:120], "total SASS", len(ins))
    addr_index = {a: i for i, (a, _) in enumerate(ins)}
    for i, (a, t) in enumerate(ins):
        m = re.search(r"BRA\S*\s+(?:\S+,\s*)?`?\(?\.?(?:L_x_\d+|0x[0-9a-f]+)", t)
        m2 = re.search(r"BRA.*?(0x[0-9a-f]+)", t)
        if m2:
            tgt = int(m2.group(1), 16)
            if tgt < a and tgt in addr_index:
                j = addr_index[tgt]
                body = [x[1] for x in ins[j:i + 1]]
                ops = {}
                for b in body:
                    tk = b.split()
                    op = (tk[1] if tk[0].startswith("@") else tk[0]).split(".")[0]
                    ops[op] = ops.get(op, 0) + 1
                if len(body) > 40:
                    print("  loop %5d..%5d  n=%4d  %s" % (j, i, len(body), sorted(ops.items(), key=lambda x: -x[1])[:12]))
