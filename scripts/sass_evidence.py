"""SASS evidence of the built library (runs on the CPU box: cuobjdump only): per kernel the instruction mix that proves what
the design claims -- UBLKCP (cp.async.bulk, the TMA engine's bulk copies) and SYNCS (mbarrier) in the streamed CG, the FP64
mix (DFMA / DMUL / DADD) and the shared-memory traffic (LDS / STS) of the assembly kernel, no tensor-core instructions
anywhere -- plus an excerpt of each hot loop.
usage: python scripts/sass_evidence.py  ->  profiles/sass/summary.md, profiles/sass/<kernel>.sass.txt (excerpts)"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "onsas.jl_b200", "libonsas_cuda.so")
OUT = os.path.join(ROOT, "profiles", "sass")
os.makedirs(OUT, exist_ok=True)
sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
arch = re.findall(r"arch = (sm_\w+)", sass)
kernels, cur = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        kernels[cur] = []
    elif cur is not None and re.match(r"\s*/\*[0-9a-f]{4,}\*/", line):
        kernels[cur].append(line)
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
WATCH = ["UBLKCP", "UBLKPF", "SYNCS", "DFMA", "DMUL", "DADD", "LDS", "STS", "LDG", "STG", "LDL", "STL", "BAR", "SHFL", "HMMA", "IMMA", "DMMA", "UTCMMA", "UTMALDG"]
rows = []
for name, lines in kernels.items():
    ops = collections.Counter()
    for l in lines:
        m = re.search(r"\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", l)
        if m:
            ops[m.group(1)] += 1
    rows.append((demangle(name), len(lines), ops))
sel = [r for r in rows if re.search(r"cg_stream<3, 12|k_assemble_reg<0, [01], 3|k_assemble<1, 0, 3, false|k_coarse_assemble<3>|k_gj_invert_blocked|k_spmv_dot<3>", r[0])]
with open(os.path.join(OUT, "summary.md"), "w") as f:
    f.write("# SASS instruction mix of `onsas.jl_b200/libonsas_cuda.so`\n\n")
    f.write(f"`cuobjdump -sass` of the library built by `make -C onsas.jl_b200/csrc` (nvcc 12.9, `-gencode arch=compute_100a,code=sm_100a`): "
            f"arch = {sorted(set(arch))}, {len(kernels)} kernels.  Static instruction counts (not execution counts).\n\n")
    f.write("| kernel | SASS instrs | " + " | ".join(WATCH) + " |\n|---|---|" + "---|" * len(WATCH) + "\n")
    for name, n, ops in sel:
        f.write(f"| `{name[:90]}` | {n} | " + " | ".join(str(ops.get(w, 0)) for w in WATCH) + " |\n")
    tot = collections.Counter()
    for _, _, ops in rows:
        tot.update(ops)
    f.write("\nWhole library: " + ", ".join(f"{w} {tot.get(w, 0)}" for w in WATCH) + ".\n\n")
    f.write("* `UBLKCP` = `cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes` (the bulk-copy / TMA engine): K's slices are "
            "streamed into the shared-memory rings of `cg_stream` by these, completion on `SYNCS` (mbarrier) objects.\n"
            "* No `HMMA` / `IMMA` / `DMMA` / `UTCMMA`: nothing on this path is a dense contraction (north star: no tensor cores).\n"
            "* `k_assemble_reg<0, KIND, 3, false, 96>`: the fused tet evaluation + assembly at 96 registers; `LDL` / `STL` are its spills.\n")
for name, n, ops in sel:
    key = re.sub(r"[^A-Za-z0-9]+", "_", name.split("(")[0])[:60]
    lines = kernels[[k for k in kernels if demangle(k) == name][0]]
    idx = [i for i, l in enumerate(lines) if re.search(r"UBLKCP|SYNCS", l)] if "cg_stream" in name else [i for i, l in enumerate(lines) if "DFMA" in l]
    if not idx:
        continue
    lo = max(0, idx[0] - 6)
    with open(os.path.join(OUT, key + ".sass.txt"), "w") as f:
        f.write(f"// {name}\n// excerpt around the first of {len(idx)} matching instructions (of {n})\n")
        f.write("\n".join(l.rstrip() for l in lines[lo:lo + 70]) + "\n")
print(open(os.path.join(OUT, "summary.md")).read())
