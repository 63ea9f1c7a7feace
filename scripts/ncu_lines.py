#!/usr/bin/env python
"""Attribute executed warp-instructions of one profiled kernel to source lines.

ncu's CLI prints per-SASS-instruction counters (`--page source --csv`) but not the CUDA-C correlation, so the
SASS stream is aligned (instruction by instruction, same order) with `nvdisasm -g` line info of a cubin built
from the SAME sources with the SAME flags.

    python scripts/ncu_lines.py <report.ncu-rep> <kernel-substring> <evals-per-launch> [--top N] [--src cb200_edt.cu]

`--src` names the translation unit under curobo_b200/csrc/ that holds the kernel (default cb200_kernels.cu).
"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FLAGS = ["-std=c++17", "-O3", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a", "--ftz=true", "--fmad=true",
         "--prec-div=false", "--prec-sqrt=false"]


def main():
    rep, kname, evals = sys.argv[1], sys.argv[2], int(sys.argv[3])
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 40
    srcdir = os.path.join(ROOT, "curobo_b200", "csrc")
    tmp = tempfile.mkdtemp()
    cubin, sass = os.path.join(tmp, "k.cubin"), os.path.join(tmp, "k.sass")
    unit = sys.argv[sys.argv.index("--src") + 1] if "--src" in sys.argv else "cb200_kernels.cu"
    subprocess.check_call(["nvcc", *FLAGS, "-cubin", "-o", cubin, os.path.join(srcdir, unit)])
    open(sass, "w").write(subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout)
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    secs, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "hdr": None, "rows": []}
            secs.append(cur)
        elif cur is not None and cur["hdr"] is None:
            cur["hdr"] = r
        elif cur is not None:
            cur["rows"].append(r)
    sec = next(s for s in secs if kname in s["name"])
    h = sec["hdr"]
    ii, ti, si = h.index("Instructions Executed"), h.index("Thread Instructions Executed"), h.index("Source")
    ncu = [(r[si].strip(), int(r[ii]), int(r[ti])) for r in sec["rows"] if len(r) > ii]
    want_samples = "--samples" in sys.argv           # rank source lines by warp-stall samples (where the time goes) instead
    if want_samples:
        sa, sb = h.index("# Samples"), h.index("stall_barrier")
        ncu = [(r[si].strip(), int(r[sa]), int(r[sb])) for r in sec["rows"] if len(r) > ii]
    mangled = re.sub(r"[^A-Za-z0-9_]", "", kname.split("<")[0].split("::")[-1])
    targs = re.search(re.escape(kname.split("<")[0].split("::")[-1]) + r"<([^>]*)>", sec["name"])
    tm = ""
    if targs:  # "(int)2, (bool)0, (int)3" -> "ILi2ELb0ELi3EE"
        for a in targs.group(1).split(","):
            m2 = re.match(r"\s*\((int|bool)\)(\d+)", a)
            if m2:
                tm += ("Li" if m2.group(1) == "int" else "Lb") + m2.group(2) + "E"
        tm = "I" + tm + "E"
    txt = open(sass).read().split("\n")
    starts = [i for i, l in enumerate(txt) if l.startswith(".text.") and mangled in l and (not tm or tm in l)]
    start = starts[0]
    end = next(i for i, l in enumerate(txt) if i > start and l.startswith("//---------------------"))
    dis, cf, cl = [], None, None
    for l in txt[start:end]:
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cf, cl = os.path.basename(m.group(1)), int(m.group(2))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+.*?;", l):
            dis.append((cf, cl))
    if len(dis) != len(ncu):
        print(f"WARNING: {len(dis)} disassembled vs {len(ncu)} profiled instructions (sources differ from the profiled build?)")
    if "--static" in sys.argv:
        # instruction-cache footprint: SASS instructions executed by >= 10 % of the evals, per source line and per function
        hot = [k for k in range(min(len(dis), len(ncu))) if ncu[k][1] >= 0.1 * evals]
        print(f"kernel: {sec['name']}\nSASS instructions: {len(ncu)} ({len(ncu) * 16 / 1024:.1f} KB); executed at all: "
              f"{sum(1 for x in ncu if x[1] > 0)}; hot (>= 10 % of evals): {len(hot)} = {len(hot) * 16 / 1024:.1f} KB")
        src = {f: open(os.path.join(srcdir, f)).read().split("\n") for f in os.listdir(srcdir)}

        def func_of(f, l):
            if f not in src or not l:
                return f"{f}:?"
            for i in range(l - 1, -1, -1):
                m = re.match(r"^(?:template.*>\s*)?(?:static\s+)?(?:__device__|CB_HD|__global__|inline|__host__)[^;{]*?([A-Za-z_0-9]+)\s*\(", src[f][i])
                if m and not src[f][i].startswith(" "):
                    return f"{f}:{m.group(1)}"
            return f"{f}:?"
        per_fn, per_line = collections.Counter(), collections.Counter()
        for k in hot:
            per_line[dis[k]] += 1
            per_fn[func_of(*dis[k])] += 1
        print("hot static instructions per function (16 B each):")
        for fn, c in per_fn.most_common(top):
            print(f"{c:6d} {c * 16 / 1024:6.2f} KB  {fn}")
        print("contiguous hot runs (>= 24 instructions): start, length, functions inside")
        run = []
        for k in hot + [10 ** 9]:
            if run and k != run[-1] + 1:
                if len(run) >= 24:
                    fc = collections.Counter(func_of(*dis[i]).split(":")[1] for i in run)
                    print(f"  @{run[0]:5d} len {len(run):4d}  " + ", ".join(f"{f} {c}" for f, c in fc.most_common(5)))
                run = []
            run.append(k)
        print("per line:")
        for (f, l), c in per_line.most_common(top):
            code = src[f][l - 1].strip()[:100] if f in src and l else ""
            print(f"{c:6d}  {f}:{l}  {code}")
        return
    agg, lanes = collections.Counter(), collections.Counter()
    for k in range(min(len(dis), len(ncu))):
        agg[dis[k]] += ncu[k][1]
        lanes[dis[k]] += ncu[k][2]
    tot = sum(agg.values())
    src = {f: open(os.path.join(srcdir, f)).read().split("\n") for f in os.listdir(srcdir)}
    if want_samples:
        print(f"kernel: {sec['name']}\nstall samples: {tot}; columns: samples, share, of which at a barrier")
        for (f, l), c in agg.most_common(top):
            code = src[f][l - 1].strip()[:100] if f in src and l else ""
            print(f"{c:10d} {100 * c / tot:5.1f}% {lanes[(f, l)]:6d}  {f}:{l}  {code}")
        return
    print(f"kernel: {sec['name']}\nwarp-instructions per eval: {tot / evals:.1f}  (total {tot}, {evals} evals)")
    print(f"{'instr/eval':>10} {'share':>6} {'lanes':>5}  source")
    for (f, l), c in agg.most_common(top):
        code = src[f][l - 1].strip()[:100] if f in src and l else ""
        print(f"{c / evals:10.1f} {100 * c / tot:5.1f}% {lanes[(f, l)] / max(c, 1):5.1f}  {f}:{l}  {code}")


if __name__ == "__main__":
    main()
