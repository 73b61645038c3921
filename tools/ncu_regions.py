#!/usr/bin/env python
"""Warp-instructions per unit by source function, from an .ncu-rep of the shallow pileup kernel
(SASS counters joined with nvdisasm line info of the in-tree libvfk.so; the library must be the
build that was profiled).

    python tools/ncu_regions.py gpurun_out/prof.ncu-rep N_UNITS [kernel-substring] [--lines]
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "vafator_b200", "csrc")


def functions_of(path):
    """[(first_line, name)] of every function definition, in file order."""
    out = []
    for n, ln in enumerate(open(path).read().split("\n"), 1):
        m = re.match(r"\s*(?:__device__|__global__|static|inline|template).*?\b([A-Za-z_]\w*)\s*\(", ln)
        if ln.startswith("template") or not m:
            m2 = re.match(r"([a-z_]\w*)\(", ln)  # continuation line of a __global__ declaration
            if not m2:
                continue
            m = m2
        if m.group(1) in ("__launch_bounds__", "if", "for", "while"):
            continue
        out.append((n, m.group(1)))
    return out


def main():
    rep, n_units = sys.argv[1], float(sys.argv[2])
    want = sys.argv[3] if len(sys.argv) > 3 and not sys.argv[3].startswith("--") else "pileup_warp_kernel"
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "vafator_b200", "libvfk.so")], cwd=tmp, capture_output=True)
    cubins = [f for f in os.listdir(tmp) if f.endswith(".cubin")]
    dis = "\n".join(subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, c)], capture_output=True, text=True).stdout for c in cubins)
    funcs, fn, cur = {}, None, None
    for ln in dis.split("\n"):
        m = re.match(r"\s*\.text\.(\S+):", ln)
        if m:
            fn, cur = m.group(1), None
            funcs[fn] = []
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m and fn:
            funcs[fn].append((m.group(2).strip(), cur))
    srows = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout)))
    blocks, names = [], []
    for r in srows:
        if r and r[0] == "Kernel Name":
            blocks.append([])
            names.append(r[1])
        elif blocks and r and r[0].startswith("0x"):
            blocks[-1].append(r)
    shdr = next(r for r in srows if r and r[0] == "Address")
    ie, te = shdr.index("Instructions Executed"), shdr.index("Thread Instructions Executed")
    src = {f: open(os.path.join(CSRC, f)).read().split("\n") for f in os.listdir(CSRC)}
    fmap = {f: functions_of(os.path.join(CSRC, f)) for f in src}

    def region(cur):
        if cur is None:
            return "(no line info)"
        f, l = cur
        if f not in fmap:
            return f
        name = "?"
        for first, nm in fmap[f]:
            if first <= l:
                name = nm
        return name

    for name, b in zip(names, blocks):
        if want not in name:
            continue
        cands = [k for k, v in funcs.items() if len(v) == len(b) and want in k]
        if not cands:
            print("could not match the SASS of", name[:60], "to the in-tree cubin")
            continue
        ins = funcs[cands[0]]
        agg, th, byl, thl = collections.Counter(), collections.Counter(), collections.Counter(), collections.Counter()
        tot = 0
        for r, (txt, cur) in zip(b, ins):
            n, t = int(r[ie]), int(r[te])
            g = region(cur)
            agg[g] += n
            th[g] += t
            byl[cur] += n
            thl[cur] += t
            tot += n
        print("== %s: %d SASS instructions, %.1f warp-instructions per unit" % (name[:48], len(ins), tot / n_units))
        for g, n in agg.most_common():
            print("   %-26s %5.1f%%  %7.1f instr/unit   %4.1f active lanes" % (g, 100 * n / tot, n / n_units, th[g] / max(n, 1)))
        if "--lines" in sys.argv:
            for cur in sorted([c for c in byl if c and c[0] in src], key=lambda c: (c[0], c[1])):
                n = byl[cur]
                if n / n_units >= 1.0:
                    print("   %-15s %4d %6.1f instr/unit %4.1f lanes  %s" % (cur[0], cur[1], n / n_units, thl[cur] / max(n, 1), src[cur[0]][cur[1] - 1].strip()[:90]))
        break


if __name__ == "__main__":
    main()
