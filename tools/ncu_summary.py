#!/usr/bin/env python
"""Summarise an .ncu-rep of the pileup kernel: headline metrics + per-source-line instruction and
stall shares (SASS counters joined with nvdisasm line info of the in-tree libvfk.so).

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [kernel-substring] > profiles/rNN_xxx.txt
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
METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "lts__t_bytes.sum", "l1tex__t_bytes.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__grid_size",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__cycles_elapsed.max", "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
]


def ncu(rep, page):
    return subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    want = sys.argv[2] if len(sys.argv) > 2 else "pileup_warp_kernel"
    rows = list(csv.reader(io.StringIO(ncu(rep, "raw"))))
    hdr, units = rows[0], rows[1]
    print("== %s" % os.path.basename(rep))
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        if want not in name:
            continue
        print("-- kernel:", name[:100])
        for m in METRICS:
            if m in hdr:
                print("   %-82s %18s %s" % (m, r[hdr.index(m)], units[hdr.index(m)]))
    # ---- SASS counters x source lines
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "vafator_b200", "libvfk.so")], cwd=tmp, capture_output=True)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")]
    if not cubin:
        return
    dis = "\n".join(subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, c)], capture_output=True, text=True).stdout for c in cubin)
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
    srows = list(csv.reader(io.StringIO(ncu(rep, "source"))))
    blocks, names = [], []
    for r in srows:
        if r and r[0] == "Kernel Name":
            blocks.append([])
            names.append(r[1])
        elif blocks and r and r[0].startswith("0x"):
            blocks[-1].append(r)
    shdr = next(r for r in srows if r and r[0] == "Address")
    ie, sm = shdr.index("Instructions Executed"), shdr.index("# Samples")
    stall_cols = [i for i, c in enumerate(shdr) if c.startswith("stall_") and "Not Issued" not in c]
    src = {}
    for f in os.listdir(os.path.join(ROOT, "vafator_b200", "csrc")):
        src[f] = open(os.path.join(ROOT, "vafator_b200", "csrc", f)).read().split("\n")
    for name, b in zip(names, blocks):
        if want not in name:
            continue
        cands = [k for k, v in funcs.items() if len(v) == len(b)]
        # prefer the instantiation whose mangled name shares the template argument
        key = next((k for k in cands if re.search(r"<\(int\)(\d+)>", name) and ("ILi%sE" % re.search(r"<\(int\)(\d+)>", name).group(1)) in k and want.split("_")[-1] in k), cands[0] if cands else None)
        if key is None:
            print("   (could not match SASS to the in-tree cubin: rebuild differs from the profiled binary)")
            continue
        ins = funcs[key]
        byline, samp, stall = collections.Counter(), collections.Counter(), collections.defaultdict(collections.Counter)
        tot = tots = 0
        for r, (txt, cur) in zip(b, ins):
            n, s = int(r[ie]), int(r[sm])
            byline[cur] += n
            samp[cur] += s
            tot += n
            tots += s
            for i in stall_cols:
                if r[i] not in ("", "0"):
                    stall[cur][shdr[i]] += int(r[i])
        print("-- source lines of %s: %d SASS instructions, %d warp-instructions executed, %d stall samples" % (key[:60], len(ins), tot, tots))
        for k, n in sorted(byline.items(), key=lambda kv: -(kv[1] / max(tot, 1) + samp[kv[0]] / max(tots, 1)))[:40]:
            f, l = k if k else ("?", 0)
            text = src[f][l - 1].strip()[:80] if f in src and 0 < l <= len(src[f]) else ""
            top = ",".join("%s:%d" % (a.replace("stall_", ""), c) for a, c in stall[k].most_common(2))
            print("   %5.1f%% inst %5.1f%% samples  %-16s:%-4d %-80s [%s]" % (100 * n / max(tot, 1), 100 * samp[k] / max(tots, 1), f, l, text, top))
        # opcode mix
        ops = collections.Counter()
        for r, (txt, cur) in zip(b, ins):
            op = txt.split()[1] if txt.startswith("@") else txt.split()[0]
            ops[op.split(".")[0]] += int(r[ie])
        print("-- opcode mix:", ", ".join("%s %.1f%%" % (o, 100 * c / max(tot, 1)) for o, c in ops.most_common(14)))
        # aggregate stall reasons
        agg = collections.Counter()
        for c in stall.values():
            agg.update(c)
        s = sum(agg.values())
        print("-- stall samples:", ", ".join("%s %.1f%%" % (a.replace("stall_", ""), 100 * c / max(s, 1)) for a, c in agg.most_common(8)))


if __name__ == "__main__":
    main()
