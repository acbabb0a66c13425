#!/usr/bin/env python
"""Attribute an ncu source-page capture to OUTERMOST source lines of one .cu file.

    python tools/ncu_lines.py <report.ncu-rep> <kernel .cu file> [--cubin-regex solve_t1] [--min 0.4]

ncu's CSV export of the source page is SASS-only; `nvdisasm --print-line-info-inline` of the same cubin (extracted from
optik_b200/lib/liboptik_b200.so) gives, per SASS instruction, the inline chain.  The i-th instruction of both listings
is the same instruction (same build), so samples / executed instructions are summed per line of the kernel file that
the instruction was inlined into.  Prints % of executed warp instructions, % of stall samples, active threads per
instruction and the share of fp64-pipe instructions per line.
"""
import argparse
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("cu")
    ap.add_argument("--kernel-regex", default=None, help="select the ncu result whose kernel name matches")
    ap.add_argument("--index", type=int, default=0, help="which captured launch of the report (0 = first)")
    ap.add_argument("--min", type=float, default=0.4, help="print lines with at least this %% of instructions or samples")
    ap.add_argument("--so", default=os.path.join(ROOT, "optik_b200", "lib", "liboptik_b200.so"))
    a = ap.parse_args()
    cu_name = os.path.basename(a.cu)
    stem = cu_name.rsplit(".", 1)[0]
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(a.so)], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
    cubin = [f for f in os.listdir(tmp) if f.startswith(stem + ".")][0]
    dis = subprocess.run(["nvdisasm", "--print-line-info-inline", os.path.join(tmp, cubin)], capture_output=True,
                         text=True, check=True).stdout
    cmd = ["ncu", "-i", a.report, "--page", "source", "--csv"]
    if a.kernel_regex:
        cmd += ["-k", "regex:" + a.kernel_regex]
    rows = list(csv.reader(io.StringIO(subprocess.run(cmd, capture_output=True, text=True, check=True).stdout)))
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
    start = starts[a.index]
    hdr = rows[start]
    data = []
    for r in rows[start + 1:]:
        if not r or r[0] in ("Kernel Name", "Address"):
            break
        data.append(r)
    kname = rows[start - 1][1] if start else "?"
    # sections of the disassembly: one per function; pick the one with len(data) instructions
    sections, cur, chain = collections.OrderedDict(), None, []
    pending = []
    for ln in dis.split("\n"):
        m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
        if m:
            cur = m.group(1)
            sections[cur] = []
            pending = []
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
        if m:
            pending.append((os.path.basename(m.group(1)), int(m.group(2))))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m and cur is not None:
            if pending:
                chain = pending
                pending = []
            sections[cur].append((m.group(2), chain))
    sec = [k for k, v in sections.items() if len(v) == len(data)]
    if not sec:
        sys.exit(f"no function with {len(data)} instructions in {cubin}: {[(k, len(v)) for k, v in sections.items()]}")
    insts = sections[sec[0]]
    ia, isamp, ith = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Thread Instructions Executed")
    agg = collections.defaultdict(lambda: [0, 0, 0, 0, 0])
    for (txt, ch), r in zip(insts, data):
        key = next((l for f, l in reversed(ch) if f == cu_name), 0) if ch else 0
        g = agg[key]
        ex = int(r[ia])
        g[0] += ex
        g[1] += int(r[isamp])
        g[2] += int(r[ith])
        g[3] += 1
        op = txt.split()[1] if txt.startswith("@") else txt.split()[0]
        if op.startswith(("DFMA", "DMUL", "DADD", "DSETP", "MUFU.RCP64H", "MUFU.RSQ64H")):
            g[4] += ex
    tot = sum(g[0] for g in agg.values())
    ts = sum(g[1] for g in agg.values())
    src = open(a.cu).read().split("\n")
    print(f"{kname}: {tot} warp instructions, {ts} samples, fp64-pipe share {100 * sum(g[4] for g in agg.values()) / tot:.1f}%")
    print("line  inst%  samp%  thr/inst fp64%  static | source")
    for k in sorted(agg):
        g = agg[k]
        if 100 * g[0] / tot >= a.min or 100 * g[1] / ts >= a.min:
            s = src[k - 1].strip()[:90] if k > 0 else "(no line)"
            print(f"{k:4d} {100 * g[0] / tot:6.1f} {100 * g[1] / ts:6.1f} {g[2] / max(g[0], 1):8.1f} {100 * g[4] / max(g[0], 1):5.0f} {g[3]:7d} | {s}")


if __name__ == "__main__":
    main()
