#!/usr/bin/env python
"""gpurun_out/r02_* (written by tools/gpu_round2.sh, gpu8.sh on the GPU box) -> the tracked files under profiles/:
bench lines, the ncu launch list of the bench command, `ncu --set full` summaries (ncu_summary.py), the per-line
attribution of the bench-pass kernel (ncu_lines.py) and the SASS instruction histogram of the built library."""
import collections, glob, os, re, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
py = sys.executable

for f in glob.glob(os.path.join(G, "r02_bench*.json")) + [os.path.join(G, "r02_launches_bench.csv"), os.path.join(G, "r02_step_counters.json")]:
    if os.path.exists(f) and os.path.getsize(f):
        shutil.copy(f, P)
titles = {"step": "the bench pass (65 536 seeds -> one Panda target)", "speed": "Speed batch, 1 Mi Panda targets, <= 32 restarts (dynamic chains)",
          "quality": "Quality batch, 262 144 Panda targets x 32 restarts", "snake": "BASELINE config 4 (20-DOF snake, 262 144 seeds): thread-per-seed kernel, then the tile kernel",
          "single": "Robot::ik single calls (tile kernel, fused selection, mapped memory)"}
for w, t in titles.items():
    rep = os.path.join(G, f"r02_prof_{w}.ncu-rep")
    if os.path.exists(rep):
        out = subprocess.run([py, os.path.join(ROOT, "tools", "ncu_summary.py"), rep,
                              f"ncu --set full --clock-control none (tools/gpu_round2.sh -> tools/profile_t1.py {w}), round 2: {t}"],
                             capture_output=True, text=True, check=True).stdout
        open(os.path.join(P, f"r02_{w}_kernels_ncu.txt"), "w").write(out)
# evaluator: summary + DRAM traffic of the roofline launch (bench.py reads r02_eval_traffic.json)
rep = os.path.join(G, "r02_prof_eval.ncu-rep")
if os.path.exists(rep):
    import json
    out = subprocess.run([py, os.path.join(ROOT, "tools", "ncu_summary.py"), rep,
                          "ncu --set full --clock-control none (tools/gpu_round2.sh -> tools/profile_target.py eval), round 2: eval_kernel<7,true,128>, "
                          "4 Mi Panda configurations (the bench.py roofline launch)"], capture_output=True, text=True, check=True).stdout
    open(os.path.join(P, "r02_eval_kernel_ncu.txt"), "w").write(out)
    def val(key):
        m = re.search(re.escape(key) + r" \[(\w+)\] = ([0-9.]+)", out)
        return float(m.group(2)) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[m.group(1)]
    rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
    B = 1 << 22
    json.dump({"kernel": "eval_kernel<7,true,128>", "evals_per_launch": B, "dram_bytes_read": int(rd), "dram_bytes_write": int(wr),
               "traffic_bytes_per_launch": int(rd + wr), "algorithmic_bytes_per_launch": B * 8 * (8 * 7 + 17),
               "source": "ncu --set full --clock-control none, profiles/r02_eval_kernel_ncu.txt (dram__bytes_read.sum + dram__bytes_write.sum), "
                         "same launch shape as bench.py's roofline pass"}, open(os.path.join(P, "r02_eval_traffic.json"), "w"), indent=1)
rep = os.path.join(G, "r02_prof_step.ncu-rep")
if os.path.exists(rep):
    out = subprocess.run([py, os.path.join(ROOT, "tools", "ncu_lines.py"), rep, os.path.join(ROOT, "optik_b200", "csrc", "solve_t1_kernel.cu")],
                         capture_output=True, text=True)
    if out.returncode == 0:
        open(os.path.join(P, "r02_step_lines.txt"), "w").write(out.stdout)
    else:
        print("ncu_lines failed:", out.stderr[-400:])
# SASS histogram
so = os.path.join(ROOT, "optik_b200", "lib", "liboptik_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
hist, name = collections.OrderedDict(), None
for ln in sass.split("\n"):
    m = re.search(r"Function : (\S+)", ln)
    if m:
        name = m.group(1); hist[name] = collections.Counter(); continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", ln)
    if m and name:
        hist[name][m.group(1)] += 1
with open(os.path.join(P, "r02_sass_histogram.txt"), "w") as f:
    f.write("SASS instruction histogram of optik_b200/lib/liboptik_b200.so (cuobjdump -sass, sm_100a), round 2 (tools/refresh_profiles.py)\n"
            "evidence: UBLKCP = 1-D TMA bulk copy (chain blob / tiles / rows), SYNCS = mbarrier ops, DFMA/DMUL/DADD = fp64 pipe; "
            "no UTMALDG / UTCMMA (no tensor-map TMA, no tcgen05: by design)\n")
    tot = collections.Counter()
    for k, c in hist.items():
        f.write(f"--- {k}\n" + " ".join(f"{op}:{n}" for op, n in c.most_common()) + "\n")
        tot.update(c)
    f.write("--- whole library\n" + " ".join(f"{op}:{tot[op]}" for op in ("UBLKCP", "SYNCS", "DFMA", "DMUL", "DADD", "UTMALDG", "UTCMMA", "HMMA")) + "\n")
print("profiles refreshed:", sorted(os.path.basename(x) for x in glob.glob(os.path.join(P, "r02_*"))))
