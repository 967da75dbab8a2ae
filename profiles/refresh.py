#!/usr/bin/env python
"""Turns the scratch outputs of tools/profile_r02.sh under gpurun_out/ into the committed evidence under profiles/
(run on the CPU box, where the .ncu-rep files are read with `ncu -i`):

    python profiles/refresh.py [tag, default r02]

  profiles/<tag>_bench_c<N>.json        one bench line per BASELINE config
  profiles/<tag>_launches.csv + _launch_summary.txt   ncu launch list of the default bench command, per-kernel shares
  profiles/<tag>_k_trace_ncu.txt, _k_shade_ncu.txt, _k_trace_c5_ncu.txt   ncu --set full summaries (profiles/ncu_summary.py,
                                        with the per-SASS-instruction stall hot spots)
  profiles/trace_kernel_ncu.json        dram bytes per k_trace launch (bench.py's roofline.traffic), per config
"""
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
G = "gpurun_out"


def summary(rep, out):
    if not os.path.exists(rep):
        return False
    with open(out, "w") as f:
        f.write("# %s -- python profiles/ncu_summary.py %s --source\n" % (os.path.basename(out), rep))
        f.write(subprocess.run([sys.executable, "profiles/ncu_summary.py", rep, "--source"], capture_output=True, text=True).stdout)
    return True


def dram_per_launch(rep):
    o = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(o.splitlines()))
    h = rr[0]
    sc = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}
    ur, uw = rr[1][h.index("dram__bytes_read.sum")], rr[1][h.index("dram__bytes_write.sum")]
    rd = [float(r[h.index("dram__bytes_read.sum")]) * sc[ur] for r in rr[2:]]
    wr = [float(r[h.index("dram__bytes_write.sum")]) * sc[uw] for r in rr[2:]]
    dur = [float(r[h.index("gpu__time_duration.sum")]) for r in rr[2:]]
    names = [r[h.index("Kernel Name")][:40] for r in rr[2:]]
    per = [a + b for a, b in zip(rd, wr)]
    return {"launches": names, "dram_bytes_read_per_launch": rd, "dram_bytes_write_per_launch": wr,
            "duration_under_ncu": dur, "duration_unit": rr[1][h.index("gpu__time_duration.sum")],
            "dram_bytes_per_launch": sum(per) / len(per)}


for c in (1, 2, 3, 4, 5):
    src = "%s/%s_bench_c%d.json" % (G, tag, c)
    if os.path.exists(src) and os.path.getsize(src) > 10:
        line = [l for l in open(src).read().splitlines() if l.startswith("{")]
        if line:
            open("profiles/%s_bench_c%d.json" % (tag, c), "w").write(line[-1] + "\n")
            d = json.loads(line[-1])
            print("config %d: value %.1f  e2e %s  roofline %s %.3f  parity %s" % (
                c, d["value"], d.get("e2e", {}).get("value"), d["roofline"]["bound"], d["roofline"]["frac"], d.get("parity")))

src = "%s/%s_launches.csv" % (G, tag)
if os.path.exists(src):
    shutil.copy(src, "profiles/%s_launches.csv" % tag)
    rows = [r for r in csv.reader(open(src)) if len(r) > 10]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        name = r[ki].split("(")[0]
        v = float(r[vi].replace(",", ""))
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    ours = lambda k: "at::" not in k and "nccl" not in k.lower() and "k_read_bw" not in k  # noqa: E731
    tot = sum(a[1] for k, a in agg.items() if ours(k))
    out = ["# ncu launch list of `python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-verify` (%s)" % tag,
           "# gpu__time_duration.sum per launch, --clock-control none; shares are what matters (cold cache, serialised);",
           "# k_read_bw = bench.py's bandwidth probe after the timed region, at:: = the L2 flush between steps", ""]
    for k, a in agg.items():
        out.append("%-44s n=%4d total=%10.1f us  share=%.3f  avg=%.1f us" % (k[:44], a[0], a[1] / 1e3, a[1] / tot if ours(k) else 0, a[1] / a[0] / 1e3))
    open("profiles/%s_launch_summary.txt" % tag, "w").write("\n".join(out) + "\n")
    print("\n".join(out))

summary("%s/%s_k_trace.ncu-rep" % (G, tag), "profiles/%s_k_trace_ncu.txt" % tag)
summary("%s/%s_k_shade.ncu-rep" % (G, tag), "profiles/%s_k_shade_ncu.txt" % tag)
summary("%s/%s_k_trace_c5.ncu-rep" % (G, tag), "profiles/%s_k_trace_c5_ncu.txt" % tag)

per_config = {}
for c, rep in ((2, "%s/%s_k_trace.ncu-rep" % (G, tag)), (5, "%s/%s_k_trace_c5.ncu-rep" % (G, tag))):
    if os.path.exists(rep):
        per_config[str(c)] = dram_per_launch(rep)
if per_config:
    old = {}
    if os.path.exists("profiles/trace_kernel_ncu.json"):
        old = json.load(open("profiles/trace_kernel_ncu.json"))
    pc = old.get("per_config", {})
    pc.update(per_config)
    doc = {"kernel": "k_trace",
           "source": "ncu --set full --clock-control none -k regex:k_trace (tools/profile_r02.sh) of `python bench.py --config N --steps 2 "
                     "--warmup 3 --no-e2e --no-cpu-baseline --no-verify`; summaries profiles/%s_k_trace*_ncu.txt" % tag,
           "note": "per launch, averaged over the captured launches; config 2: DRAM traffic is path-record and list I/O, the BVH is "
                   "served by L1/L2; config 5 (10 M triangles): the BVH itself streams from HBM",
           "per_config": pc}
    if "2" in pc:
        doc["dram_bytes_per_launch"] = pc["2"]["dram_bytes_per_launch"]
    json.dump(doc, open("profiles/trace_kernel_ncu.json", "w"), indent=1)
    print({k: v["dram_bytes_per_launch"] for k, v in pc.items()})
