#!/usr/bin/env python
"""Turns the scratch ncu outputs under gpurun_out/ into the committed summaries under profiles/ (run on the CPU box)."""
import collections, csv, json, os, shutil, subprocess, sys
tag = sys.argv[1] if len(sys.argv) > 1 else "r01_final"
shutil.copy("gpurun_out/launches_r01.csv", "profiles/%s_launches.csv" % tag)
for n in ("bench_r01.json", "bench_ref_r01.json"):
    if os.path.exists("gpurun_out/" + n):
        shutil.copy("gpurun_out/" + n, "profiles/%s_%s" % (tag, n.replace("_r01", "").replace("bench_ref", "bench_reference")))
for k in ("trace", "shade"):
    with open("profiles/%s_k_%s_ncu.txt" % (tag, k), "w") as f:
        f.write(subprocess.run([sys.executable, "profiles/ncu_summary.py", "gpurun_out/prof_%s_r01.ncu-rep" % k], capture_output=True, text=True).stdout)
rows = [r for r in csv.reader(open("profiles/%s_launches.csv" % tag)) if len(r) > 10]
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    name = r[ki].split("(")[0]; v = float(r[vi].replace(",", ""))
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for k, a in agg.items() if "at::" not in k)
out = ["# ncu launch list of `python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline` (%s)" % tag,
       "# gpu__time_duration.sum per launch, --clock-control none; shares are what matters (cold cache, serialised)", ""]
for k, a in agg.items():
    out.append("%-44s n=%4d total=%10.1f us  share=%.3f  avg=%.1f us" % (k[:44], a[0], a[1] / 1e3, a[1] / tot if "at::" not in k else 0, a[1] / a[0] / 1e3))
open("profiles/%s_launch_summary.txt" % tag, "w").write("\n".join(out) + "\n")
print("\n".join(out))
o = subprocess.run(["ncu", "-i", "gpurun_out/prof_trace_r01.ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(o.splitlines())); h = rr[0]
sc = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}
ur, uw = rr[1][h.index("dram__bytes_read.sum")], rr[1][h.index("dram__bytes_write.sum")]
rd = [float(r[h.index("dram__bytes_read.sum")]) * sc[ur] for r in rr[2:]]
wr = [float(r[h.index("dram__bytes_write.sum")]) * sc[uw] for r in rr[2:]]
dur = [float(r[h.index("gpu__time_duration.sum")]) for r in rr[2:]]
names = [r[h.index("Kernel Name")][:40] for r in rr[2:]]
per = [a + b for a, b in zip(rd, wr)]
json.dump({"kernel": "k_trace", "launches": names,
           "source": "ncu --set full --clock-control none -k regex:k_trace -s 10 -c 5 (= the five traversal launches of one wave: camera+primary, bounce 1..4) of `python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline`; summary profiles/%s_k_trace_ncu.txt" % tag,
           "dram_bytes_read_per_launch": rd, "dram_bytes_write_per_launch": wr, "duration_under_ncu": dur,
           "duration_unit": rr[1][h.index("gpu__time_duration.sum")], "dram_bytes_per_launch": sum(per) / len(per),
           "note": "average over the 5 launches of a wave; DRAM traffic is path-record and list I/O, the BVH itself is served by L1/L2"},
          open("profiles/trace_kernel_ncu.json", "w"), indent=1)
print(names, rd, wr, dur)
