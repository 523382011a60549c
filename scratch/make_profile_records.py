"""profiles/<tag>_launch_shares.md and profiles/ncu_traffic.json from the records of one
GPU session (scratch/gpu_run18.sh): <tag>_launches.csv, <tag>_bench_n1.json,
<tag>_ncu_summary.txt, <tag>_traffic_nocachectl.csv."""
import csv, collections, json, os, re, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
P = lambda f: os.path.join(ROOT, "profiles", f)


def launches(fn):
    rows = list(csv.reader(open(fn)))
    for i, r in enumerate(rows):
        if "Kernel Name" in r:
            return r, rows[i + 1:]


hdr, rows = launches(P(tag + "_launches.csv"))
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(",", ""))
    agg.setdefault(r[ki].split("(")[0], []).append(v / 1e3 if r[ui] == "ns" else v)
sweep = [k for k in agg if "microbench" not in k and "thin" not in k]
tot = sum(sum(agg[k]) / len(agg[k]) for k in sweep)
out = ["# Round 2 (final build): kernel shares of one C4 sweep (ncu --metrics gpu__time_duration.sum --clock-control none, %s_launches.csv;" % tag,
       "# cold-cache serialised launch times: the SHARE is what compares with the CUDA-event stage times of the bench line)", "",
       "| kernel | launches | mean us | share of the sweep |", "|---|---|---|---|"]
for k in sweep:
    m = sum(agg[k]) / len(agg[k])
    out.append("| `%s` | %d | %.1f | %.1f %% |" % (k.replace("void ", ""), len(agg[k]), m, 100 * m / tot))
d = json.load(open(P(tag + "_bench_n1.json")))
r = d["roofline"]
ev = {"k_far_moments + k_sphere_columns_smem": r["kernels"]["k_sphere_columns"]["ms"], "k_nu_rates": r["kernels"]["k_nu_rates"]["ms"],
      "k_cell_solve_pce_spec": r["ms_cell_solve"], "k_converge_finish": r["ms_finish"]}
te = sum(ev.values())
out += ["", "CUDA-event stage times of the same build (bench line %s_bench_n1.json, ms_per_step %.4f):" % (tag, d["ms_per_step"]), ""]
out += ["* %s: %.4f ms = %.1f %%" % (k, v, 100 * v / te) for k, v in ev.items()]
open(P(tag + "_launch_shares.md"), "w").write("\n".join(out) + "\n")
print("\n".join(out[3:10]))

hdr, rows = launches(P(tag + "_traffic_nocachectl.csv"))
ki, mi, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tr = collections.OrderedDict()
for r in rows:
    if len(r) <= vi:
        continue
    n = r[ki].split("(")[0].replace("void ", "").split("<")[0]
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(r[ui], 1)
    tr.setdefault(n, {}).setdefault(r[mi], []).append(float(r[vi].replace(",", "")) * mult)
t = {n: {k: sum(v) / len(v) for k, v in m.items()} for n, m in tr.items()}
fl = {}
for blk in open(P(tag + "_ncu_summary.txt")).read().split("----"):
    m = re.search(r"Kernel Name\s+(?:void )?(\w+)", blk)
    rd = re.search(r"dram__bytes_read.sum\s+([\d.]+)", blk)
    wr = re.search(r"dram__bytes_write.sum\s+([\d.]+)", blk)
    if m and rd and wr:
        fl.setdefault(m.group(1), []).append((float(rd.group(1)) + float(wr.group(1))) * 1e6)
ren = lambda k: k.replace("k_sphere_columns_smem", "k_sphere_columns")
js = {"source": "profiles/%s_ncu_summary.txt: ncu --set full --clock-control none capture of `bench.py --steps 2 --warmup 1`, dram__bytes_read.sum + "
                "dram__bytes_write.sum per launch (the smaller of the two captured launches: the column kernel's first launch also writes back dirty "
                "lines of the bench's L2-flush scratch buffer), C4; ncu's default cache control flushes L2 before every replay pass" % tag,
      "bytes_per_launch": {ren(k): min(v) for k, v in fl.items() if k in ("k_nu_rates", "k_sphere_columns_smem", "k_far_moments")},
      "in_pipeline_source": "profiles/%s_traffic_nocachectl.csv: the same command with --cache-control none (caches as the previous kernel left "
                            "them, i.e. what the sweep sees): the 33.5 MB ray-record buffer written by the column kernel is read back from L2" % tag,
      "bytes_per_launch_in_pipeline": {ren(k): v["dram__bytes_read.sum"] + v["dram__bytes_write.sum"] for k, v in t.items()},
      "l2_bytes_per_launch_in_pipeline": {ren(k): v["lts__t_bytes.sum"] for k, v in t.items()}}
json.dump(js, open(P("ncu_traffic.json"), "w"), indent=1)
print(js["bytes_per_launch"], js["bytes_per_launch_in_pipeline"])
