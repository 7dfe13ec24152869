"""Turn the outputs of scripts/profile_bench.sh (gpurun_out/<tag>_launches_bench.csv, <tag>_prof_sweep.ncu-rep) into the tracked
summaries under profiles/: a compact launch list, per-kernel metric tables, and profiles/spmv_traffic.json (DRAM bytes of the
sweep kernels per sweep, read by bench.py for roofline.traffic).      python scripts/summarize_ncu.py r02"""
import collections
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "rXX"
out_dir = os.path.join(ROOT, "profiles")
src_dir = os.path.join(ROOT, "gpurun_out")

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_red.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_op_read.sum", "lts__t_sectors_op_red.sum",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__cycles_elapsed.avg", "smsp__inst_executed_op_global_red.sum", "smsp__sass_inst_executed_op_shared_ld.sum"]


def short(name):
    return re.sub(r"\(.*", "", name).replace("void ", "").replace("unnamed>::", "").replace("b200::", "")


# ---- launch list
lpath = os.path.join(src_dir, f"{tag}_launches_bench.csv")
if os.path.exists(lpath):
    rows = [r for r in csv.reader(open(lpath)) if len(r) > 14 and r[0].isdigit()]
    d = collections.OrderedDict()
    for r in rows:
        d.setdefault((int(r[0]), short(r[4]), r[8]), {})[r[12]] = float(r[14].replace(",", ""))
    with open(os.path.join(out_dir, f"{tag}_launches_bench.csv"), "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:^k_ "
                "python bench.py --steps 1 --warmup 1   (scripts/profile_bench.sh; times are cold-cache, serialised)\n")
        f.write("id,kernel,grid,time_us,dram_read_MB,dram_write_MB\n")
        for (i, name, grid), m in d.items():
            f.write(f"{i},{name},\"{grid}\",{m.get('gpu__time_duration.sum', 0) / 1e3:.1f},{m.get('dram__bytes_read.sum', 0) / 1e6:.1f},"
                    f"{m.get('dram__bytes_write.sum', 0) / 1e6:.1f}\n")
    agg = collections.OrderedDict()
    for (i, name, grid), m in d.items():
        a = agg.setdefault(name.split("<")[0], [0, 0.0, 0.0])
        a[0] += 1
        a[1] += m.get("gpu__time_duration.sum", 0) / 1e3
        a[2] += m.get("dram__bytes_read.sum", 0) + m.get("dram__bytes_write.sum", 0)
    total = sum(a[1] for a in agg.values())
    print(f"{'kernel':34s} {'n':>5s} {'total us':>10s} {'avg us':>9s} {'share':>6s} {'DRAM MB/launch':>15s}")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:34s} {a[0]:5d} {a[1]:10.1f} {a[1] / a[0]:9.1f} {100 * a[1] / total:5.1f}% {a[2] / a[0] / 1e6:15.1f}")
    sweep = {k: a for k, a in agg.items() if k in ("k_sweep", "k_sweep_finish")}
    if sweep:
        js = {"dram_bytes_per_sweep": sum(a[2] / a[0] for a in sweep.values()),  # k_sweep + k_sweep_finish, per launch
              "per_kernel": {k: {"dram_bytes": a[2] / a[0], "time_us_under_ncu": a[1] / a[0], "launches_averaged": a[0]}
                             for k, a in sweep.items()},
              "source": f"profiles/{tag}_launches_bench.csv: ncu dram__bytes_read.sum + dram__bytes_write.sum per launch, averaged over "
                        f"the launches of `python bench.py --steps 1 --warmup 1` (the bench command itself, same workload)"}
        json.dump(js, open(os.path.join(out_dir, "spmv_traffic.json"), "w"), indent=1)
        print("spmv_traffic.json:", js["dram_bytes_per_sweep"] / 1e9, "GB per sweep")
else:
    print("no", lpath)

# ---- full-set captures
rep = os.path.join(src_dir, f"{tag}_prof_sweep.ncu-rep")
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = short(r[hdr.index("Kernel Name")]).split("<")[0]
        path = os.path.join(out_dir, f"{tag}_ncu_{name}.csv")
        with open(path, "w") as f:
            f.write(f"# ncu --set full --clock-control none --import-source on; {r[hdr.index('Kernel Name')][:90]}\n")
            f.write("metric,unit,value\n")
            for k, u, v in zip(hdr, units, r):
                if k in KEEP or ("average_warps_issue_stalled" in k and k.endswith("_per_issue_active.ratio") and "not_issued" not in k):
                    f.write(f"{k},{u},{v}\n")
        print("wrote", path)
else:
    print("no", rep)
