"""Turn gpurun_out/*.ncu-rep and launch lists into the tracked text summaries under profiles/.
usage: python scripts/make_profile_summaries.py TAG launches.csv stage.ncu-rep poisson.ncu-rep"""
import collections, csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, launches, stage_rep, poisson_rep = sys.argv[1:5]
out_dir = os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)

def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[0], rows[1], rows[2:]

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio"]

def kernel_table(rep, path, title):
    hdr, units, data = raw(rep)
    ki = hdr.index("Kernel Name")
    with open(path, "w") as f:
        f.write(f"# {title}\n# source: ncu --set full --clock-control none (one capture per kernel launch), file {os.path.basename(rep)}\n")
        for r in data:
            f.write(f"\n## {r[ki][:100]}\n")
            for k in KEYS:
                if k in hdr:
                    i = hdr.index(k)
                    f.write(f"{k:95s} {r[i]:>18s} {units[i]}\n")
    return hdr, units, data

# launch list
rows = list(csv.reader(open(launches)))
for n, r in enumerate(rows):
    if "Kernel Name" in r:
        hdr, start = r, n + 1
        break
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[start:]:
    if len(r) <= vi:
        continue
    name = r[ki].split("(")[0]
    agg.setdefault(name, [0, 0.0])
    agg[name][0] += 1
    agg[name][1] += float(r[vi].replace(",", ""))
tot = sum(v[1] for v in agg.values())
with open(os.path.join(out_dir, f"{tag}_launches_512.txt"), "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none: python scripts/profile_run.py 512 2 (set! + 2 time steps), 1 x B200\n")
    f.write("# cold-cache, serialised per-launch times: compare SHARES, not absolutes\n")
    for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
        f.write(f"{k:45s} n={v[0]:4d} total={v[1] / 1e6:9.3f} ms avg={v[1] / v[0] / 1e6:8.3f} ms share={v[1] / tot * 100:5.1f}%\n")
hdr2, units2, data2 = kernel_table(stage_rep, os.path.join(out_dir, f"{tag}_stage_kernel_512.txt"), "stage_kernel, 512^3, stages of a time step")
kernel_table(poisson_rep, os.path.join(out_dir, f"{tag}_poisson_projection_512.txt"), "Poisson solver + projection kernels, 512^3")
# DRAM traffic per launch of the stage kernel (mean over captured launches) for bench.py's roofline.traffic
def val(r, k):
    i = hdr2.index(k)
    v = float(r[i].replace(",", ""))
    u = units2[i].lower()
    return v * {"gbyte": 1e9, "mbyte": 1e6, "kbyte": 1e3, "byte": 1}.get(u, 1)
tr = [val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum") for r in data2]
json.dump({"size": 512, "n_gpus": 1, "dram_bytes_per_launch": sum(tr) / len(tr), "launches_captured": len(tr), "source": os.path.basename(stage_rep)},
          open(os.path.join(out_dir, "stage_kernel_traffic.json"), "w"))
print("wrote", os.listdir(out_dir))
