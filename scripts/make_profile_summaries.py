"""Turn the reduced ncu outputs of scripts/gpu_r2_profile.sh (raw-page CSVs + launch list) into the tracked text summaries under profiles/.
usage: python scripts/make_profile_summaries.py TAG [COMMIT]      (reads gpurun_out/TAG_*.csv)"""
import collections, csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
commit = sys.argv[2] if len(sys.argv) > 2 else subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
G = os.path.join(ROOT, "gpurun_out")
out_dir = os.path.join(ROOT, "profiles")

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]


def kernel_table(raw_csv, path, title):
    rows = list(csv.reader(open(raw_csv)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ki = hdr.index("Kernel Name")
    with open(path, "w") as f:
        f.write(f"# {title}\n# source: ncu --set full --clock-control none --import-source on (one capture per kernel launch), library of commit {commit},\n"
                f"# reduced on the GPU box to {os.path.basename(raw_csv)} (raw page) by scripts/gpu_r2_profile.sh\n")
        for r in data:
            f.write(f"\n## {r[ki][:110]}\n")
            for k in KEYS:
                if k in hdr:
                    i = hdr.index(k)
                    f.write(f"{k:95s} {r[i]:>18s} {units[i]}\n")
    return hdr, units, data


# launch list of the bench command
rows = list(csv.reader(open(os.path.join(G, f"{tag}_launches_bench.csv"))))
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
step_kernels = ("stage_kernel", "poisson_", "fft_x", "thomas_z", "remove_mean", "project_momentum", "halo_fill")
tot_step = sum(v[1] for k, v in agg.items() if any(s in k for s in step_kernels) and "stage_hi" not in k)
with open(os.path.join(out_dir, f"{tag}_launches_bench_512.txt"), "w") as f:
    f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none -c 400: python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-ensemble, 1 x B200, commit {commit}\n")
    f.write("# cold-cache, serialised per-launch times: compare SHARES, not absolutes. share = of all captured device time (incl. the sub-records' kernels);\n")
    f.write("# share_step = of the kernels of the 512^3 anelastic time step alone (stage, Poisson, projection, halo fills)\n")
    for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
        in_step = any(s in k for s in step_kernels) and "stage_hi" not in k
        f.write(f"{k[:60]:60s} n={v[0]:4d} total={v[1] / 1e6:9.3f} ms avg={v[1] / v[0] / 1e6:8.3f} ms share={v[1] / tot * 100:5.1f}%"
                + (f" share_step={v[1] / tot_step * 100:5.1f}%" if in_step else "") + "\n")
hdr2, units2, data2 = kernel_table(os.path.join(G, f"{tag}_stage_raw.csv"), os.path.join(out_dir, f"{tag}_stage_kernel_512.txt"), "stage_kernel, 512^3, the three stages of a time step")
kernel_table(os.path.join(G, f"{tag}_poisson_raw.csv"), os.path.join(out_dir, f"{tag}_poisson_projection_512.txt"), "Poisson solver + projection kernels, 512^3")
kernel_table(os.path.join(G, f"{tag}_stage_hi_raw.csv"), os.path.join(out_dir, f"{tag}_stage_hi_kernel_256.txt"), "specific_fields_kernel + stage_hi_kernel<5> (WENO9), 256^3")


def val(r, k):
    i = hdr2.index(k)
    v = float(r[i].replace(",", ""))
    u = units2[i].lower()
    return v * {"gbyte": 1e9, "mbyte": 1e6, "kbyte": 1e3, "byte": 1}.get(u, 1)


tr = [val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum") for r in data2]
json.dump({"size": 512, "n_gpus": 1, "dram_bytes_per_launch": sum(tr) / len(tr), "launches_captured": len(tr), "source": f"gpurun_out/{tag}_stage_raw.csv (ncu --set full, raw page)",
           "commit": commit}, open(os.path.join(out_dir, "stage_kernel_traffic.json"), "w"))
print("wrote summaries for", tag, "commit", commit)
