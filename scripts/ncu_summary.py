"""Build profiles/<tag>_fused_pass_summary.json from an ncu launch list (csv) and a full capture
(.ncu-rep) of `python bench.py --steps 1 --warmup 3 --no-cpu-baseline`.
   usage: ncu_summary.py TAG LAUNCHES.csv REPORT.ncu-rep [BENCH.json]"""
import csv, json, subprocess, sys, collections
tag, launches_csv, rep = sys.argv[1:4]
bench = json.loads(open(sys.argv[4]).readline()) if len(sys.argv) > 4 else None
rows = [r for r in csv.reader(open(launches_csv)) if len(r) > 10 and r[0].isdigit()]
per = collections.defaultdict(list)
for r in rows:
    per[r[4].split("(")[0].strip()].append(float(r[-1]) / 1e6)
tot = sum(sum(v) for v in per.values())
launch_list = {k: dict(launches=len(v), total_ms=round(sum(v), 3), mean_ms=round(sum(v) / len(v), 3), min_ms=round(min(v), 3),
                       max_ms=round(max(v), 3), share=round(sum(v) / tot, 4)) for k, v in per.items()}
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
hdr, units, data = rr[0], rr[1], rr[2:]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.avg.per_second"]
want += [h for h in hdr if "issue_stalled" in h and "per_issue_active" in h]
full = {}
for k in want:
    if k in hdr:
        i = hdr.index(k)
        full[k] = {"unit": units[i], "values": [d[i] for d in data]}
kname = data[0][hdr.index("Kernel Name")] if data else None
def gb(k):
    i = hdr.index(k)
    u = units[i].lower()
    f = {"gbyte": 1e9, "mbyte": 1e6, "kbyte": 1e3, "byte": 1.0, "tbyte": 1e12}[u]
    return [float(d[i]) * f for d in data]
dram = [a + b for a, b in zip(gb("dram__bytes_read.sum"), gb("dram__bytes_write.sum"))]
alg = 32.0 * (1 << 30)
out = {"round": tag, "what": "ncu captures of `python bench.py --steps 1 --warmup 3 --no-cpu-baseline` on one B200 (30 qubits, QFT-30 + 20 random layers); "
       "launch list = --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare shares, not absolutes); "
       "full capture = --set full --clock-control none --import-source on, 3 launches of the fused pass inside the timed step",
       "kernel": kname, "launch_list": launch_list,
       "kernel_share_of_step_ncu": max(v["share"] for v in launch_list.values()),
       "kernel_share_of_step_bench_events": bench["roofline"]["kernel_share_of_step"] if bench else None,
       "algorithmic_bytes_per_launch": alg, "dram_bytes_per_launch": dram, "dram_bytes_per_launch_mean": sum(dram) / len(dram),
       "traffic_over_algorithmic": sum(dram) / len(dram) / alg,
       "dram_bytes_per_launch_scaled_to": {str(L): sum(dram) / len(dram) * 2.0 ** (L - 30) for L in (29, 30, 31, 32)},
       "full_capture": full}
if bench:
    out["bench_line"] = {k: bench[k] for k in ("value", "ms_per_step", "e2e", "roofline", "clocks", "gpu_launches")}
json.dump(out, open(f"profiles/{tag}_fused_pass_summary.json", "w"), indent=1)
print("wrote", f"profiles/{tag}_fused_pass_summary.json", "traffic/alg", out["traffic_over_algorithmic"], launch_list)
