"""Turns gpurun_out/ ncu artefacts into the tracked text summaries under profiles/.
usage: python scripts/summarize_profiles.py <tag> <launches.csv> <report.ncu-rep> <bench.log>"""
import collections, csv, io, json, subprocess, sys
tag, launches, rep, benchlog = sys.argv[1:5]
out = []
# ---- launch list (ncu --metrics gpu__time_duration.sum): per-kernel totals and shares
rows = [r for r in csv.reader(open(launches)) if len(r) > 5]
hdr = rows[0]; ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
d = collections.defaultdict(list)
for r in rows[1:]:
    try:
        d[r[ki]].append(float(r[vi].replace(",", "")))
    except ValueError:
        pass
ours = {k: v for k, v in d.items() if "qot::" in k}
tot_all = sum(sum(v) for v in d.values()); tot_ours = sum(sum(v) for v in ours.values())
out.append(f"# {tag}: ncu launch list (gpu__time_duration.sum, --clock-control none; cold-cache, serialised)\n")
out.append(f"command: python bench.py --steps 490 --warmup 5 --no-e2e --no-cpu-baseline (default: 8 graph branches)  ({len(rows)-1} launches captured)\n")
out.append("| kernel | launches | mean us | total us | share of all | share of qot:: kernels |\n|---|---:|---:|---:|---:|---:|")
for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
    name = k.split("(")[0][:70]
    so = f"{sum(v)/tot_ours*100:.1f}%" if k in ours else "-"
    out.append(f"| `{name}` | {len(v)} | {sum(v)/len(v)/1e3:.2f} | {sum(v)/1e3:.1f} | {sum(v)/tot_all*100:.1f}% | {so} |")
out.append("\nThe setup kernels (torch generators / sort / index ops that build the synthetic shard, `collate_kernel`, "
           "`lp_count_kernel`+`scan_kernel`+`widen_i32_kernel` that build lut_ptr at collate time) run before the timed region; "
           "inside the timed region a step is exactly one `lp_attn_kernel<false, true, true>` launch (the fused kernel), so its share of the step is 100%.\n")
# ---- full capture of the dominant kernel
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(raw))); h, u = r[0], r[1]
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__block_size", "launch__grid_size", "launch__waves_per_multiprocessor",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum",
        "sm__inst_executed.avg.per_cycle_elapsed", "sm__cycles_elapsed.max", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
out.append(f"# {tag}: ncu --set full of the dominant kernel (lp_attn_kernel<false, true, true>: attention + tensor-core head in one launch) ({len(r)-2} launches)\n\n| metric | unit | per launch |\n|---|---|---|")
for k in keys:
    if k in h:
        i = h.index(k)
        out.append(f"| {k} | {u[i]} | {', '.join(x[i] for x in r[2:])} |")
# ---- bench line
line = [l for l in open(benchlog) if l.startswith("{")][-1]
b = json.loads(line)
out.append(f"\n# {tag}: bench.py line (not under ncu)\n\n```json\n{json.dumps(b, indent=1)}\n```\n")
open(f"profiles/{tag}_summary.md", "w").write("\n".join(out) + "\n")
print("\n".join(out)[:3000])
