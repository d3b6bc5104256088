"""Turns gpurun_out/ ncu artefacts into the tracked text summaries under profiles/.
usage: python scripts/summarize_profiles.py <tag> <launches.csv> <report.ncu-rep> <bench.log> <batches_in_profiled_launch>
Writes profiles/<tag>_summary.md, profiles/lp_infer_traffic.json (read by bench.py for roofline.traffic) and
profiles/<tag>_sass_opcodes.txt (opcode histogram of lp_stream_kernel from the built .so)."""
import collections, csv, io, json, os, re, subprocess, sys, tempfile
tag, launches, rep, benchlog = sys.argv[1:5]
nb_prof = float(sys.argv[5]) if len(sys.argv) > 5 else None
out = []
rows = [r for r in csv.reader(open(launches)) if len(r) > 5]
hdr = rows[0]; ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
d = collections.defaultdict(list)
for r in rows[1:]:
    try:
        d[r[ki]].append(float(r[vi].replace(",", "")))
    except ValueError:
        pass
tot = sum(sum(v) for v in d.values())
out.append(f"# {tag}: ncu launch list of the TIMED REGION (gpu__time_duration.sum, --clock-control none; serialised)\n")
out.append("command: `QOT_PROFILE_TIMED_REGION=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none "
           "python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-secondary --min-timed-ms 5` "
           f"({len(rows) - 1} launches: bench.py brackets the timed replays with cudaProfilerStart/Stop)\n")
out.append("| kernel | launches | mean us | total us | share of the timed region |\n|---|---:|---:|---:|---:|")
for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
    out.append(f"| `{k.split('(')[0][:80]}` | {len(v)} | {sum(v) / len(v) / 1e3:.2f} | {sum(v) / 1e3:.1f} | {sum(v) / tot * 100:.1f}% |")
out.append("\nA step is one batch; one launch of `lp_stream_kernel` covers a run of consecutive batches (130 per launch in this "
           "command: a 260-step unit wraps the 245-batch shard once).  The only other launch is torch's 4-byte-per-batch fill that "
           "clears the status words.\n")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(raw))); h, u = r[0], r[1]
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__block_size",
        "launch__grid_size", "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum",
        "sm__inst_executed.avg.per_cycle_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__cycles_elapsed.max", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
out.append(f"# {tag}: ncu --set full of one launch of lp_stream_kernel<head = tcgen05, verified layout>\n\n| metric | unit | value |\n|---|---|---|")
val = {}
for k in keys:
    if k in h:
        i = h.index(k)
        out.append(f"| {k} | {u[i]} | {r[2][i]} |")
        val[k] = (float(r[2][i].replace(",", "")), u[i])
def to_bytes(k):
    v, unit = val[k]
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
if nb_prof:
    rd, wr = to_bytes("dram__bytes_read.sum") / nb_prof, to_bytes("dram__bytes_write.sum") / nb_prof
    json.dump({"kernel": "lp_stream_kernel<tcgen05 head, verified layout>", "batches_in_profiled_launch": nb_prof,
               "dram_bytes_read_per_batch": rd, "dram_bytes_write_per_batch": wr,
               "source": f"ncu --set full, {os.path.basename(rep)} ({tag})"}, open("profiles/lp_infer_traffic.json", "w"), indent=1)
    out.append(f"\nPer batch ({nb_prof:.0f} batches in the profiled launch): DRAM read {rd / 1e6:.3f} MB, write {wr / 1e6:.3f} MB "
               "(algorithmic: 6.74 MB) -> profiles/lp_infer_traffic.json.\n")
out.append(f"\n# {tag}: bench.py line (not under ncu)\n\n```json\n" +
           json.dumps(json.loads([l for l in open(benchlog) if l.startswith("{")][-1]), indent=1) + "\n```\n")
open(f"profiles/{tag}_summary.md", "w").write("\n".join(out))
# ---- SASS opcode histogram of the kernel
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath("gnn_qot_estimation_b200/libqot_b200.so")], cwd=tmp, check=True, capture_output=True)
cubin = [f for f in os.listdir(tmp) if "lightpath_stream" in f][0]
sass = subprocess.run(["nvdisasm", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
fn, hist = None, collections.defaultdict(collections.Counter)
for l in sass:
    if l.startswith(".text."):
        fn = l.strip()[6:]
    m = re.search(r"\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]+)", l)
    if m and fn:
        hist[fn][m.group(1)] += 1
with open(f"profiles/{tag}_sass_opcodes.txt", "w") as f:
    f.write("SASS opcode histogram (static counts, nvdisasm of gnn_qot_estimation_b200/libqot_b200.so, lightpath_stream.cu)\n"
            "Blackwell-native: UBLKCP = cp.async.bulk, UTCHMMA = tcgen05.mma, STTM / LDTM = tcgen05.st / ld, UTCBAR = tcgen05.commit,\n"
            "SYNCS.* = mbarrier, LDGSTS = cp.async, UTCATOMSWS = tcgen05.alloc / dealloc\n\n")
    for k, c in hist.items():
        if "lp_stream_kernel" in k or "st_head" in k or "st_producer" in k or "lp_wire" in k:
            key = [o for o in c if re.match(r"(UBLKCP|UTC|STTM|LDTM|SYNCS|LDGSTS|HMMA|UTMA)", o)]
            f.write(f"== {k}\n  native: " + ", ".join(f"{o} x{c[o]}" for o in sorted(key)) + "\n  top: " +
                    ", ".join(f"{o} x{n}" for o, n in c.most_common(14)) + f"\n  total {sum(c.values())}\n")
print(open(f"profiles/{tag}_summary.md").read()[:3000])
