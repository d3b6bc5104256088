"""Developer tool: map ncu warp-sampling data (SASS level) of one kernel to source lines.
usage: python scripts/ncu_lines.py <report.ncu-rep> <cubin-name-substr> <kernel-mangled-substr> <source.cu> [top]"""
import collections, csv, io, re, subprocess, sys, tempfile, os
rep, cubin_sub, kern_sub, srcfile = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.environ.get("NCU_LINES_SO", os.path.abspath("gnn_qot_estimation_b200/libqot_b200.so"))], cwd=tmp, check=True, capture_output=True)
cubin = [f for f in os.listdir(tmp) if cubin_sub in f][0]
sass = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
starts = [i for i, l in enumerate(sass) if l.startswith(".text.")]
s0 = [i for i in starts if kern_sub in sass[i]][0]
s1 = min([i for i in starts if i > s0] + [len(sass)])
cur, lines = None, []
for l in sass[s0:s1]:
    m = re.search(r'//## File ".*?", line (\d+)', l)
    if m:
        cur = int(m.group(1)); continue
    if re.match(r"\s+/\*[0-9a-f]{4}\*/", l):
        lines.append(cur)
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", "0", "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, data = rows[1], rows[2:]
isamp, ie = hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
agg = collections.defaultdict(lambda: [0, 0, collections.Counter()])
for i in range(min(len(lines), len(data))):
    a = agg[lines[i]]
    a[0] += int(data[i][isamp]); a[1] += int(data[i][ie])
    for c in stall_cols:
        v = int(data[i][c] or 0)
        if v: a[2][hdr[c][6:]] += v
ts = sum(v[0] for v in agg.values()); te = sum(v[1] for v in agg.values())
src = open(srcfile).read().split("\n")
print(f"sass {len(data)} mapped {len(lines)}; samples {ts}; warp-instructions {te}")
for ln, (s, e, st) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    why = ",".join(f"{k}:{v}" for k, v in st.most_common(2))
    print(f"{ln or 0:4d} samp {s/ts*100:5.1f}% inst {e/te*100:5.1f}% [{why:28s}] {src[ln-1].strip()[:90] if ln else ''}")
