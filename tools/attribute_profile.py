"""Scratch: attribute executed warp instructions of a kernel in an ncu report to source regions
(joins `ncu --page source --csv` SASS rows with `nvdisasm -g` line info by instruction order)."""
import csv, re, collections, subprocess, sys, os, tempfile
rep, so, kernel_sym = os.path.abspath(sys.argv[1]), os.path.abspath(sys.argv[2]), sys.argv[3]
tmp = tempfile.mkdtemp()
subprocess.run(f"cd {tmp} && cuobjdump -xelf all {so} > /dev/null 2>&1 && nvdisasm -g -c *.cubin > dis.txt 2>/dev/null", shell=True)
csvtxt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = open(os.path.join(tmp, "dis.txt")).read().split("\n")
start = [i for i, l in enumerate(lines) if l.startswith(f".text.{kernel_sym}:")][0]
seq, cur = [], None
for l in lines[start + 1:]:
    if l.startswith("//---------------------"): break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', l)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s", l): seq.append(cur)
rows = list(csv.reader(csvtxt.splitlines()))
hdr = rows[1]; ia = hdr.index("Instructions Executed"); iat = hdr.index("Avg. Threads Executed")
data = [r for r in rows[2:] if len(r) > ia and r[ia].isdigit()]
print("sass rows", len(seq), len(data))
tot = sum(int(r[ia]) for r in data); print("total warp instr", tot)
src = open("realtime-collision-detection_b200/csrc/rcd_pairs.cuh").read().split("\n")
pats = {"predict_coef": "PredictCoef predict_coef(", "offset_may_hit": "bool offset_may_hit(", "g2_at": "float g2_at(",
        "predict_window": "bool predict_window(", "predict_scan": "u32 predict_scan(", "sample_predict": "u32 sample_predict(",
        "narrow_compute_node": "bool narrow_compute_node(", "lower_bound/global_push": "u32 lower_bound_keys(",
        "kernel prologue": "__global__ void __launch_bounds__(PAIR_THREADS", "run_scan (S2b)": "auto run_scan = [&]",
        "groups/rows/spans": "// a tile that crosses a cell-row boundary", "stage": "auto stage = [&](u32 c)",
        "S1 filter": "// ---- S1 filter: one query per lane", "list scan": "// exclusive scan of the list lengths",
        "S2 dispatch (S2a)": "// ---- S2 on up to 32 pairs", "end of tile": "// ---- end of tile",
        "narrow_detect": "bool narrow_detect(", "exact/emit": "struct EmitRec {", "helpers": "void cp_async16("}
marks = sorted(((n, [i + 1 for i, l in enumerate(src) if p in l][0]) for n, p in pats.items()), key=lambda x: x[1])
def region(fn, ln):
    if fn != "rcd_pairs.cuh": return fn
    name = "top"
    for n, l in marks:
        if ln >= l: name = n
    return name
agg, thr = collections.Counter(), collections.Counter()
for k in range(min(len(seq), len(data))):
    fn, ln = seq[k] if seq[k] else ("?", 0)
    r = region(fn, ln); e = int(data[k][ia]); agg[r] += e
    try: thr[r] += float(data[k][iat]) * e
    except ValueError: pass
for r, c in agg.most_common():
    print(f"{100 * c / tot:5.1f}%  {c:>13}  avg threads {thr[r] / max(c, 1):4.1f}  {r}")
