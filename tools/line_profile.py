"""Scratch: warp instructions executed per source line of one kernel in an ncu report
(joins `ncu --page source --csv` SASS rows with `nvdisasm -g` line info by instruction order).
usage: line_profile.py report.ncu-rep lib.so kernel_symbol_substring [kernel-name-regex]"""
import csv, re, collections, subprocess, sys, os, tempfile
rep, so, sym = os.path.abspath(sys.argv[1]), os.path.abspath(sys.argv[2]), sys.argv[3]
kre = sys.argv[4] if len(sys.argv) > 4 else sym
tmp = tempfile.mkdtemp()
subprocess.run(f"cd {tmp} && cuobjdump -xelf all {so} > /dev/null 2>&1 && nvdisasm -g -c *.cubin > dis.txt 2>/dev/null", shell=True)
csvtxt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre], capture_output=True, text=True).stdout
lines = open(os.path.join(tmp, "dis.txt")).read().split("\n")
starts = [i for i, l in enumerate(lines) if l.startswith(".text.") and sym in l]
print("symbols:", [lines[i][:90] for i in starts])
start = starts[0]
seq, cur, stack = [], None, None
for l in lines[start + 1:]:
    if l.startswith("//---------------------"): break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)), m.group(3)); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s", l): seq.append((cur, l.strip()))
rows = list(csv.reader(csvtxt.splitlines()))
hi = [k for k, r in enumerate(rows) if "Instructions Executed" in r][0]
hdr = rows[hi]; ia = hdr.index("Instructions Executed"); iat = hdr.index("Avg. Threads Executed")
isamp = hdr.index("Warp Stall Sampling (All Samples)") if "Warp Stall Sampling (All Samples)" in hdr else None
data = []
for r in rows[hi + 1:]:  # the first launch only (the report repeats the table per launch)
    if r and r[0] == "Kernel Name":
        break
    if len(r) > ia and r[ia].isdigit():
        data.append(r)
print("sass rows", len(seq), len(data))
tot = sum(int(r[ia]) for r in data); print("total warp instr", tot)
agg, thr, smp = collections.Counter(), collections.Counter(), collections.Counter()
for k in range(min(len(seq), len(data))):
    c = seq[k][0]
    key = (c[0], c[1]) if c else ("?", 0)
    e = int(data[k][ia]); agg[key] += e
    try: thr[key] += float(data[k][iat]) * e
    except ValueError: pass
    if isamp is not None:
        try: smp[key] += int(data[k][isamp])
        except ValueError: pass
stot = max(1, sum(smp.values()))
if "--sass" in sys.argv:
    for k in range(min(len(seq), len(data))):
        print(data[k][ia], data[k][iat], seq[k][0][1] if seq[k][0] else 0, seq[k][1][:100])
else:
    for key in sorted(agg, key=lambda k: (k[0], k[1])):
        c = agg[key]
        if c * 1000 < tot and smp[key] * 1000 < stot: continue
        print(f"{key[0]}:{key[1]:5d} {100 * c / tot:5.1f}% inst  {100 * smp[key] / stot:5.1f}% stall-samples  avg thr {thr[key] / max(c, 1):4.1f}")
