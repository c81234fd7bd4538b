"""Scratch: print the SASS rows (with executed-instruction counts) of one source region of a kernel in an ncu report.
usage: sass_region.py report.ncu-rep lib.so kernel_symbol first_line last_line [file]"""
import csv, re, subprocess, sys, os, tempfile
rep, so, kernel_sym = os.path.abspath(sys.argv[1]), os.path.abspath(sys.argv[2]), sys.argv[3]
l0, l1 = int(sys.argv[4]), int(sys.argv[5])
fname = sys.argv[6] if len(sys.argv) > 6 else "rcd_pairs.cuh"
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
hdr = rows[1]; ia = hdr.index("Instructions Executed"); isrc = hdr.index("Source"); ist = hdr.index("# Samples")
data = [r for r in rows[2:] if len(r) > ia and r[ia].isdigit()]
tot = 0
for k in range(min(len(seq), len(data))):
    fn, ln = seq[k] if seq[k] else ("?", 0)
    if fn == fname and l0 <= ln <= l1:
        tot += int(data[k][ia])
        print(f"{k:5d} L{ln:4d} {int(data[k][ia]):>12} smp {data[k][ist]:>6}  {data[k][isrc].strip()}")
print("total", tot)
