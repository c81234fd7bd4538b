"""Per-kernel summary of an `ncu --set full` report: duration, DRAM bytes, issue utilisation, threads per instruction,
registers, local-memory instructions.  Writes a text table and (optionally) profiles/traffic.json.
    python tools/ncu_summary.py report.ncu-rep [out.txt] [traffic.json]
Launch durations under ncu are cold-cache and serialised (clocks not locked): shares, not absolutes, compare with bench.py."""
import csv, json, subprocess, sys, collections
rep = sys.argv[1]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr, units = rows[0], rows[1]
c = {n: i for i, n in enumerate(hdr)}
M = {"ms": "gpu__time_duration.sum", "rd": "dram__bytes_read.sum", "wr": "dram__bytes_write.sum",
     "thr": "smsp__thread_inst_executed_per_inst_executed.ratio", "issue": "smsp__issue_active.avg.pct_of_peak_sustained_active",
     "warps": "sm__warps_active.avg.pct_of_peak_sustained_active", "regs": "launch__registers_per_thread",
     "inst": "smsp__inst_executed.sum", "lld": "smsp__inst_executed_op_local_ld.sum", "lst": "smsp__inst_executed_op_local_st.sum",
     "l2hit": "lts__t_sector_hit_rate.pct", "dram_pct": "dram__throughput.avg.pct_of_peak_sustained_elapsed",
     "grid": "launch__grid_size", "block": "launch__block_size", "fma": "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
     "alu": "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
     "fp64": "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"}
def val(r, k):
    n = M[k]
    if n not in c or r[c[n]] in ("", "n/a"): return None
    v = float(r[c[n]].replace(",", ""))
    u = units[c[n]]
    if k == "ms": v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1.0)
    if k in ("rd", "wr"): v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
    return v
per = collections.OrderedDict()
for r in rows[2:]:
    name = r[c["Kernel Name"]].split("(")[0].replace("void ", "").replace("rcd::", "")
    d = {k: val(r, k) for k in M}
    per.setdefault(name, []).append(d)
lines = [f"# {rep}", "# kernel | launches in report | per launch (last frame captured): ms, DRAM read MB, DRAM write MB, issue-active %, "
         "threads/inst, warps-active %, regs, warp-inst (M), local ld+st inst (M), L2 hit %", ""]
summary = {}
tot = 0.0
for name, ds in per.items():
    d = ds[-1]
    tot += sum(x["ms"] for x in ds[len(ds) // 2:]) if len(ds) > 1 else d["ms"]
for name, ds in per.items():
    d = ds[-1]
    loc = ((d["lld"] or 0) + (d["lst"] or 0)) / 1e6
    lines.append(f"{name:28s} x{len(ds):<3d} {d['ms']:8.4f} ms  rd {d['rd'] / 1e6:8.2f} MB  wr {d['wr'] / 1e6:8.2f} MB  issue {d['issue']:5.1f}%  "
                 f"thr/inst {d['thr']:5.2f}  warps {d['warps']:5.1f}%  regs {int(d['regs']):3d}  inst {d['inst'] / 1e6:9.2f} M  local {loc:7.2f} M  "
                 f"L2 hit {d['l2hit']:5.1f}%  grid {int(d['grid'])}x{int(d['block'])}")
    summary[name] = {"ms_under_ncu": round(d["ms"], 5), "dram_read_bytes": d["rd"], "dram_write_bytes": d["wr"],
                     "issue_active_pct": round(d["issue"], 2), "threads_per_inst": round(d["thr"], 2), "warps_active_pct": round(d["warps"], 2),
                     "registers": int(d["regs"]), "warp_inst": d["inst"], "local_ldst_inst": (d["lld"] or 0) + (d["lst"] or 0),
                     "l2_hit_pct": round(d["l2hit"], 2), "launches_in_report": len(ds),
                     "fma_pipe_pct": d.get("fma"), "alu_pipe_pct": d.get("alu"), "fp64_pipe_pct": d.get("fp64")}
out = "\n".join(lines) + "\n"
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(out)
print(out)
if len(sys.argv) > 3:
    stage_of = {"k_pack_keys": "predict.keys", "k_onesweep_pass<1>": "predict.sort", "k_onesweep_pass<0>": "predict.sort",
                "k_reorder": "predict.reorder", "k_pairs<3, 0, 0>": "predict.pairs", "k_narrow<3, 0>": "predict.narrow",
                "k_exact<3, 2>": "predict.exact", "k_exact<3, 1>": "predict.exact"}
    tj = {"source": rep.split("/")[-1] + " (ncu --set full, one launch each; dram__bytes_read.sum + dram__bytes_write.sum; stages of several "
                                         "kernels: bytes summed, figures of the longest one)", "ncu": {}}
    longest = {}
    for name, key in stage_of.items():
        if name in summary:
            s = summary[name]
            tj[key] = tj.get(key, 0.0) + s["dram_read_bytes"] + s["dram_write_bytes"]
            if key not in longest or s["ms_under_ncu"] > longest[key]:
                longest[key] = s["ms_under_ncu"]
                tj["ncu"][key] = {k: s[k] for k in ("issue_active_pct", "threads_per_inst", "warps_active_pct", "registers", "local_ldst_inst",
                                                   "warp_inst", "fma_pipe_pct", "alu_pipe_pct", "fp64_pipe_pct")}
                tj["ncu"][key]["kernel"] = name
    json.dump(tj, open(sys.argv[3], "w"), indent=1)
