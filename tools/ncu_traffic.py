"""Not a test: summarise an ncu report of the aggregation launch into profiles/ and stamp profiles/traffic.json.
usage: python tools/ncu_traffic.py <report.ncu-rep> <workload> <round-tag>   (runs `ncu -i` here, no GPU needed)"""
import csv, io, json, os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import kernel_source_sha
rep, workload, tag = sys.argv[1:4]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__block_size", "launch__grid_size", "launch__shared_mem_per_block_dynamic",
        "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.max", "sm__cycles_active.avg"]
short = workload.split("_")[0]
out = ["# ncu --set full --clock-control none --import-source on: %s (%s)" % (workload, tag), "",
       "Captured with `tools/call_ncu.sh` (bench.py --workload %s --steps 1 --warmup 1 --no-cpu-baseline), kernel sources %s." % (workload, kernel_source_sha()), ""]
entry = None
for r in data:
    name = r[ix["Kernel Name"]]
    out += ["## " + name, "", "| metric | value |", "|---|---|"]
    for k in keys:
        if k in ix:
            out.append("| %s | %s %s |" % (k, r[ix[k]], units[ix[k]]))
    st = sorted(((h, float(r[i].replace(",", "") or 0)) for i, h in enumerate(hdr) if "smsp__average_warps_issue_stalled" in h and "not_issued" not in h and r[i]), key=lambda t: -t[1])
    out += ["", "Stall reasons (warps per issue-active cycle): " + ", ".join("%s %.2f" % (h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v) for h, v in st[:8]), ""]
    def val(k):
        v = float(r[ix[k]].replace(",", ""))
        return v * {"gbyte": 1e9, "mbyte": 1e6, "kbyte": 1e3, "byte": 1.0, "tbyte": 1e12}.get(units[ix[k]].lower(), 1.0)
    rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
    entry = {"dram_bytes_read": rd, "dram_bytes_write": wr, "traffic": rd + wr, "round": tag, "kernel_source_sha": kernel_source_sha(),
             "capture": "profiles/%s_ncu_summary_%s.md" % (tag, short)}
open(os.path.join(root, "profiles", "%s_ncu_summary_%s.md" % (tag, short)), "w").write("\n".join(out) + "\n")
tj = os.path.join(root, "profiles", "traffic.json")
t = json.load(open(tj)) if os.path.exists(tj) else {}
t.setdefault(workload, {})["mgm_aggregate_kernel_fused"] = entry
json.dump(t, open(tj, "w"), indent=1)
print("\n".join(out)[:2500])
