"""Not a test: summarise an `ncu --page raw --csv` export into a markdown table per kernel launch and update
profiles/traffic.json.  usage: ncu_summary.py <raw.csv> <out.md> <workload> <command line used>"""
import csv, json, os, sys
raw, out_md, workload, cmd = sys.argv[1:5]
rows = list(csv.reader(open(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__block_size", "launch__grid_size", "launch__shared_mem_per_block_dynamic",
        "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
        "smsp__average_warp_latency_per_inst_issued.ratio", "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
lines = ["# ncu --set full --clock-control none summary (B200)", "", "Command: `%s`." % cmd, ""]
traffic = {}
for r in data:
    name = r[ix["Kernel Name"]]
    lines += ["## " + name, "", "| metric | value |", "|---|---|"]
    for k in keys:
        if k in ix:
            lines.append("| %s | %s %s |" % (k, r[ix[k]], units[ix[k]]))
    lines.append("")
    short = "mgm_aggregate_kernel" if "mgm_aggregate" in name else ("mgm_wta_kernel" if "mgm_wta" in name else None)
    if short:
        def val(k):
            v = float(r[ix[k]].replace(",", ""))
            u = units[ix[k]].lower()
            return v * {"gbyte": 1e9, "mbyte": 1e6, "kbyte": 1e3, "byte": 1.0, "tbyte": 1e12}.get(u, 1.0)
        rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
        if short == "mgm_aggregate_kernel" and rd > 1.5 * wr:   # the launch with the finish tiles fused in re-reads the sweeps
            short = "mgm_aggregate_kernel_fused"
        traffic[short] = {"dram_bytes_read": rd, "dram_bytes_write": wr, "traffic": rd + wr, "round": 1}
open(out_md, "w").write("\n".join(lines))
tj = os.path.join(os.path.dirname(out_md), "traffic.json")
t = json.load(open(tj)) if os.path.exists(tj) else {}
t.setdefault(workload, {}).update(traffic)
json.dump(t, open(tj, "w"), indent=1)
print(open(out_md).read()[:3000])
