"""Summarise an `ncu --set full` report (exported with `ncu -i X.ncu-rep --page raw --csv`) into
profiles/ncu_summary_rNN.json + .md: per kernel, averaged over its captured launches."""
import csv, json, sys, collections

raw, out_json, out_md = sys.argv[1], sys.argv[2], sys.argv[3]
rows = list(csv.reader(open(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
M = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
     "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
     "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
     "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
     "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "smsp__cycles_elapsed.avg.per_second",
     "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__t_output_wavefronts_pipe_lsu_mem_local_op_ld.sum",
     "l1tex__t_output_wavefronts_pipe_lsu_mem_local_op_st.sum", "smsp__thread_inst_executed_per_inst_executed.ratio"]
scale = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1, "Ghz": 1e9, "Mhz": 1e6}
acc = collections.OrderedDict()
for r in rows[2:]:
    name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "").strip()
    d = acc.setdefault(name, collections.defaultdict(list))
    for m in M:
        if m in idx and r[idx[m]] != "":
            try:
                d[m].append(float(r[idx[m]].replace(",", "")) * scale.get(units[idx[m]], 1))
            except ValueError:
                pass
summary = {}
for name, d in acc.items():
    s = {m: sum(v) / len(v) for m, v in d.items()}
    s["launches_captured"] = len(d["gpu__time_duration.sum"])
    s["dram_bytes_per_launch"] = s.get("dram__bytes_read.sum", 0) + s.get("dram__bytes_write.sum", 0)
    summary[name] = s
# bench.py looks traffic up by mode name
alias = {"local": "k_local_s16<1>", "semi_global": "k_semiglobal_s16<1>", "pack": "k_pack"}
for k, v in alias.items():
    if v in summary:
        summary[k] = {"kernel": v, "dram_bytes_per_launch": summary[v]["dram_bytes_per_launch"], "note": "per launch of the ncu capture (262144 pairs 150x600)"}
json.dump(summary, open(out_json, "w"), indent=1)
with open(out_md, "w") as f:
    f.write("| kernel | launches | time ms | DRAM rd MB | DRAM wr MB | ALU pipe % | FMA pipe % | issue active % | warps active % | regs | SM GHz | smem bank conflicts | local-mem wavefronts |\n|---|---|---|---|---|---|---|---|---|---|---|---|---|\n")
    for name, s in summary.items():
        if "gpu__time_duration.sum" not in s:
            continue
        g = lambda m, sc=1.0: ("%.3g" % (s[m] * sc)) if m in s else "-"
        f.write(f"| {name} | {s['launches_captured']} | {g('gpu__time_duration.sum', 1e3)} | {g('dram__bytes_read.sum', 1e-6)} | {g('dram__bytes_write.sum', 1e-6)} | "
                f"{g('sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active')} | {g('sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active')} | "
                f"{g('smsp__issue_active.avg.pct_of_peak_sustained_active')} | {g('sm__warps_active.avg.pct_of_peak_sustained_active')} | {g('launch__registers_per_thread')} | "
                f"{g('smsp__cycles_elapsed.avg.per_second', 1e-9)} | {g('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum')} | "
                f"{('%.3g' % (s.get('l1tex__t_output_wavefronts_pipe_lsu_mem_local_op_ld.sum', 0) + s.get('l1tex__t_output_wavefronts_pipe_lsu_mem_local_op_st.sum', 0)))} |\n")
print(open(out_md).read())
