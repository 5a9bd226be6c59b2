"""Summarise an .ncu-rep (one kernel, `ncu --set full --import-source on`) into the two files kept under
profiles/:  python tools/ncu_extract.py gpurun_out/prof.ncu-rep profiles/ncu_rg_kernel_rNN"""
import csv, json, subprocess, sys
from collections import Counter

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
keep = ['gpu__time_duration.sum', 'sm__cycles_elapsed.avg', 'sm__cycles_elapsed.avg.per_second',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'sm__warps_active.avg.per_cycle_active',
        'launch__registers_per_thread', 'launch__block_size', 'launch__grid_size', 'launch__shared_mem_per_block_dynamic',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'lts__t_bytes.sum', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__t_sector_hit_rate.pct']
keep += [h for h in hdr if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio')]
res = {h: {"unit": u, "value": v} for h, u, v in zip(hdr, units, vals) if h in keep}
json.dump(res, open(out + "_metrics.json", "w"), indent=1, sort_keys=True)

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
h = next(i for i, r in enumerate(rows) if "Source" in r)
ix = {k: i for i, k in enumerate(rows[h])}
def num(x):
    try: return float(x)
    except ValueError: return 0.0
samples, execd = Counter(), Counter()
for r in rows[h + 1:]:
    toks = r[ix['Source']].strip().split()
    if not toks:
        continue
    op = (toks[1] if toks[0].startswith('@') and len(toks) > 1 else toks[0]).split('.')[0]
    samples[op] += num(r[ix['# Samples']]); execd[op] += num(r[ix['Instructions Executed']])
tot_s, tot_e = sum(samples.values()) or 1.0, sum(execd.values()) or 1.0
with open(out + "_source_summary.csv", "w") as f:
    f.write("opcode,warp_instructions_executed,share_of_executed,stall_samples,share_of_samples\n")
    for op, e in execd.most_common(25):
        f.write("%s,%d,%.4f,%d,%.4f\n" % (op, e, e / tot_e, samples[op], samples[op] / tot_s))
print(json.dumps({k: v["value"] for k, v in res.items() if "stalled" not in k}, indent=1))
