#!/usr/bin/env python
"""Turns the raw gpurun_out/<tag>_* artefacts of profiles/run_profiles.sh into the tracked summaries under profiles/:
   <tag>_<workload>_launches.md   per-kernel launch counts / device time / share (ncu gpu__time_duration, cold-cache, serialised)
   <tag>_<workload>_full.md       selected `ncu --set full` metrics of the top kernel + top stall sites
   roofline_traffic.json          dram bytes per launch of the dominant kernels (read by bench.py)
Usage: python profiles/summarise.py <tag> [traffic-key suffix, e.g. _f16 when the tag profiled the FP16-operand stress]"""
import csv
import json
import os
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")

FULL = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_bytes.sum', 'lts__t_sector_hit_rate.pct', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__ops_path_tensor_op_utchmma_src_tf32_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed',
        'sm__ops_path_tensor_op_utchmma_src_fp16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_tc.sum', 'sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'smsp__inst_executed.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__waves_per_multiprocessor', 'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_shared_mem',
        'launch__occupancy_limit_registers', 'sm__cycles_elapsed.max', 'smsp__cycles_active.avg']


def launches(tag, wl):
    p = os.path.join(OUT, f"{tag}_{wl}_launches.csv")
    if not os.path.exists(p):
        return None
    rows = [r for r in csv.reader(open(p)) if len(r) > 10 and r[0].isdigit()]
    agg = defaultdict(lambda: [0, 0.0, None, None])
    for r in rows:
        name = r[4].split('(')[0].replace('void ', '').replace('unnamed>::', '').replace('dsrl::<', '').strip()
        a = agg[name]
        a[0] += 1
        a[1] += float(r[-1].replace(',', ''))
        a[2], a[3] = r[7], r[8]
    tot = sum(a[1] for a in agg.values())
    lines = [f"# {tag} {wl}: kernel launches under `ncu --metrics gpu__time_duration.sum --clock-control none`", "",
             "Per-launch times under ncu are cold-cache and serialised: compare SHARES with the bench's own CUDA-event time, not absolutes.",
             f"Source: `gpurun_out/{tag}_{wl}_launches.csv` (command in `profiles/run_profiles.sh`).", "",
             "| kernel | launches | block | grid | total µs | mean µs | share |", "|---|---|---|---|---|---|---|"]
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"| `{name}` | {a[0]} | {a[2]} | {a[3]} | {a[1] / 1e3:.1f} | {a[1] / a[0] / 1e3:.2f} | {100 * a[1] / tot:.1f} % |")
    open(os.path.join(PROF, f"{tag}_{wl}_launches.md"), "w").write("\n".join(lines) + "\n")
    return agg


def full(tag, wl):
    rep = os.path.join(OUT, f"{tag}_{wl}_full.ncu-rep")
    if not os.path.exists(rep):
        return None
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rd = list(csv.reader(raw.splitlines()))
    hdr, units, row = rd[0], rd[1], rd[2]
    idx = {h: i for i, h in enumerate(hdr)}
    lines = [f"# {tag} {wl}: `ncu --set full --clock-control none` of `{row[idx['Kernel Name']].split('(')[0][:90]}`", "",
             f"Source: `gpurun_out/{tag}_{wl}_full.ncu-rep` (one launch; ncu replays the kernel ~40 times, durations are not bench numbers).", "",
             "| metric | value | unit |", "|---|---|---|"]
    vals = {}
    for m in FULL:
        if m in idx and row[idx[m]] != '':
            lines.append(f"| `{m}` | {row[idx[m]]} | {units[idx[m]]} |")
            vals[m] = (row[idx[m]], units[idx[m]])
    stalls = sorted(((float(row[i].replace(',', '') or 0), h) for h, i in idx.items()
                     if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('_per_issue_active.ratio') and row[i]), reverse=True)[:6]
    lines += ["", "Top warp stall reasons (warps per issue-active cycle):", ""]
    for v, h in stalls:
        lines.append(f"* {h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]}: {v:.2f}")
    src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    srows = list(csv.reader(src.splitlines()))
    if len(srows) > 3:
        sh = srows[1]
        si = {h: i for i, h in enumerate(sh)}
        data = [r for r in srows[2:] if len(r) == len(sh)]
        tot = sum(int(r[si['# Samples']] or 0) for r in data) or 1
        lines += ["", f"Hottest SASS sites by warp-state samples (total {tot}):", "", "| samples | share | executed | instruction | top stall |", "|---|---|---|---|---|"]
        for r in sorted(data, key=lambda r: -int(r[si['# Samples']] or 0))[:10]:
            st = sorted(((int(r[si[h]] or 0), h) for h in sh if h.startswith('stall_') and '(' not in h), reverse=True)[0]
            lines.append(f"| {r[si['# Samples']]} | {100 * int(r[si['# Samples']] or 0) / tot:.1f} % | {r[si['Instructions Executed']]} | `{r[si['Source']].strip()[:70]}` | {st[1]} |")
    open(os.path.join(PROF, f"{tag}_{wl}_full.md"), "w").write("\n".join(lines) + "\n")
    return vals


def main():
    tag = sys.argv[1]
    suffix = sys.argv[2] if len(sys.argv) > 2 else ""
    traffic_path = os.path.join(PROF, "roofline_traffic.json")
    traffic = json.load(open(traffic_path)) if os.path.exists(traffic_path) else {}
    for wl in ("fa_stress", "seg_counts", "fa_train", "fa_dsign", "fa_resolve", "fa_pack", "fa_finish"):
        launches(tag, wl)
        v = full(tag, wl)
        if v and 'dram__bytes_read.sum' in v:
            scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            rd = float(v['dram__bytes_read.sum'][0].replace(',', '')) * scale[v['dram__bytes_read.sum'][1]]
            wr = float(v['dram__bytes_write.sum'][0].replace(',', '')) * scale[v['dram__bytes_write.sum'][1]]
            traffic[wl + suffix] = rd + wr
            traffic[wl + suffix + "_source"] = f"profiles/{tag}_{wl}_full.md (dram__bytes_read.sum + dram__bytes_write.sum, one launch)"
    json.dump(traffic, open(traffic_path, "w"), indent=1)
    print(json.dumps(traffic, indent=1))


if __name__ == "__main__":
    main()
