"""Dev probe: prints the headline metrics + hottest SASS sites of a .ncu-rep (ncu -i ... --page raw/source --csv)."""
import csv, subprocess, sys
rep = sys.argv[1]
M = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_sector_hit_rate.pct', 'lts__t_bytes.sum',
     'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
     'sm__ops_path_tensor_op_utchmma_src_fp16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed',
     'l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active',
     'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'launch__registers_per_thread', 'launch__waves_per_multiprocessor',
     'sm__cycles_elapsed.max', 'smsp__cycles_active.avg', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'launch__grid_size']
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rd = list(csv.reader(raw.splitlines()))
hdr, units, row = rd[0], rd[1], rd[2]
idx = {h: i for i, h in enumerate(hdr)}
print(row[idx['Kernel Name']][:80])
for m in M:
    if m in idx: print(f"  {m:100s} {row[idx[m]]} {units[idx[m]]}")
st = sorted(((float(row[i].replace(',', '') or 0), h) for h, i in idx.items() if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('_per_issue_active.ratio') and row[i]), reverse=True)[:6]
print("  stalls:", ", ".join(f"{h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]} {v:.2f}" for v, h in st))
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
sr = list(csv.reader(src.splitlines()))
if len(sr) > 3:
    sh = sr[1]; si = {h: i for i, h in enumerate(sh)}
    data = [r for r in sr[2:] if len(r) == len(sh)]
    tot = sum(int(r[si['# Samples']] or 0) for r in data) or 1
    for r in sorted(data, key=lambda r: -int(r[si['# Samples']] or 0))[:int(sys.argv[2]) if len(sys.argv) > 2 else 12]:
        stl = sorted(((int(r[si[h]] or 0), h) for h in sh if h.startswith('stall_') and '(' not in h), reverse=True)[0]
        print(f"  {100*int(r[si['# Samples']] or 0)/tot:5.1f}%  exec {r[si['Instructions Executed']]:>10s}  {r[si['Source']].strip()[:80]:80s} {stl[1]}")
