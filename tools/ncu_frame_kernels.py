"""`ncu --set full` report of one frame's kernels -> profiles/ncu_frame_kernels.json (read by bench.py) + a CSV summary.

  python tools/ncu_frame_kernels.py gpurun_out/prof_X.ncu-rep profiles/r02_ncu_full_X_summary.csv

Per kernel and launch: duration, DRAM bytes, warp instructions, FP32 FLOPs (fadd + fmul + 2 * ffma thread instructions, predicated
on), XU (MUFU) thread operations (XU-pipe warp instructions x 32), and the utilisation of DRAM, the FMA / ALU / XU / LSU pipes and
the issue slots -- the numbers behind bench.py's per-stage `bound` and `frac_of_bound`.  The capture command is in
tools/gpu_r2_profile.sh (one 3840x2160 frame inside `bench.py --frames 2`, --clock-control none).
"""
import csv
import json
import subprocess
import sys

NAMES = {'rcd3_kernel': 'rcd_demosaic', 'rcd_strip_kernel': 'rcd_demosaic', 'smooth_kernel': 'color_smoothing', 'frame_stats_kernel': 'frame_stats',
         'prepare_kernel': 'frame_prepare', 'wiener32_kernel': 'wiener_tiles', 'wiener32_shared_kernel': 'wiener_tiles',
         'wiener_normalize_kernel': 'wiener_normalize_lum', 'wiener_normalize_lum4_kernel': 'wiener_normalize_lum', 'grid_build_kernel': 'bilateral_grid_build',
         'metrics_sliced_kernel': 'metrics_sliced', 'tonemap_kernel': 'bilateral_slice_tonemap'}
COLS = {
  'time_us': 'gpu__time_duration.sum', 'dram_read': 'dram__bytes_read.sum', 'dram_write': 'dram__bytes_write.sum',
  'dram_pct': 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm_pct': 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
  'warps_active_pct': 'sm__warps_active.avg.pct_of_peak_sustained_active', 'issue_pct': 'smsp__issue_active.avg.pct_of_peak_sustained_active',
  'warp_insts': 'smsp__inst_executed.sum', 'alu_pipe_pct': 'sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active',
  'fma_pipe_pct': 'sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active', 'lsu_pipe_pct': 'sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active',
  'xu_pipe_pct': 'sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active', 'xu_warp_insts': 'sm__inst_executed_pipe_xu.sum',
  'fadd': 'smsp__sass_thread_inst_executed_op_fadd_pred_on.sum', 'fmul': 'smsp__sass_thread_inst_executed_op_fmul_pred_on.sum',
  'ffma': 'smsp__sass_thread_inst_executed_op_ffma_pred_on.sum', 'regs': 'launch__registers_per_thread',
  'smem_dyn': 'launch__shared_mem_per_block_dynamic', 'grid': 'launch__grid_size', 'block': 'launch__block_size',
  'smem_conflicts': 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
  'smem_wavefront_pct': 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'l2_hit_pct': 'lts__t_sector_hit_rate.pct'}
SCALE = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'us': 1, 'ms': 1e3, 'ns': 1e-3, 's': 1e6, 'usecond': 1, 'msecond': 1e3, 'nsecond': 1e-3, 'second': 1e6}

rep, out_csv = sys.argv[1], sys.argv[2]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
ix = {k: (hdr.index(c) if c in hdr else -1) for k, c in COLS.items()}
kname = hdr.index('Kernel Name')


def val(r, key):
  i = ix[key]
  if i < 0 or r[i] in ('', 'n/a'):
    return None
  v = float(r[i].replace(',', ''))
  return v * SCALE.get(units[i], 1)


kernels = {}
with open(out_csv, 'w', newline='') as f:
  w = csv.writer(f)
  w.writerow(['kernel', 'launch name'] + list(COLS))
  for r in rows[2:]:
    name = next((n for k, n in NAMES.items() if k in r[kname]), None)
    v = {k: val(r, k) for k in COLS}
    w.writerow([name or '', r[kname][:80]] + [v[k] for k in COLS])
    if name is None or name in kernels:
      continue
    flops = None if v['ffma'] is None else (v['fadd'] or 0) + (v['fmul'] or 0) + 2 * v['ffma']
    kernels[name] = {
      'gpu_time_us': v['time_us'], 'dram_bytes': None if v['dram_read'] is None else int(v['dram_read'] + v['dram_write']),
      'warp_insts': v['warp_insts'], 'flops': flops, 'xu_thread_ops': None if v['xu_warp_insts'] is None else v['xu_warp_insts'] * 32,
      'dram_pct': v['dram_pct'], 'fma_pipe_pct': v['fma_pipe_pct'], 'alu_pipe_pct': v['alu_pipe_pct'], 'xu_pipe_pct': v['xu_pipe_pct'],
      'lsu_pipe_pct': v['lsu_pipe_pct'], 'issue_pct': v['issue_pct'], 'warps_active_pct': v['warps_active_pct'], 'registers': v['regs'],
      'smem_bank_conflicts': v['smem_conflicts'], 'smem_wavefront_pct': v['smem_wavefront_pct']}
doc = {'source': f'ncu --set full --clock-control none, one 3840x2160 frame inside bench.py --frames 2 ({out_csv}); all values per launch',
       'kernels': kernels}
json.dump(doc, open('profiles/ncu_frame_kernels.json', 'w'), indent=1)
total = sum(k['dram_bytes'] or 0 for k in kernels.values())
print(len(kernels), 'kernels,', round(total / 1e6, 1), 'MB DRAM per frame =', round(total / (3840 * 2160), 1), 'B/px')
for n, k in kernels.items():
  print(f"{n:26s} {k['gpu_time_us']:8.1f} us  dram {k['dram_pct']}%  fma {k['fma_pipe_pct']}%  xu {k['xu_pipe_pct']}%  issue {k['issue_pct']}%  flops {k['flops']}")
