"""profiles/rNN_ncu_full_X_summary.csv -> profiles/ncu_traffic.json (DRAM bytes per launch of the frame kernels, read by bench.py).
  python tools/ncu_traffic.py profiles/r01_ncu_full_v7_summary.csv"""
import csv, json, sys
NAMES = {'rcd3_kernel': 'rcd_demosaic', 'smooth_kernel': 'color_smoothing', 'frame_stats_kernel': 'frame_stats', 'prepare_kernel': 'frame_prepare',
         'wiener32_kernel': 'wiener_tiles', 'wiener32_shared_kernel': 'wiener_tiles', 'wiener_normalize_kernel': 'wiener_normalize_lum', 'grid_build_kernel': 'bilateral_grid_build',
         'metrics_sliced_kernel': 'metrics_sliced', 'tonemap_kernel': 'bilateral_slice_tonemap'}
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
col = {n: hdr.index(n) for n in ('Kernel Name', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__time_duration.sum')}
scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
out = {}
for r in rows[2:]:
  for key, name in NAMES.items():
    if key in r[col['Kernel Name']]:
      out[name] = {'dram_read_bytes': int(float(r[col['dram__bytes_read.sum']]) * scale[units[col['dram__bytes_read.sum']]]),
                   'dram_write_bytes': int(float(r[col['dram__bytes_write.sum']]) * scale[units[col['dram__bytes_write.sum']]]),
                   'gpu_time_us': float(r[col['gpu__time_duration.sum']])}
doc = {'source': f'ncu --set full --clock-control none, one 3840x2160 frame inside bench.py --frames 2 ({sys.argv[1]}); bytes per launch', 'kernels': out}
json.dump(doc, open('profiles/ncu_traffic.json', 'w'), indent=1)
total = sum(v['dram_read_bytes'] + v['dram_write_bytes'] for v in out.values())
print(len(out), 'kernels,', total / 1e6, 'MB per frame =', total / (3840 * 2160), 'B/px')
