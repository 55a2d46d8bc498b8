"""Per-source-line executed-instruction / stall-sample shares from `ncu --page source --csv --print-source cuda,sass`.
  ncu -i rep.ncu-rep --page source --csv --print-source cuda,sass --kernel-name regex:NAME > k.csv ; python tools/src_hist.py k.csv [N]"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1], errors='replace')))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur = None; hdr = None; inst = collections.Counter(); samp = collections.Counter(); text = {}
seen_kernel = 0
for r in rows:
  if not r: continue
  if r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
  if r[0] == 'Function Name': continue
  if r[0] == 'Line No': hdr = {h: i for i, h in enumerate(r)}; continue
  if hdr is None or r[0] == '': continue
  try:
    ln = int(r[0]); n = int(r[hdr['Instructions Executed']] or 0); s = int(r[hdr['# Samples']] or 0)
  except ValueError:
    continue
  inst[(cur, ln)] += n; samp[(cur, ln)] += s; text[(cur, ln)] = r[1]
ti, ts = sum(inst.values()), sum(samp.values())
print(f'total warp instructions {ti}, samples {ts}')
for k, n in inst.most_common(top):
  print(f'{100 * n / ti:5.1f}% inst {100 * samp[k] / max(ts, 1):5.1f}% samples  {k[0]}:{k[1]}: {text[k][:100]}')
