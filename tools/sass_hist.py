"""Summarise an `ncu --page source --csv` export: executed-instruction histogram by opcode and stall-reason totals.
  ncu -i rep.ncu-rep --page source --csv --kernel-name regex:NAME > k.csv ; python tools/sass_hist.py k.csv"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1], errors='replace')))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
ops = collections.Counter(); static = collections.Counter(); stalls = collections.Counter(); samples = 0
total = 0
for r in rows[2:]:
  if r and r[0] == 'Kernel Name': break  # only the first profiled instance
  if len(r) < len(hdr) - 5 or r[0] == 'Address': continue
  src = r[ix['Source']].strip()
  toks = src.split()
  if toks and toks[0].startswith('@'): toks = toks[1:]
  op = toks[0].split('.')[0] if toks else '?'
  full = toks[0] if toks else '?'
  n = int(r[ix['Instructions Executed']] or 0)
  ops[full if op in ('LDS', 'STS', 'LDG', 'STG', 'RED', 'REDG', 'ATOMG', 'LDL', 'STL', 'SHFL', 'MUFU') else op] += n
  static[op] += 1
  total += n
  for h in hdr:
    if h.startswith('stall_') and 'Not Issued' not in h:
      stalls[h] += int(r[ix[h]] or 0)
print('static instructions', sum(static.values()), 'executed warp-instr', total)
for k, v in ops.most_common(28): print(f'  {k:28s} {v:12d} {100*v/total:5.1f}%')
ts = sum(stalls.values())
print('stall samples', ts)
for k, v in stalls.most_common(10): print(f'  {k:28s} {v:8d} {100*v/ts:5.1f}%')
