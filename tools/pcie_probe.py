"""Pinned host<->device copy bandwidth of the box (the bound of bench.py's `e2e` leg: 1.5 B/px in, 3 B/px out per frame).
  python tools/pcie_probe.py  ->  one JSON line: H2D alone, D2H alone, both directions at once (GB/s)"""
import json
import torch

dev = torch.device('cuda:0')
n_in, n_out = 32 * 3840 * 2160 * 3 // 2, 32 * 3840 * 2160 * 3  # one bench step: 398 MB in, 796 MB out
h_in, h_out = torch.empty(n_in, dtype=torch.uint8).pin_memory(), torch.empty(n_out, dtype=torch.uint8).pin_memory()
d_in, d_out = torch.empty(n_in, dtype=torch.uint8, device=dev), torch.empty(n_out, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)


def timed(fn, reps=5):
  fn(); torch.cuda.synchronize()
  a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  a.record()
  for _ in range(reps):
    fn()
  for s in (s1, s2):
    torch.cuda.current_stream(dev).wait_stream(s)
  b.record(); torch.cuda.synchronize()
  return a.elapsed_time(b) / reps


def h2d():
  with torch.cuda.stream(s1):
    s1.wait_stream(torch.cuda.current_stream(dev)); d_in.copy_(h_in, non_blocking=True)


def d2h():
  with torch.cuda.stream(s2):
    s2.wait_stream(torch.cuda.current_stream(dev)); h_out.copy_(d_out, non_blocking=True)


def both():
  h2d(); d2h()


t_in, t_out, t_both = timed(h2d), timed(d2h), timed(both)
print(json.dumps({'h2d_gbs': round(n_in / t_in / 1e6, 1), 'd2h_gbs': round(n_out / t_out / 1e6, 1), 'both_ms': round(t_both, 3),
                  'both_d2h_gbs': round(n_out / t_both / 1e6, 1), 'e2e_bound_mp_per_s': round(32 * 3840 * 2160 / 1e6 / (t_both / 1e3), 1),
                  'note': 'e2e_bound = pixels of one bench step / time of its copies in both directions at once'}))
