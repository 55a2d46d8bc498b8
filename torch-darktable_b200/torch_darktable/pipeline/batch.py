"""Streaming frames through an ImageProcessor from HOST memory (B200 addition).

The reference processes one image set at a time and synchronises the device after most ops, so host<->device copies
never overlap its compute.  HostFrameRunner keeps three CUDA streams busy instead: while frame i is processed, frame i+1
is copied in from pinned memory and the uint8 result of frame i-1 is copied out.  Each frame is its own image set
(`ImageProcessor.process`), exactly what a caller of the reference does per camera frame.
"""

from __future__ import annotations

import os
from pathlib import Path

import torch

from .image_processor import ImageProcessor


def _parse_cpulist(text: str) -> set[int]:
  cpus: set[int] = set()
  for part in text.strip().split(','):
    if not part:
      continue
    lo, _, hi = part.partition('-')
    cpus.update(range(int(lo), int(hi or lo) + 1))
  return cpus


def gpu_locality(index: int) -> dict:
  """NUMA node and local CPUs of CUDA device `index`, from sysfs through its PCI address ({} fields are None when the platform does
  not say, e.g. a VM without NUMA topology)."""
  info = {'numa_node': None, 'local_cpus': None, 'pci': None}
  try:
    p = torch.cuda.get_device_properties(index)
    pci = f'{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0'
    info['pci'] = pci
    base = Path('/sys/bus/pci/devices') / pci
    node = int((base / 'numa_node').read_text())
    info['numa_node'] = node if node >= 0 else None
    cpus = (base / 'local_cpulist').read_text().strip()
    info['local_cpus'] = cpus or None
  except (OSError, ValueError, AttributeError, RuntimeError):
    pass
  return info


def bind_host_to_gpu(index: int) -> dict:
  """Restrict the calling thread (and the threads it starts) to the CPUs that are local to CUDA device `index`, so that the pinned
  buffers it allocates afterwards are placed (first touch) on the GPU's own NUMA node and its copies do not cross the socket
  interconnect.  With eight ranks streaming 43 GB/s each through pinned memory (bench.py's e2e leg) unbound ranks all land on one
  node's memory controllers; see tools/pcie_probe.py for the measurement.  Call it before allocating pinned memory.  Returns what it
  did; never raises (a container may forbid the affinity call or hide the topology)."""
  info = {'bound': False, **gpu_locality(index)}
  try:
    allowed = os.sched_getaffinity(0)
    local = _parse_cpulist(info['local_cpus']) & allowed if info['local_cpus'] else set()
    if local and local != allowed:
      os.sched_setaffinity(0, local)
      info['bound'] = True
    info['cpus_used'] = len(local) if local else len(allowed)
  except (OSError, AttributeError, ValueError):
    pass
  return info


class HostFrameRunner:
  def __init__(self, processor: ImageProcessor, slots: int = ImageProcessor.LANES + 1):
    self.processor = processor
    self.device = processor.device
    self.slots = slots
    self._in = [torch.empty(processor.expected_bytes, dtype=torch.uint8, device=self.device) for _ in range(slots)]
    self._s_in = torch.cuda.Stream(self.device)
    self._s_compute = torch.cuda.Stream(self.device)
    self._s_out = torch.cuda.Stream(self.device)
    self._copied = [torch.cuda.Event() for _ in range(slots)]
    self._consumed = [torch.cuda.Event() for _ in range(slots)]
    self._ready = [torch.cuda.Event() for _ in range(slots)]
    self._out_done = [torch.cuda.Event() for _ in range(slots)]
    self._results: list[torch.Tensor | None] = [None] * slots  # device results whose copy-out may still be in flight
    self._done = torch.cuda.Event()
    self._used = [False] * slots  # slot has held a frame before (its events are recorded)
    self._next = 0                # slot of the next frame: batches continue round robin across calls

  def run(self, host_frames: list[torch.Tensor], host_out: list[torch.Tensor], name: str = 'cam', after_caller: bool = True) -> None:
    """host_frames: pinned uint8 packed frames; host_out: pinned uint8 (H', W', 3) buffers that receive the results.
    Returns after everything has been enqueued; call `wait()` (or synchronise the device) before reading host_out.

    after_caller: order the three streams after the work already queued on the caller's current stream (needed when that work
    touches the processor).  A caller that only streams batches through the runner passes False: consecutive `run` calls then
    pipeline into each other (the copy-in of the next batch overlaps the tail of this one) -- the slots are handed over by events
    that outlive a call, and the caller's stream is still ordered after the end of every batch.

    A result tensor stays referenced in its slot until the compute stream has been ordered after its copy-out, so that the
    caching allocator hands its memory out again in plain stream order (no record_stream bookkeeping, no allocator growth)."""
    assert len(host_frames) == len(host_out)
    caller = torch.cuda.current_stream(self.device)
    if after_caller:
      for s in (self._s_in, self._s_compute, self._s_out):
        s.wait_stream(caller)
    for i, frame in enumerate(host_frames):
      slot = (self._next + i) % self.slots
      with torch.cuda.stream(self._s_in):
        if self._used[slot]:
          self._s_in.wait_event(self._consumed[slot])  # the previous user of this slot (this call's or an earlier one's) has been processed
        self._in[slot].copy_(frame, non_blocking=True)
        self._copied[slot].record(self._s_in)
      with torch.cuda.stream(self._s_compute):
        self._s_compute.wait_event(self._copied[slot])
        if self._results[slot] is not None:
          self._s_compute.wait_event(self._out_done[slot])  # the copy-out of the result this slot held has finished
          self._results[slot] = None
        # two frames in flight on the processor's lanes; `done` marks both "input consumed" and "result ready"
        result, done = self.processor.submit(self._in[slot], name, track=False)  # slots and results are kept alive here until their copies are ordered
        self._consumed[slot] = self._ready[slot] = done
        self._used[slot] = True
        self._results[slot] = result
      with torch.cuda.stream(self._s_out):
        self._s_out.wait_event(self._ready[slot])
        host_out[i].copy_(result, non_blocking=True)
        self._out_done[slot].record(self._s_out)
    self._next = (self._next + len(host_frames)) % self.slots
    self._done.record(self._s_out)
    caller.wait_event(self._done)

  def wait(self) -> None:
    self._done.synchronize()
