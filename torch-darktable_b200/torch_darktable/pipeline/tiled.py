"""One oversize frame split into row tiles across GPUs (B200 addition, no counterpart in the reference).

Every rank owns a contiguous band of rows of the 12-bit packed frame.  The pipeline's dependency radius in rows
(RCD 10 + smoothing passes + Wiener tile 32 + bilateral ~4 sigma_s) is covered by a halo of packed rows that neighbouring
ranks exchange once, point to point (NCCL send/recv over NVLink between GPUs; gloo in the CPU tests), after which every
rank runs the ordinary stage kernels on its padded band.  Only three tiny reductions cross ranks, exactly where the
single-GPU pipeline has its global barriers:

  green sums (2 floats, SUM)   -> global green-equilibration ratio          (postprocess.cu:355-366)
  bounds     (2 floats, MIN/MAX) -> normalisation, with the EMA of ImageProcessor (image_processor.py:288-290)
  metrics    (6 floats, SUM)   -> tone-mapping metrics, with their EMA      (image_processor.py:292-294)
(the fused path carries the first two in ONE all-gather of six floats per rank: two latency-bound collectives per frame)

Band boundaries are multiples of 8 rows, which keeps the Bayer phase, the Wiener tile phase (stride 8), the sampling phase of
the bounds / metrics (stride 8) and -- for sigma_s in {1, 2, 4, 8} -- the bilateral grid phase identical to the untiled frame, so
the owned rows equal the single-GPU result up to the summation order of the three reductions.  Known deviations, both
inherited from reference artefacts that depend on absolute positions: the RCD stale-cell band (3 px just inside the 7-px
margin at the LEFT/RIGHT borders) and bilateral grids that saturate in y (height / sigma_s > 3000), which the
split refuses.  The Laplacian local contrast has a support of thousands of rows and is not available here.

By default a band runs through the FUSED frame kernels of libtdb200 (include/tdb200.h "Fused frame pipeline": the statistics
kernels have band forms that count the owned rows only and leave raw sums for the all-reduce); band boundaries and the halo are
then multiples of 32 rows, the tile height of the smoothing kernel.  With an `ops` object the stages are called one by one:
`CudaOps` is the stage-by-stage product path, the CPU tests inject an oracle-backed object so that the partitioning, the halo
exchange and the reductions are exercised under gloo without a GPU.
"""

from __future__ import annotations

from dataclasses import dataclass

import torch

from .config import ImageProcessingSettings, ToneMapper
from .util import lerp

ROW_ALIGN = 8
FUSED_ALIGN = 32  # the statistics of the fused smoothing kernel come per 32-row tile


def partition_rows(height: int, world: int, align: int = ROW_ALIGN) -> list[tuple[int, int]]:
  """Contiguous row bands [y0, y1) for `world` ranks; every boundary is a multiple of `align`."""
  if height % 2:
    raise ValueError('frame height must be even')
  units = height // align
  if units < world:
    raise ValueError(f'{height} rows cannot be split into {world} bands of at least {align} rows')
  bands = []
  for r in range(world):
    y0 = (units * r // world) * align
    y1 = (units * (r + 1) // world) * align if r + 1 < world else height
    bands.append((y0, y1))
  return bands


def halo_rows(settings: ImageProcessingSettings) -> int:
  """Rows of context a band needs on each side: RCD 10 (PPG 6, bilinear 2), one per smoothing pass, Wiener tile 32,
  bilateral blur/trilinear 4 cells; rounded up to the band alignment."""
  radius = 10 + 2
  if settings.postprocess:
    radius += settings.color_smoothing_passes + 2
  if settings.enable_denoise:
    radius += 32
  if settings.enable_bilateral:
    radius += int(4 * settings.bil_sigma_spatial + 0.999)
  return (radius + ROW_ALIGN - 1) // ROW_ALIGN * ROW_ALIGN


@dataclass
class Band:
  rank: int
  world: int
  y0: int          # owned rows [y0, y1) of the full frame
  y1: int
  top: int         # halo rows actually present above / below (0 at the frame border)
  bottom: int

  @property
  def padded(self) -> tuple[int, int]:
    return self.y0 - self.top, self.y1 + self.bottom


def make_band(height: int, rank: int, world: int, halo: int, align: int = ROW_ALIGN) -> Band:
  y0, y1 = partition_rows(height, world, align)[rank]
  if world > 1 and y1 - y0 < halo:
    raise ValueError(f'band of {y1 - y0} rows is thinner than the {halo}-row halo: use fewer ranks')
  return Band(rank, world, y0, y1, top=halo if rank > 0 else 0, bottom=halo if rank + 1 < world else 0)


# ---- collectives ---------------------------------------------------------------------------------------------------
class DistCollective:
  """torch.distributed plumbing: NCCL between GPUs (halo rows travel GPU to GPU over NVLink), gloo on CPU tensors."""

  def __init__(self, group=None):
    import torch.distributed as dist
    self.dist = dist
    self.group = group
    self.rank = dist.get_rank(group)
    self.world = dist.get_world_size(group)

  def exchange_halos(self, padded: torch.Tensor, row_bytes: int, band: Band) -> torch.Tensor:
    """padded: the rank's persistent band buffer (halo above | own rows | halo below, uint8) whose middle part already holds the own
    rows.  The first / last `halo` own rows go to the neighbours and theirs arrive straight in the halo parts: one batched
    isend / irecv, no staging copies (slices of a 1-D tensor are contiguous), no concatenation."""
    dist = self.dist
    t, b, n = band.top * row_bytes, band.bottom * row_bytes, padded.numel()
    ops = []
    if band.top:      # my first rows go up, the upper neighbour's last rows come down
      ops.append(dist.P2POp(dist.isend, padded[t: 2 * t], self.rank - 1, self.group))
      ops.append(dist.P2POp(dist.irecv, padded[:t], self.rank - 1, self.group))
    if band.bottom:
      ops.append(dist.P2POp(dist.isend, padded[n - 2 * b: n - b], self.rank + 1, self.group))
      ops.append(dist.P2POp(dist.irecv, padded[n - b:], self.rank + 1, self.group))
    if ops:
      for req in dist.batch_isend_irecv(ops):
        req.wait()
    return padded

  def all_gather(self, t: torch.Tensor) -> torch.Tensor:
    """(world, n) stack of every rank's vector: ONE collective carries the green sums, the minima and the maxima of a band; each
    rank then reduces the rows itself, in rank order (bit-identical on every rank)."""
    out = torch.empty((self.world, t.numel()), dtype=t.dtype, device=t.device)
    self.dist.all_gather_into_tensor(out, t.contiguous().reshape(1, -1), group=self.group)
    return out

  def all_reduce(self, t: torch.Tensor, op: str) -> torch.Tensor:
    dist = self.dist
    red = {'sum': dist.ReduceOp.SUM, 'min': dist.ReduceOp.MIN, 'max': dist.ReduceOp.MAX}[op]
    out = t.clone()
    dist.all_reduce(out, op=red, group=self.group)
    return out


class ThreadCollective:
  """The same collectives between `world` Python threads of ONE process (one thread per band, all on one device): used to check
  a split against the untiled result on a single GPU, and by the CPU tests.  Create one shared `ThreadCollective.Hub(world)`."""

  class Hub:
    def __init__(self, world: int):
      import threading
      self.world = world
      self.barrier = threading.Barrier(world)
      self.slots: list = [None] * world

  def __init__(self, hub: 'ThreadCollective.Hub', rank: int):
    self.hub, self.rank, self.world = hub, rank, hub.world

  def _gather(self, value):
    hub = self.hub
    hub.slots[self.rank] = value
    hub.barrier.wait()
    values = list(hub.slots)
    hub.barrier.wait()  # nobody overwrites a slot before everybody has read it
    return values

  def exchange_halos(self, padded: torch.Tensor, row_bytes: int, band: Band) -> torch.Tensor:
    bufs = self._gather((padded, band))
    t, b = band.top * row_bytes, band.bottom * row_bytes
    if band.top:     # the upper neighbour's last own rows
      up, ub = bufs[self.rank - 1]
      end = up.numel() - ub.bottom * row_bytes
      padded[:t].copy_(up[end - t: end])
    if band.bottom:  # the lower neighbour's first own rows
      dn, db = bufs[self.rank + 1]
      start = db.top * row_bytes
      padded[padded.numel() - b:].copy_(dn[start: start + b])
    self.hub.barrier.wait()  # nobody refills its buffer before the neighbours have read it
    return padded

  def all_gather(self, t: torch.Tensor) -> torch.Tensor:
    return torch.stack(self._gather(t))

  def all_reduce(self, t: torch.Tensor, op: str) -> torch.Tensor:
    stacked = torch.stack(self._gather(t))
    return {'sum': stacked.sum(0), 'min': stacked.min(0).values, 'max': stacked.max(0).values}[op]


# ---- the product stage calls ---------------------------------------------------------------------------------------
class CudaOps:
  """Stage kernels of libtdb200 through the reference-facing Python API."""

  def __init__(self):
    import torch_darktable as td
    from ..extension import extension
    self.td, self.ext = td, extension

  def demosaic(self, packed, size, pattern, fmt, settings, white_balance):
    return self.td.demosaic_packed(packed, size, pattern, method=settings.debayer.name, format_type=fmt, white_balance=white_balance,
                                   ppg_median_threshold=settings.ppg_median_threshold)

  def smooth(self, rgb, size, pattern, passes):
    pp = self.td.PostProcess(rgb.device, size, pattern, color_smoothing_passes=passes, green_eq_local=False, green_eq_global=False)
    return pp.process(rgb)

  def green_sums(self, rgb, pattern): return self.ext.green_sums(rgb, pattern.value)
  def green_eq_apply(self, rgb, ratio, pattern): return self.ext.green_eq_apply(rgb, ratio, pattern.value)
  def bounds(self, rgb, stride): return self.td.compute_image_bounds([rgb], stride=stride)
  def normalize(self, rgb, bounds): return self.ext.normalize(rgb, bounds)

  def denoise(self, rgb, size, noise):
    return self.td.Wiener(rgb.device, size).process_log_luminance(rgb, noise)

  def bilateral(self, rgb, size, sigma_s, sigma_r, detail):
    return self.td.Bilateral(rgb.device, size, sigma_s=sigma_s, sigma_r=sigma_r).process_rgb(rgb, detail)

  def metric_sums(self, rgb, stride): return self.ext.image_metric_sums([rgb], stride=stride)
  def metrics_from_sums(self, sums): return self.ext.metrics_from_sums(sums)

  def tonemap(self, rgb, op, metrics, params):
    return self.ext.tonemap(rgb, op, None if op == 'aces' else metrics, params.to_cpp(), None, 'none')


_TONEMAP_OPS = {ToneMapper.reinhard: 'reinhard', ToneMapper.linear: 'linear', ToneMapper.aces: 'aces',
                ToneMapper.adaptive_aces: 'adaptive_aces'}


class TiledFrameProcessor:
  """ImageProcessor for ONE rank's band of an oversize frame.  `image_size` is the FULL frame (width, height)."""

  def __init__(self, image_size: tuple[int, int], bayer_pattern, packed_format, settings: ImageProcessingSettings, device: torch.device,
               white_balance: tuple[float, float, float] | None, collective, ops=None):
    import torch_darktable as td
    self.td = td
    self.width, self.height = image_size
    self.bayer_pattern, self.packed_format, self.settings, self.device = bayer_pattern, packed_format, settings, device
    self.collective = collective
    # fused kernels unless stage objects are given (or the settings leave the fused path: no post-process smoothing)
    self.fused = ops is None and settings.postprocess and settings.color_smoothing_passes >= 1
    self.ops = ops if ops is not None else CudaOps()
    align = FUSED_ALIGN if self.fused else ROW_ALIGN
    self.halo = (halo_rows(settings) + align - 1) // align * align
    self.band = make_band(self.height, collective.rank, collective.world, self.halo, align)
    self.white_balance = torch.tensor(white_balance, dtype=torch.float32, device=device) if white_balance is not None else None
    self.bounds: torch.Tensor | None = None   # EMA state, identical on every rank
    self.metrics: torch.Tensor | None = None
    self.profile = False                      # record a CUDA event after every phase of `process` (see `breakdown`)
    self._events: list | None = None
    if self.width % 2:
      raise ValueError('frame width must be even')
    if settings.enable_bilateral:
      s = float(settings.bil_sigma_spatial)
      if round(self.height / s) > 3000:
        raise ValueError('the bilateral grid of the full frame saturates in y (height / sigma_s > 3000): not reproducible per band')
      if abs(self.band.padded[0] / s - round(self.band.padded[0] / s)) > 1e-6:
        raise ValueError(f'bil_sigma_spatial={s} does not divide the band origin {self.band.padded[0]}: the grid phase would shift')

  def _mark(self, name: str):
    if self._events is not None:
      ev = torch.cuda.Event(enable_timing=True)
      ev.record()
      self._events.append((name, ev))

  def breakdown(self) -> dict[str, float]:
    """Milliseconds per phase of the last `process` call made with `profile = True` (synchronises the device)."""
    if not self._events:
      return {}
    torch.cuda.synchronize(self.device)
    out: dict[str, float] = {}
    for (_, a), (name, b) in zip(self._events, self._events[1:]):
      out[name] = out.get(name, 0.0) + a.elapsed_time(b)
    return out

  @property
  def row_bytes(self) -> int:
    return self.width * 3 // 2

  @property
  def owned_rows(self) -> tuple[int, int]:
    return self.band.y0, self.band.y1

  def own_rows_buffer(self) -> torch.Tensor:
    """The part of the persistent padded band buffer that holds this rank's own packed rows.  A producer that writes the rows
    straight into it (a raw-file reader, an H2D copy) and then passes this very tensor to `process` saves the band-sized copy."""
    b = self.band
    if getattr(self, '_padded', None) is None:
      p0, p1 = b.padded
      self._padded = torch.empty((p1 - p0) * self.row_bytes, dtype=torch.uint8, device=self.device)
    t = b.top * self.row_bytes
    return self._padded[t: t + (b.y1 - b.y0) * self.row_bytes]

  def _fill_padded(self, own_packed_rows: torch.Tensor) -> torch.Tensor:
    view = self.own_rows_buffer()
    if own_packed_rows.data_ptr() != view.data_ptr():
      view.copy_(own_packed_rows)
    return self._padded

  def _own(self, image: torch.Tensor) -> torch.Tensor:
    """Rows this rank owns, out of a tensor laid out over the padded band."""
    return image[self.band.top: self.band.top + (self.band.y1 - self.band.y0)]

  def _fused_objects(self, size):
    if getattr(self, '_fused_size', None) != size:
      from ..extension import extension
      s = self.settings
      self._frame = extension.FramePipeline(self.device, size[0], size[1], self.bayer_pattern.value)
      self._post = self.td.PostProcess(self.device, size, self.bayer_pattern, color_smoothing_passes=s.color_smoothing_passes,
                                       green_eq_local=False, green_eq_global=True)._postprocess
      self._wiener = self.td.Wiener(self.device, size)._wiener if s.enable_denoise else None
      self._bil = (self.td.Bilateral(self.device, size, sigma_s=s.bil_sigma_spatial, sigma_r=s.bil_sigma_luminance)._bilateral
                   if s.enable_bilateral else None)
      self._fused_size = size
    return self._frame, self._post, self._wiener, self._bil

  def _process_fused(self, packed: torch.Tensor, size) -> torch.Tensor:
    """The band through the fused frame kernels; TWO collectives of six floats per frame -- an all-gather of the smoothing statistics
    (green sums, minima, maxima) and an all-reduce of the metric sums -- where the single-GPU pipeline has its barriers."""
    from ..extension import extension
    b, s, col = self.band, self.settings, self.collective
    frame, post, wiener, bil = self._fused_objects(size)
    lo, hi = b.top, b.top + (b.y1 - b.y0)  # owned rows inside the padded band
    rgb = self.td.demosaic_packed(packed, size, self.bayer_pattern, method=s.debayer.name, format_type=self.packed_format,
                                  white_balance=self.white_balance, ppg_median_threshold=s.ppg_median_threshold)
    smoothed, raw = frame.smooth_band(post, rgb, lo, hi)
    self._mark('demosaic + smoothing')
    stats = col.all_gather(raw)  # (world, 6): one collective for the green sums, the minima and the maxima
    # ratio, bounds of the equilibrated image (the G1 greens take the ratio: x -> max(0, x * ratio) is monotone) and their EMA: one kernel,
    # the state is updated in place
    prev = self.bounds
    if self.bounds is None:
      self.bounds = torch.empty(2, dtype=torch.float32, device=self.device)
    ratio = extension.band_stats_finish(stats, prev, s.moving_average, self.bounds)
    self._mark('all-gather of the band statistics + bounds')

    image = frame.prepare(smoothed, ratio, self.bounds, wiener)
    if wiener is not None:
      image = frame.denoise(wiener, image, s.denoise, True, bil)
    elif bil is not None:
      frame.bilateral_grid(bil, image)
    lab = wiener is not None and bil is not None
    local_sums = frame.metric_sums_band(image, bil, s.bilateral, lo, hi, lab_input=lab)
    self._mark('prepare + Wiener + bilateral grid + metric sums')
    msums = col.all_reduce(local_sums, 'sum')
    prev = self.metrics
    if self.metrics is None:
      self.metrics = torch.empty(5, dtype=torch.float32, device=self.device)
    extension.band_metrics_finish(msums, prev, s.moving_average, self.metrics)
    self._mark('all-reduce of the metric sums + metrics')

    params = self.td.TonemapParameters(s.tone_gamma, s.tone_intensity, s.light_adapt, s.vibrance).to_cpp()
    op = _TONEMAP_OPS[s.tone_mapping]
    if bil is not None:
      out = frame.slice_tonemap(image, bil, s.bilateral, op, self.metrics, params, None, 'none', lab_input=lab)
    else:
      out = extension.tonemap(image, op, None if op == 'aces' else self.metrics, params, None, 'none')
    own = self._own(out).contiguous()
    self._mark('slice + tone map + crop')
    return own

  def process(self, own_packed_rows: torch.Tensor) -> torch.Tensor:
    """own_packed_rows: uint8 tensor with the packed bytes of this rank's rows.  Returns the uint8 (rows, W, 3) sRGB band."""
    b, s, ops, col = self.band, self.settings, self.ops, self.collective
    if own_packed_rows.numel() != (b.y1 - b.y0) * self.row_bytes:
      raise ValueError(f'expected {(b.y1 - b.y0) * self.row_bytes} packed bytes for rows {b.y0}..{b.y1}, got {own_packed_rows.numel()}')
    self._events = [] if self.profile else None
    self._mark('start')
    packed = col.exchange_halos(self._fill_padded(own_packed_rows), self.row_bytes, b)
    self._mark('halo exchange')
    p0, p1 = b.padded
    size = (self.width, p1 - p0)
    if self.fused:
      return self._process_fused(packed, size)

    rgb = ops.demosaic(packed, size, self.bayer_pattern, self.packed_format, s, self.white_balance)
    if s.postprocess:
      rgb = ops.smooth(rgb, size, self.bayer_pattern, s.color_smoothing_passes)
      sums = col.all_reduce(ops.green_sums(self._own(rgb), self.bayer_pattern), 'sum')
      ratio = torch.where((sums[0] > 0) & (sums[1] > 0), sums[1] / sums[0], torch.ones_like(sums[0])).reshape(1)
      rgb = ops.green_eq_apply(rgb, ratio, self.bayer_pattern)

    local = ops.bounds(self._own(rgb), 8)
    bounds = torch.stack([col.all_reduce(local[0:1], 'min')[0], col.all_reduce(local[1:2], 'max')[0]])
    self.bounds = lerp(self.bounds if self.bounds is not None else bounds, bounds, s.moving_average)

    rgb = ops.normalize(rgb, self.bounds)
    if s.enable_denoise:
      rgb = ops.denoise(rgb, size, s.denoise)
    if s.enable_bilateral:
      rgb = ops.bilateral(rgb, size, s.bil_sigma_spatial, s.bil_sigma_luminance, s.bilateral)

    msums = col.all_reduce(ops.metric_sums(self._own(rgb), 8), 'sum')
    metrics = ops.metrics_from_sums(msums)
    self.metrics = lerp(self.metrics if self.metrics is not None else metrics, metrics, s.moving_average)

    params = self.td.TonemapParameters(s.tone_gamma, s.tone_intensity, s.light_adapt, s.vibrance)
    return ops.tonemap(self._own(rgb).contiguous(), _TONEMAP_OPS[s.tone_mapping], self.metrics, params)
