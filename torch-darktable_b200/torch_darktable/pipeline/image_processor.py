"""ImageProcessor: packed 12-bit Bayer bytes -> display-ready uint8 RGB.

Same constructor, methods, EMA state and error types as the reference's pipeline/image_processor.py:31-319, but the
stages run as fused libtdb200 kernels on the current CUDA stream without any host synchronisation:

  load_image        packed bytes -> RGB in ONE kernel (unpack + white balance + demosaic), then the post-process
  process_rgb       normalise, Wiener on log-luminance (fused), bilateral local contrast on RGB (fused)
  process_image_set joint bounds / metrics with device-resident EMA, tone map with the camera transform fused in
"""

from __future__ import annotations

import os

from beartype import beartype
import torch

import torch_darktable as td
from torch_darktable import _lib
from torch_darktable.extension import extension

from .camera_settings import CameraSettings
from .config import Debayer, ImageProcessingSettings, ToneMapper
from .transform import ImageTransform, transform
from .util import lerp, normalize_image, resize_longest_edge


class ImageSizeMismatchError(Exception):
  """The byte count of a raw frame does not match the configured sensor."""

  def __init__(self, message: str, image_size: tuple[int, int], packed_format: td.PackedFormat, padding: int):
    super().__init__(message)
    self.image_size = image_size
    self.packed_format = packed_format
    self.padding = padding


_TONEMAP_OPS = {ToneMapper.reinhard: 'reinhard', ToneMapper.linear: 'linear', ToneMapper.aces: 'aces',
                ToneMapper.adaptive_aces: 'adaptive_aces'}


@beartype
class ImageProcessor:
  LANES = 3  # most frames in flight through `submit` / `process_batch` / HostFrameRunner (three below 3 MP, two above: `_n_lanes`)

  @beartype
  def __init__(self, image_size: tuple[int, int], bayer_pattern: td.BayerPattern, packed_format: td.PackedFormat,
               settings: ImageProcessingSettings, device: torch.device, white_balance: tuple[float, float, float] | None,
               transforms: ImageTransform | dict[str, ImageTransform] = ImageTransform.none, padding: int = 0):
    assert device.index is not None, f'Device not fully specified: {device}'
    self.device = device
    self.settings = settings
    self.image_size = image_size  # (width, height)
    self.bayer_pattern = bayer_pattern
    self.packed_format = packed_format
    self.transforms = transforms
    self.padding = padding

    self.metrics: torch.Tensor | None = None  # EMA state across image sets
    self.bounds: torch.Tensor | None = None
    # the fused path keeps both in these two device buffers and updates them IN PLACE (the kernels read the previous value and write
    # the blended one, pipeline/util.py:4 lerp): a captured CUDA graph of an image set then advances the state on every replay
    self._bounds_state = torch.empty(2, dtype=torch.float32, device=device)
    self._metrics_state = torch.empty(5, dtype=torch.float32, device=device)
    self._batch_graph = None  # (key, graph, static input, static output) of process_batch
    # Several frames in flight (`submit`): consecutive frames rotate over LANES lanes -- a CUDA stream plus its own workspaces -- so
    # that the tail of one frame's kernels and its single-CTA statistics steps overlap the next frames' work (measured at 4K: +9 %
    # with two lanes, +10.5 % with three).  The EMA chain stays sequential: frame i reads the state frame i - 1 wrote, through a ring
    # of LANES device buffers ([0] = the resting buffers above; frame i writes slot i % LANES on lane i % LANES, so the slot is next
    # written by the frame that follows frame i on the SAME stream) and one event per statistic.
    self._bounds_pp = [self._bounds_state] + [torch.empty(2, dtype=torch.float32, device=device) for _ in range(self.LANES - 1)]
    self._metrics_pp = [self._metrics_state] + [torch.empty(5, dtype=torch.float32, device=device) for _ in range(self.LANES - 1)]
    # measured on a B200 (profiles/r02_batch_graph.jsonl): 1080p 0.224 ms per frame with two lanes, 0.219 with three; 4K 0.805 / 0.815 ms
    # (three 4K frames' working sets evict each other from L2)
    self._n_lanes = 3 if image_size[0] * image_size[1] < 3_000_000 else 2
    if os.environ.get('TDB_LANES'):  # A/B runs
      self._n_lanes = max(1, min(self.LANES, int(os.environ['TDB_LANES'])))
    self._lanes: list | None = None
    self._ev_bounds: torch.cuda.Event | None = None
    self._ev_metrics: torch.cuda.Event | None = None

    self.white_balance = (torch.tensor(white_balance, device=device).to(torch.float32) if white_balance is not None else None)
    self._frame = extension.FramePipeline(device, image_size[0], image_size[1], bayer_pattern.value)
    self._bil_slots: list = []
    self.wiener_workspace = td.Wiener(device, image_size)
    self.rcd_workspace = td.RCD(device, image_size, bayer_pattern)
    self._make_bilateral(settings)
    self._make_ppg(settings)
    self._make_postprocess(settings)

  # -- workspaces -----------------------------------------------------------------------------------------------
  def _make_bilateral(self, s: ImageProcessingSettings):
    self._bil_slots = []
    self.bil_workspace = td.Bilateral(self.device, self.image_size, sigma_s=s.bil_sigma_spatial, sigma_r=s.bil_sigma_luminance)

  def _make_ppg(self, s: ImageProcessingSettings):
    self.ppg_workspace = td.PPG(self.device, self.image_size, self.bayer_pattern, median_threshold=s.ppg_median_threshold)

  def _make_postprocess(self, s: ImageProcessingSettings):
    self.postprocess_workspace = td.PostProcess(self.device, self.image_size, self.bayer_pattern,
                                                color_smoothing_passes=s.color_smoothing_passes, green_eq_local=False,
                                                green_eq_global=True, green_eq_threshold=s.green_eq_threshold)

  def update_settings(self, settings: ImageProcessingSettings):
    old, self.settings = self.settings, settings
    self._join_lanes()
    self._batch_graph, self._lanes = None, None

    def changed(*names: str) -> bool:
      return any(getattr(old, n) != getattr(settings, n) for n in names)

    if changed('bil_sigma_spatial', 'enable_bilateral', 'bil_sigma_luminance'):
      self._make_bilateral(settings)
    if changed('ppg_median_threshold'):
      self._make_ppg(settings)
    if changed('color_smoothing_passes', 'green_eq_threshold'):
      self._make_postprocess(settings)

  @staticmethod
  def from_camera_settings(camera_settings: CameraSettings, device: torch.device):
    return ImageProcessor(camera_settings.image_size, camera_settings.bayer_pattern, camera_settings.packed_format,
                          camera_settings.image_processing, device=device, white_balance=camera_settings.white_balance,
                          transforms=camera_settings.transform, padding=camera_settings.padding)

  def __repr__(self) -> str:
    wb = 'None' if self.white_balance is None else '({:.3f}, {:.3f}, {:.3f})'.format(*self.white_balance.tolist())
    if isinstance(self.transforms, ImageTransform):
      tf = self.transforms.name
    else:
      tf = '{' + ', '.join(f'{k}: {v.name}' for k, v in self.transforms.items()) + '}'
    return (f'ImageProcessor(size={self.image_size}, bayer={self.bayer_pattern.name}, format={self.packed_format.name}, '
            f'device={self.device}, wb={wb}, padding={self.padding}, transform={tf}, debayer={self.settings.debayer.name}, '
            f'tonemap={self.settings.tone_mapping.name})')

  # -- sizes ----------------------------------------------------------------------------------------------------
  @property
  def final_size(self):
    return resize_longest_edge(self.image_size, self.settings.resize_width)

  @property
  def expected_bytes(self) -> int:
    width, height = self.image_size
    if self.packed_format not in (td.PackedFormat.Packed12, td.PackedFormat.Packed12_IDS):
      raise ValueError(f'Unsupported packed format: {self.packed_format}')
    return (width * height * 3) // 2 + self.padding

  def _mismatch(self, message: str) -> ImageSizeMismatchError:
    return ImageSizeMismatchError(message, image_size=self.image_size, packed_format=self.packed_format, padding=self.padding)

  def _strip(self, raw: torch.Tensor) -> torch.Tensor:
    if raw.numel() != self.expected_bytes:
      raise self._mismatch(f'Image size mismatch: expected {self.expected_bytes} bytes for {self.image_size} '
                           f'{self.packed_format.name} with {self.padding} padding, got {raw.numel()} bytes. ')
    return raw[: -self.padding] if self.padding > 0 else raw

  # -- stages ---------------------------------------------------------------------------------------------------
  @beartype
  def load_bytes(self, bytes: torch.Tensor) -> torch.Tensor:
    """Packed bytes -> (H, W) float32 CFA in [0, 1] (no white balance)."""
    decoded = td.decode12(self._strip(bytes), output_dtype=torch.float32, format_type=self.packed_format)
    width, height = self.image_size
    if decoded.numel() != width * height:
      raise self._mismatch(f'Decoded image size mismatch: expected {width * height} pixels ({width}x{height}), '
                           f'got {decoded.numel()} pixels.')
    return decoded.view(height, width)

  @beartype
  def load_image(self, bytes: torch.Tensor) -> torch.Tensor:
    """Packed bytes -> demosaiced (and post-processed) linear RGB; one fused kernel up to the demosaic."""
    s = self.settings
    rgb = td.demosaic_packed(self._strip(bytes), self.image_size, self.bayer_pattern, method=s.debayer.name,
                             format_type=self.packed_format, white_balance=self.white_balance,
                             ppg_median_threshold=s.ppg_median_threshold)
    return self.postprocess_workspace.process(rgb) if s.postprocess else rgb

  def debayer(self, bayer_image: torch.Tensor) -> torch.Tensor:
    """(H, W) float CFA -> RGB, stage by stage (kept for API parity; load_image uses the fused path)."""
    assert bayer_image.ndim == 2, f'Bayer image must have 2 dimensions, got {bayer_image.shape}'
    if self.white_balance is not None:
      bayer_image = td.apply_white_balance(bayer_image, self.white_balance, self.bayer_pattern)
    cfa = bayer_image.unsqueeze(-1)
    if self.settings.debayer == Debayer.bilinear:
      rgb = td.bilinear5x5_demosaic(cfa, self.bayer_pattern)
    elif self.settings.debayer == Debayer.rcd:
      rgb = self.rcd_workspace.process(cfa)
    elif self.settings.debayer == Debayer.ppg:
      rgb = self.ppg_workspace.process(cfa)
    else:
      raise AssertionError(f'Invalid debayer method: {self.settings.debayer}')
    return self.postprocess_workspace.process(rgb) if self.settings.postprocess else rgb

  @beartype
  def process_rgb(self, rgb_raw: torch.Tensor, bounds: torch.Tensor | None = None) -> torch.Tensor:
    if bounds is not None:
      rgb_raw = normalize_image(rgb_raw, bounds)
    if self.settings.enable_denoise:
      rgb_raw = self.wiener_workspace.process_log_luminance(rgb_raw, self.settings.denoise)
    if self.settings.enable_bilateral:
      rgb_raw = self.bil_workspace.process_rgb(rgb_raw, self.settings.bilateral)
    return rgb_raw

  def _transform_for(self, image_name: str) -> ImageTransform:
    return self.transforms[image_name] if isinstance(self.transforms, dict) else self.transforms

  def transform(self, image: torch.Tensor, image_name: str) -> torch.Tensor:
    return transform(image, self._transform_for(image_name))

  def tonemap(self, rgb_raw: torch.Tensor, metrics: torch.Tensor | None = None,
              image_transform: ImageTransform = ImageTransform.none) -> torch.Tensor:
    s = self.settings
    params = td.TonemapParameters(s.tone_gamma, s.tone_intensity, s.light_adapt, s.vibrance)
    if metrics is None:
      metrics = td.compute_image_metrics([rgb_raw], stride=4, min_gray=1e-4)
    op = _TONEMAP_OPS[s.tone_mapping]
    return extension.tonemap(rgb_raw, op, None if op == 'aces' else metrics, params.to_cpp(), None, image_transform.name)

  @beartype
  def process(self, bytes: torch.Tensor, image_name: str) -> torch.Tensor:
    return self.process_image_set({image_name: bytes})[image_name]

  @beartype
  def process_image_set(self, image_set_bytes: dict[str, torch.Tensor]) -> dict[str, torch.Tensor]:
    if os.environ.get('TDB_UNFUSED'):
      return self.process_image_set_by_stage(image_set_bytes)
    return self._process_image_set_fused(image_set_bytes)

  @beartype
  def process_image_set_by_stage(self, image_set_bytes: dict[str, torch.Tensor]) -> dict[str, torch.Tensor]:
    """The image-set composite written with the public stage calls, line by line as in the reference
    (pipeline/image_processor.py:284-300).  `process_image_set` computes the same thing with the fused kernels."""
    names = list(image_set_bytes.keys())
    rgb_raw = [self.load_image(b) for b in image_set_bytes.values()]

    bounds = td.compute_image_bounds(rgb_raw, stride=8)
    self.bounds = lerp(self.bounds if self.bounds is not None else bounds, bounds, self.settings.moving_average)
    rgb = [self.process_rgb(image, self.bounds) for image in rgb_raw]

    metrics = td.compute_image_metrics(rgb, stride=8)
    self.metrics = lerp(self.metrics if self.metrics is not None else metrics, metrics, self.settings.moving_average)

    return {name: self.tonemap(image, self.metrics, self._transform_for(name)) for name, image in zip(names, rgb, strict=True)}

  def _state_in_place(self, t: torch.Tensor | None, state: torch.Tensor) -> torch.Tensor | None:
    """The EMA state as the kernels see it: None before the first image set, else `state` (holding t's values)."""
    if t is None:
      return None
    assert t.numel() == state.numel(), f'expected {state.numel()} values, got {t.numel()}'
    if t is not state:  # set by the stage-by-stage path or by the caller
      state.copy_(t.to(device=self.device, dtype=torch.float32).reshape(-1))
    return state

  def _bilateral_for(self, slot: int):
    """One grid per frame of an image set: the slice of frame i runs after the statistics of the whole set are known."""
    while len(self._bil_slots) <= slot:
      s = self.settings
      self._bil_slots.append(self.bil_workspace if not self._bil_slots else
                             td.Bilateral(self.device, self.image_size, sigma_s=s.bil_sigma_spatial, sigma_r=s.bil_sigma_luminance))
    return self._bil_slots[slot]._bilateral

  def _process_image_set_fused(self, image_set_bytes: dict[str, torch.Tensor], outs: dict[str, torch.Tensor] | None = None) -> dict[str, torch.Tensor]:
    """Fused kernels (include/tdb200.h, "Fused frame pipeline"): per frame
         demosaic from packed bytes -> smoothing (+ green ratio, bounds of the set, EMA)            [barrier: bounds]
         green-eq + normalise + log-luminance -> Wiener tiles -> normalise + splat -> grid blur
         metrics at the sampled pixels of the sliced image (+ EMA)                                  [barrier: metrics]
         slice + tone map + transform -> uint8
       with no host synchronisation and no statistic launch of its own."""
    s = self.settings
    names = list(image_set_bytes.keys())
    n = len(names)
    if n == 0:
      return self.process_image_set_by_stage(image_set_bytes)
    frame, ma = self._frame, float(s.moving_average)
    self._join_lanes()  # frames still in flight on the lanes (submit) come first
    prev_bounds = self._state_in_place(self.bounds, self._bounds_state)
    prev_metrics = self._state_in_place(self.metrics, self._metrics_state)

    # -- A: load; bounds of the set
    if s.postprocess and s.color_smoothing_passes >= 1:
      bounds = self._bounds_state
      raw, ratios = [], []
      for i, b in enumerate(image_set_bytes.values()):
        rgb = td.demosaic_packed(self._strip(b), self.image_size, self.bayer_pattern, method=s.debayer.name, format_type=self.packed_format,
                                 white_balance=self.white_balance, ppg_median_threshold=s.ppg_median_threshold)
        smoothed, ratio = frame.smooth_deferred(self.postprocess_workspace._postprocess, rgb, i == 0, i == n - 1, prev_bounds, ma, bounds)
        raw.append(smoothed), ratios.append(ratio)
      self.bounds = bounds
    else:
      raw, ratios = [self.load_image(b) for b in image_set_bytes.values()], [None] * n
      bounds = td.compute_image_bounds(raw, stride=8)
      self._bounds_state.copy_(lerp(prev_bounds if prev_bounds is not None else bounds, bounds, ma))
      self.bounds = self._bounds_state

    # -- B, C, D: per frame up to the blurred bilateral grid; metrics of the set
    wiener = self.wiener_workspace._wiener if s.enable_denoise else None
    metrics = self._metrics_state
    images = []
    for i, (image, ratio) in enumerate(zip(raw, ratios, strict=True)):
      bil = self._bilateral_for(i) if s.enable_bilateral else None
      rgb = frame.prepare(image, ratio, self.bounds, wiener)
      if wiener is not None:
        rgb = frame.denoise(wiener, rgb, s.denoise, True, bil)  # Lab if the bilateral stage follows
      elif bil is not None:
        frame.bilateral_grid(bil, rgb)
      lab = wiener is not None and bil is not None
      frame.metrics(rgb, bil, s.bilateral, i == 0, i == n - 1, prev_metrics, ma, metrics, lab_input=lab)
      images.append(rgb)
    self.metrics = metrics

    # -- E: tone map (the bilateral slice happens inside)
    params = td.TonemapParameters(s.tone_gamma, s.tone_intensity, s.light_adapt, s.vibrance).to_cpp()
    op = _TONEMAP_OPS[s.tone_mapping]
    out = {}
    for i, (name, rgb) in enumerate(zip(names, images, strict=True)):
      tf = self._transform_for(name).name
      dst = outs.get(name) if outs is not None else None
      if s.enable_bilateral:
        out[name] = frame.slice_tonemap(rgb, self._bilateral_for(i), s.bilateral, op, self.metrics, params, None, tf,
                                        lab_input=s.enable_denoise, out=dst)
      else:
        out[name] = extension.tonemap(rgb, op, None if op == 'aces' else self.metrics, params, None, tf, out=dst)
    return out

  # -- two frames in flight -------------------------------------------------------------------------------------------
  class _Lane:
    __slots__ = ('stream', 'frame', 'post', 'wiener', 'bil', 'done')

  def _ensure_lanes(self) -> list:
    if self._lanes is None:
      s, lanes = self.settings, []
      for k in range(self._n_lanes):
        lane = ImageProcessor._Lane()
        lane.stream, lane.done = torch.cuda.Stream(self.device), None
        if k == 0:  # the processor's own workspaces
          lane.frame, lane.post = self._frame, self.postprocess_workspace._postprocess
          lane.wiener, lane.bil = self.wiener_workspace._wiener, self._bilateral_for(0)
        else:
          lane.frame = extension.FramePipeline(self.device, self.image_size[0], self.image_size[1], self.bayer_pattern.value)
          lane.post = td.PostProcess(self.device, self.image_size, self.bayer_pattern, color_smoothing_passes=s.color_smoothing_passes,
                                     green_eq_local=False, green_eq_global=True, green_eq_threshold=s.green_eq_threshold)._postprocess
          lane.wiener = td.Wiener(self.device, self.image_size)._wiener
          lane.bil = td.Bilateral(self.device, self.image_size, sigma_s=s.bil_sigma_spatial, sigma_r=s.bil_sigma_luminance)._bilateral
        lanes.append(lane)
      self._lanes = lanes
    return self._lanes

  def _join_lanes(self):
    """Order the current stream after every frame that is still in flight on a lane."""
    if self._lanes is not None:
      cur = torch.cuda.current_stream(self.device)
      for lane in self._lanes:
        if lane.done is not None:
          cur.wait_event(lane.done)
          lane.done = None
    self._ev_bounds = self._ev_metrics = None

  def _state_parity(self) -> int:
    """Index of the ping-pong buffers that hold the latest EMA state (values set from outside are copied into the resting pair)."""
    if self.bounds is None or self.metrics is None:
      return 0
    for k in range(self.LANES):
      if self.bounds is self._bounds_pp[k] and self.metrics is self._metrics_pp[k]:
        return k
    self._bounds_pp[0].copy_(self.bounds.to(device=self.device, dtype=torch.float32).reshape(-1))
    self._metrics_pp[0].copy_(self.metrics.to(device=self.device, dtype=torch.float32).reshape(-1))
    self.bounds, self.metrics = self._bounds_pp[0], self._metrics_pp[0]
    return 0

  @beartype
  def submit(self, bytes: torch.Tensor, image_name: str, out: torch.Tensor | None = None, track: bool = True):
    """`process(bytes, image_name)` without waiting for the previous frame to leave the GPU (B200 addition): the frame -- its own image
    set -- is enqueued on one of LANES rotating lanes (stream + workspaces), ordered after the work already queued on the current
    stream, and (result, event) is returned; the result is complete when the event is (`join()` orders the current stream after
    all frames in flight).  The bounds / metrics EMA is chained exactly as by consecutive `process` calls: frame i reads the state
    frame i - 1 wrote (a ring of device buffers, one event per statistic), so the results are those of the sequential calls."""
    s = self.settings
    if not (s.postprocess and s.color_smoothing_passes >= 1):  # not the fused nine-launch configuration: run in line
      res = self._process_image_set_fused({image_name: bytes}, {image_name: out} if out is not None else None)[image_name]
      done = torch.cuda.Event()
      done.record()
      return res, done
    lanes = self._ensure_lanes()
    cur = torch.cuda.current_stream(self.device)
    par = self._state_parity()  # (a copy, if any, is queued on the current stream, before the lane picks it up)
    prev_b = None if self.bounds is None else self._bounds_pp[par]
    prev_m = None if self.metrics is None else self._metrics_pp[par]
    nxt = (par + 1) % self._n_lanes  # ring slot AND lane of this frame
    out_b, out_m = self._bounds_pp[nxt], self._metrics_pp[nxt]
    lane = lanes[nxt]
    ma = float(s.moving_average)
    lane.stream.wait_stream(cur)
    _lib.lib.tdb_set_concurrency_hint(self._n_lanes)  # kernels that would fill an SM's register file alone leave room for the other lane's
    try:
      with torch.cuda.stream(lane.stream):
        rgb = td.demosaic_packed(self._strip(bytes), self.image_size, self.bayer_pattern, method=s.debayer.name, format_type=self.packed_format,
                                 white_balance=self.white_balance, ppg_median_threshold=s.ppg_median_threshold)
        if self._ev_bounds is not None:
          lane.stream.wait_event(self._ev_bounds)  # the previous frame's bounds are final (and its read of the buffer written here is done)
        smoothed, ratio = lane.frame.smooth_deferred(lane.post, rgb, True, True, prev_b, ma, out_b)
        self._ev_bounds = torch.cuda.Event()
        self._ev_bounds.record(lane.stream)
        wiener = lane.wiener if s.enable_denoise else None
        bil = lane.bil if s.enable_bilateral else None
        image = lane.frame.prepare(smoothed, ratio, out_b, wiener)
        if wiener is not None:
          image = lane.frame.denoise(wiener, image, s.denoise, True, bil)
        elif bil is not None:
          lane.frame.bilateral_grid(bil, image)
        lab = wiener is not None and bil is not None
        if self._ev_metrics is not None:
          lane.stream.wait_event(self._ev_metrics)
        lane.frame.metrics(image, bil, s.bilateral, True, True, prev_m, ma, out_m, lab_input=lab)
        self._ev_metrics = torch.cuda.Event()
        self._ev_metrics.record(lane.stream)
        params = td.TonemapParameters(s.tone_gamma, s.tone_intensity, s.light_adapt, s.vibrance).to_cpp()
        op, tf = _TONEMAP_OPS[s.tone_mapping], self._transform_for(image_name).name
        if bil is not None:
          res = lane.frame.slice_tonemap(image, bil, s.bilateral, op, out_m, params, None, tf, lab_input=s.enable_denoise, out=out)
        else:
          res = extension.tonemap(image, op, None if op == 'aces' else out_m, params, None, tf, out=out)
        lane.done = torch.cuda.Event()
        lane.done.record(lane.stream)
    finally:
      _lib.lib.tdb_set_concurrency_hint(0)
    if track and not torch.cuda.is_current_stream_capturing():
      # allocator bookkeeping across streams (track=False: the caller keeps `bytes` and the result alive until its own streams are done): the input is read on the lane, a result allocated on the lane is consumed on the caller's stream
      bytes.record_stream(lane.stream)
      if out is None:
        res.record_stream(cur)
    self.bounds, self.metrics = out_b, out_m
    return res, lane.done

  def join(self):
    """Order the current stream after every submitted frame and bring the EMA state back to its resting buffers."""
    self._join_lanes()
    for k in range(1, self.LANES):
      if self.bounds is self._bounds_pp[k]:
        self._bounds_pp[0].copy_(self._bounds_pp[k]), self._metrics_pp[0].copy_(self._metrics_pp[k])
        self.bounds, self.metrics = self._bounds_pp[0], self._metrics_pp[0]

  # -- batches ----------------------------------------------------------------------------------------------------
  @beartype
  def process_batch(self, frames: torch.Tensor, image_name: str = 'cam', graph: bool = True) -> torch.Tensor:
    """frames: (N, expected_bytes) uint8 packed frames on the device -> (N, H', W', 3) uint8, every frame its own image set -- N calls
    of `process(frame, image_name)` (reference pipeline/image_processor.py:274-300), EMA state carried from frame to frame.

    graph=True (B200 addition): the first call runs the batch eagerly and, while doing so, captures its launches
    (nine per frame, consecutive frames on two lanes: see `submit`) into ONE CUDA graph; later calls with the same batch size replay it
    -- one cudaGraphLaunch instead of 9 N kernel launches and ~40 N tensor allocations, which is what bounds small frames (at 256 x 192 a frame is 30 us of GPU work behind 150 us
    of Python).  The graph reads a static input buffer and writes a static output buffer: `frames` is copied in unless it IS that
    buffer (`batch_input_buffer(N)`), and the returned tensor is the static output, valid until the next call.  The EMA tensors are
    updated in place by the kernels, so a replay continues the state exactly as an eager call would."""
    if frames.dim() != 2 or frames.dtype != torch.uint8 or not frames.is_cuda:
      raise RuntimeError('frames must be an (N, bytes) uint8 CUDA tensor')
    n = frames.size(0)
    if frames.size(1) != self.expected_bytes:
      raise self._mismatch(f'Image size mismatch: expected {self.expected_bytes} bytes per frame, got {frames.size(1)} bytes. ')
    tf = self._transform_for(image_name)
    w, h = self.image_size
    shape = (n, w, h, 3) if tf in (ImageTransform.rotate_90, ImageTransform.rotate_270, ImageTransform.transpose) else (n, h, w, 3)
    def run(src, dst):  # two frames in flight; the state ends in its resting buffers, so that a captured run can be replayed
      self._join_lanes()
      for i in range(n):
        self.submit(src[i], image_name, dst[i], track=False)
      self.join()

    if not graph:
      out = torch.empty(shape, dtype=torch.uint8, device=self.device)
      run(frames, out)
      return out
    key = (n, image_name, self.settings, tf)
    if self._batch_graph is None or self._batch_graph[0] != key:
      static_in = self.batch_input_buffer(n)
      static_in.copy_(frames)
      static_out = torch.empty(shape, dtype=torch.uint8, device=self.device)
      run(static_in, static_out)  # the real thing for this call; it also initialises the EMA state and every lazily set kernel attribute
      result = static_out.clone()
      g = torch.cuda.CUDAGraph()
      with torch.cuda.graph(g):  # recorded, not executed: both lanes fork from and join the capturing stream
        run(static_in, static_out)
      self._batch_graph = (key, g, static_in, static_out)
      return result
    _, g, static_in, static_out = self._batch_graph
    if frames.data_ptr() != static_in.data_ptr():
      static_in.copy_(frames)
    self.join()  # the graph starts from the resting state buffers
    g.replay()
    return static_out

  def batch_input_buffer(self, n: int) -> torch.Tensor:
    """The (n, expected_bytes) device buffer the captured graph of `process_batch` reads: fill it (e.g. by H2D copies) and pass it
    to `process_batch` to save the device-to-device copy of the batch."""
    if self._batch_graph is not None and self._batch_graph[2].size(0) == n:
      return self._batch_graph[2]
    return torch.empty((n, self.expected_bytes), dtype=torch.uint8, device=self.device)
