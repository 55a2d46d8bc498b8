"""Tone mapping and image statistics (public names of the reference's torch_darktable/tonemap.py)."""

from dataclasses import dataclass

from beartype import beartype
import torch

from .extension import extension

_METRIC_KEYS = ('log_mean', 'linear_mean', 'rgb_mean')


@dataclass(frozen=True)
class TonemapParameters:
  """gamma, exposure `intensity` (stops), local/global `light_adapt` blend, `vibrance`."""

  gamma: float = 1.0
  intensity: float = 0.0
  light_adapt: float = 0.8
  vibrance: float = 0.0

  def to_cpp(self):
    return extension.TonemapParams(self.gamma, self.intensity, self.light_adapt, self.vibrance)

  @classmethod
  def from_cpp(cls, p) -> 'TonemapParameters':
    return cls(p.gamma, p.intensity, p.light_adapt, p.vibrance)


@beartype
def metrics_to_dict(metrics: torch.Tensor) -> dict[str, float | tuple[float, float, float]]:
  assert metrics.numel() == 5, f'Expected 5 elements, got {metrics.numel()}'
  m = [float(v) for v in metrics.detach().cpu().tolist()]
  return {'log_mean': m[0], 'linear_mean': m[1], 'rgb_mean': (m[2], m[3], m[4])}


@beartype
def metrics_from_dict(metrics_dict: dict[str, float | tuple[float, float, float]],
                      device: torch.device = torch.device('cuda')) -> torch.Tensor:
  rgb_mean = metrics_dict['rgb_mean']
  assert isinstance(rgb_mean, tuple), 'RGB mean must be a tuple'
  values = [metrics_dict['log_mean'], metrics_dict['linear_mean'], *rgb_mean]
  return torch.tensor(values, device=device, dtype=torch.float32)


@beartype
def print_metrics(metrics: torch.Tensor):
  d = metrics_to_dict(metrics)
  r, g, b = d['rgb_mean']
  print('Image Metrics:')
  print(f'  Log Mean: {d["log_mean"]:.4f}')
  print(f'  Linear Mean: {d["linear_mean"]:.4f}')
  print(f'  RGB Mean: ({r:.4f}, {g:.4f}, {b:.4f})')


def _check_hdr(image: torch.Tensor):
  assert image.dim() == 3 and image.size(2) == 3, 'Input must be (H, W, 3)'
  assert image.dtype == torch.float32, 'Input must be float32'
  assert image.device.type == 'cuda', 'Input must be on CUDA device'


def _check_metrics(metrics: torch.Tensor):
  assert metrics.numel() == 5 and metrics.dtype == torch.float32 and metrics.device.type == 'cuda', \
    'Metrics tensor must have 5 float32 elements on CUDA'


@beartype
def reinhard_tonemap(image: torch.Tensor, metrics: torch.Tensor, params: TonemapParameters) -> torch.Tensor:
  _check_hdr(image)
  assert metrics.numel() == 5, 'Metrics tensor must have 5 elements'
  return extension.reinhard_tonemap(image, metrics, params.to_cpp())


@beartype
def aces_tonemap(image: torch.Tensor, params: TonemapParameters, metrics: torch.Tensor | None = None) -> torch.Tensor:
  _check_hdr(image)
  if metrics is None:
    return extension.aces_tonemap(image, params.to_cpp())
  _check_metrics(metrics)
  return extension.adaptive_aces_tonemap(image, metrics, params.to_cpp())


@beartype
def linear_tonemap(image: torch.Tensor, metrics: torch.Tensor, params: TonemapParameters) -> torch.Tensor:
  _check_hdr(image)
  _check_metrics(metrics)
  return extension.linear_tonemap(image, metrics, params.to_cpp())


compute_image_bounds = beartype(extension.compute_image_bounds)


@beartype
def compute_image_metrics(images: list[torch.Tensor], stride: int = 8, min_gray: float = 1e-4, rescale: bool = False) -> torch.Tensor:
  return extension.compute_image_metrics(images, stride, min_gray, rescale)


__all__ = ['TonemapParameters', 'aces_tonemap', 'compute_image_bounds', 'compute_image_metrics', 'linear_tonemap',
           'metrics_from_dict', 'metrics_to_dict', 'print_metrics', 'reinhard_tonemap']
