"""GPU: estimate_channel_noise (reference denoise.py:131-158) as a sampled kernel + exact on-device selection (csrc/noise.cu), against the
reference's own torch formulation (conv2d + strided slice + two torch.median) on the same device tensor, and a float64 numpy restatement.
Tolerance 1e-6: the response is five float32 terms whose order of addition differs between a direct loop and the library convolution; the
selection itself is exact (the result is one of the samples' values)."""

import numpy as np
import pytest

import synth

pytestmark = pytest.mark.gpu


def reference_formula(image, stride):
  import torch
  lap = torch.tensor([[0, -1, 0], [-1, 4, -1], [0, -1, 0]], dtype=image.dtype, device=image.device)
  hf = torch.conv2d(image.permute(2, 0, 1).unsqueeze(0), lap.unsqueeze(0).repeat(3, 1, 1).unsqueeze(1), groups=3, padding=1)
  flat = hf[0, :, ::stride, ::stride].flatten(1)
  med = torch.median(flat, dim=1).values
  return torch.median(torch.abs(flat - med.unsqueeze(1)), dim=1).values / 0.6745


@pytest.mark.parametrize('h,w,stride', [(192, 256, 8), (250, 372, 8), (251, 373, 4), (130, 202, 1), (2160, 3840, 8)])
def test_channel_noise_matches_the_reference_formula(h, w, stride):
  import torch
  import torch_darktable as td
  assert torch.cuda.is_available()
  rng = np.random.default_rng(h + w)
  sig = np.array([0.01, 0.03, 0.02], np.float32)
  img = (synth.scene_rgb(h, w, 3) + rng.normal(0, 1, (h, w, 3)).astype(np.float32) * sig).astype(np.float32)
  x = torch.from_numpy(img).cuda()
  got = td.estimate_channel_noise(x, stride=stride)
  assert got.is_cuda and got.shape == (3,) and got.dtype == torch.float32
  want = reference_formula(x, stride)
  assert float((got - want).abs().max()) <= 1e-6, (got.tolist(), want.tolist())
  # float64 restatement on the host (lower median, zero padding)
  p = np.pad(img.astype(np.float64), ((1, 1), (1, 1), (0, 0)))
  r = (4 * p[1:-1, 1:-1] - p[:-2, 1:-1] - p[2:, 1:-1] - p[1:-1, :-2] - p[1:-1, 2:])[::stride, ::stride].reshape(-1, 3)
  lower = lambda a: np.sort(a, axis=0)[(a.shape[0] - 1) // 2]  # noqa: E731
  med = lower(r)
  ref64 = lower(np.abs(r - med)) / 0.6745
  assert np.abs(got.cpu().numpy() - ref64).max() <= 2e-6


def test_channel_noise_recovers_injected_sigma():
  """A flat grey image + white noise of known sigma per channel: the response has gain sqrt(4^2 + 4) = sqrt(20)."""
  import torch
  import torch_darktable as td
  sig = np.array([0.01, 0.03, 0.02], np.float32)
  img = (0.5 + np.random.default_rng(2).normal(0, 1, (1080, 1920, 3)) * sig).astype(np.float32)
  got = td.estimate_channel_noise(torch.from_numpy(img).cuda(), stride=4).cpu().numpy() / np.sqrt(20.0)
  assert np.all(np.abs(got - sig) < 0.05 * sig), got


def test_channel_noise_feeds_wiener_without_a_host_round_trip():
  import torch
  import torch_darktable as td
  h, w = 192, 256
  x = torch.from_numpy((synth.scene_rgb(h, w, 5) + np.random.default_rng(1).normal(0, 0.02, (h, w, 3))).astype(np.float32)).cuda()
  sigma = td.estimate_channel_noise(x) / float(np.sqrt(20.0))
  out = td.Wiener(torch.device('cuda:0'), (w, h)).process(x, sigma)
  assert out.shape == x.shape and bool(torch.isfinite(out).all())
  assert float((out - x).abs().mean()) > 1e-4
