"""JPEG output (`torch_darktable.Jpeg`, reference jpeg.py:23-29 / csrc/jpeg_encoder.cu:104-180).

The encoder is nvJPEG in both packages, so the parity bar is the strongest one: the stream must equal the reference's byte for
byte (tests/golden/jpeg.npz, written by tests/golden/make_golden.py from the unmodified reference on a B200).  Independently of
the fixture, every stream must decode (OpenCV's libjpeg) to the input picture within the PSNR a JPEG of that quality gives."""

import json
from pathlib import Path

import numpy as np
import pytest

import synth

GOLDEN = Path(__file__).resolve().parent / 'golden' / 'jpeg.npz'


def picture(h, w, seed):
  return np.floor(synth.scene_rgb(h, w, seed) ** (1 / 2.2) * 255.0 + 0.5).clip(0, 255).astype(np.uint8)


def psnr(a, b):
  mse = np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2)
  return 10 * np.log10(255.0 ** 2 / max(mse, 1e-12))


def test_enum_values_are_the_references():
  import torch_darktable as td
  ext = td.extension.extension
  assert [int(ext.JpegInputFormat[n]) for n in ('BGR', 'RGB', 'BGRI', 'RGBI')] == [0, 1, 2, 3]  # csrc/jpeg_encoder.h:6-11
  assert [int(ext.JpegSubsampling[n]) for n in ('CSS_444', 'CSS_422', 'CSS_GRAY')] == [0, 1, 2]  # :13-17
  assert int(ext.RGBI) == 3 and int(ext.CSS_422) == 1  # export_values()
  assert int(td.jpeg.InputFormat.RGBI) == 3 and int(td.jpeg.Subsampling.CSS_GRAY) == 2
  assert issubclass(td.JpegException, Exception)


def test_no_gpu_means_an_exception_not_a_fallback():
  import torch
  import torch_darktable as td
  if torch.cuda.is_available():
    pytest.skip('needs a machine without a GPU')
  with pytest.raises((td.JpegException, RuntimeError)):
    td.Jpeg()


def decode(stream):
  import cv2
  img = cv2.imdecode(np.frombuffer(stream, dtype=np.uint8), cv2.IMREAD_COLOR)
  assert img is not None, 'stream does not decode'
  return img[:, :, ::-1]  # BGR -> RGB


@pytest.mark.gpu
@pytest.mark.parametrize('fmt,quality,css,progressive,min_psnr', [
  ('RGBI', 94, 'CSS_422', False, 36.0), ('RGBI', 80, 'CSS_444', True, 33.0), ('BGRI', 94, 'CSS_444', False, 38.0),
  ('RGB', 90, 'CSS_422', False, 34.0), ('BGR', 60, 'CSS_444', False, 30.0)])
def test_stream_decodes_to_the_input(fmt, quality, css, progressive, min_psnr):
  import torch
  import torch_darktable as td
  img = picture(200, 312, 31)
  src = img if fmt in ('RGBI', 'RGB') else img[:, :, ::-1]
  x = torch.from_numpy(np.ascontiguousarray(src if fmt.endswith('I') else src.transpose(2, 0, 1))).cuda()
  stream = td.Jpeg().encode(x, quality=quality, input_format=td.jpeg.InputFormat[fmt], subsampling=td.jpeg.Subsampling[css],
                            progressive=progressive)
  assert stream.device.type == 'cpu' and stream.dtype == torch.uint8 and stream.dim() == 1
  raw = stream.numpy().tobytes()
  assert raw[:2] == b'\xff\xd8' and raw[-2:] == b'\xff\xd9'
  assert (b'\xff\xc2' in raw) == progressive  # SOF2 marks a progressive stream
  got = decode(raw)
  assert got.shape == img.shape
  assert psnr(got, img) >= min_psnr, psnr(got, img)


@pytest.mark.gpu
def test_gray_subsampling_and_pitched_rows():
  import torch
  import torch_darktable as td
  img = picture(96, 160, 37)
  coder = td.Jpeg()
  gray = decode(coder.encode(torch.from_numpy(img).cuda(), subsampling=td.jpeg.Subsampling.CSS_GRAY).numpy().tobytes())
  luma = img.astype(np.float64) @ np.array([0.299, 0.587, 0.114])
  assert np.abs(gray.astype(np.float64)[:, :, 0] - luma).mean() < 2.0 and np.ptp(gray.astype(np.int32), axis=2).max() == 0
  # a view with a row pitch (left part of a wider image) is encoded in place and equals the dense copy's stream
  wide = torch.from_numpy(np.concatenate([img, img[:, ::-1]], axis=1)).cuda()
  view = wide[:, :160]
  assert not view.is_contiguous()
  a = coder.encode(view, quality=90, subsampling=td.jpeg.Subsampling.CSS_444)
  b = coder.encode(view.contiguous(), quality=90, subsampling=td.jpeg.Subsampling.CSS_444)
  assert torch.equal(a, b)


@pytest.mark.gpu
def test_pipeline_result_feeds_the_encoder():
  """The uint8 result of ImageProcessor.process (transformed by the tone-map kernel) goes straight into Jpeg.encode."""
  import torch
  import torch_darktable as td
  from torch_darktable.pipeline import ImageProcessingSettings, ImageProcessor, ImageTransform
  w, h = 256, 192
  frame = synth.packed_frame(h, w, seed=5)
  proc = ImageProcessor((w, h), td.BayerPattern.RGGB, td.PackedFormat.Packed12, ImageProcessingSettings(moving_average=1.0),
                        torch.device('cuda:0'), (1.8, 1.0, 2.1), ImageTransform.rotate_90)
  out = proc.process(torch.from_numpy(frame).cuda(), 'cam')
  got = decode(td.Jpeg().encode(out, quality=95, subsampling=td.jpeg.Subsampling.CSS_444).numpy().tobytes())
  assert got.shape == tuple(out.shape) == (w, h, 3)
  assert psnr(got, out.cpu().numpy()) >= 36.0


@pytest.mark.gpu
def test_argument_errors():
  import torch
  import torch_darktable as td
  coder = td.Jpeg()
  with pytest.raises(RuntimeError, match='uint8'):
    coder.encode(torch.zeros(8, 8, 3, device='cuda'))
  with pytest.raises(RuntimeError, match='CUDA'):
    coder.encode(torch.zeros(8, 8, 3, dtype=torch.uint8))
  with pytest.raises(RuntimeError, match='interleaved'):
    coder.encode(torch.zeros(3, 8, 8, dtype=torch.uint8, device='cuda'))
  with pytest.raises(RuntimeError, match='planar'):
    coder.encode(torch.zeros(8, 8, 3, dtype=torch.uint8, device='cuda'), input_format=td.jpeg.InputFormat.RGB)
  with pytest.raises(RuntimeError, match='quality'):
    coder.encode(torch.zeros(8, 8, 3, dtype=torch.uint8, device='cuda'), quality=0)


@pytest.mark.gpu
def test_streams_equal_the_references():
  import torch
  import torch_darktable as td
  if not GOLDEN.exists():
    pytest.skip('tests/golden/jpeg.npz not generated yet (python tests/golden/make_golden.py <dir> jpeg on a GPU box)')
  data = np.load(GOLDEN)
  manifest = json.loads(str(data['manifest']))
  coder = td.Jpeg()
  assert manifest
  for name, case in manifest.items():
    p = case['params']
    img = picture(p['synth']['h'], p['synth']['w'], p['synth']['seed'])
    x = torch.from_numpy(np.ascontiguousarray(img if p['input_format'].endswith('I') else img.transpose(2, 0, 1))).cuda()
    got = coder.encode(x, quality=p['quality'], input_format=td.jpeg.InputFormat[p['input_format']],
                       subsampling=td.jpeg.Subsampling[p['subsampling']], progressive=p['progressive']).numpy()
    want = data[f'{name}/out/out']
    assert got.shape == want.shape and np.array_equal(got, want), f'{name}: stream differs from the reference ({got.size} vs {want.size} bytes)'
