"""GPU: ImageProcessor.process_batch -- the (N, bytes) batch entry and its CUDA-graph form.

A batch is N image sets of one frame each, i.e. N calls of ImageProcessor.process (reference pipeline/image_processor.py:274-300)
with the EMA of bounds / metrics carried from frame to frame.  The graph form must be indistinguishable from the eager calls, on the
capturing call AND on replays with new input.  "Indistinguishable" = what two eager runs of the same frame guarantee: the Wiener
overlap-add accumulates with float atomics whose order varies from run to run (as in the reference, denoise.cu:173-177), so a
uint8 sample may differ by one LSB on a few samples per million; the tests allow 1 LSB on at most 1e-4 of the samples and 1e-6 on
the EMA state."""

import numpy as np
import pytest

import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def td():
  import torch
  assert torch.cuda.is_available(), 'GPU tests need a CUDA device'
  import torch_darktable
  return torch_darktable


def make_processor(td, w, h, ma, transform='rotate_270', debayer='rcd'):
  import torch
  from torch_darktable.pipeline import ImageProcessingSettings, ImageProcessor, ImageTransform
  from torch_darktable.pipeline.config import Debayer, ToneMapper
  settings = ImageProcessingSettings(debayer=Debayer[debayer], tone_mapping=ToneMapper.adaptive_aces, enable_denoise=True, enable_bilateral=True,
                                     postprocess=True, tone_gamma=1.5, tone_intensity=2.0, light_adapt=0.8, vibrance=0.5, moving_average=ma)
  return ImageProcessor((w, h), td.BayerPattern.RGGB, td.PackedFormat.Packed12, settings, torch.device('cuda:0'), (1.8, 1.0, 2.1),
                        ImageTransform[transform])


def assert_same(got, want, what=''):
  import torch
  d = (got.to(torch.int16) - want.to(torch.int16)).abs()
  assert int(d.max()) <= 1 and float((d > 0).float().mean()) <= 1e-4, f'{what}: max {int(d.max())} LSB, {float((d > 0).float().mean()):.2e} of the samples differ'


def frames_of(h, w, seeds, gain=1.0):
  import torch
  fr = []
  for s in seeds:
    cfa = synth.mosaic(synth.scene_rgb(h, w, s), 'RGGB') * (gain * (0.6 + 0.1 * (s % 5)))  # exposures differ: the EMA has something to do
    fr.append(synth.pack12(np.floor(np.clip(cfa, 0, 1) * 4095.0 + 0.5).astype(np.uint16)))
  return torch.from_numpy(np.stack(fr)).cuda()


@pytest.mark.parametrize('h,w,transform', [(192, 256, 'rotate_270'), (250, 372, 'none')])
@pytest.mark.parametrize('ma', [1.0, 0.3])
def test_batch_graph_equals_eager_calls(td, h, w, transform, ma):
  import torch
  batches = [frames_of(h, w, range(10 * b, 10 * b + 4)) for b in range(3)]
  eager = make_processor(td, w, h, ma, transform)
  want = [[eager.process(f, 'cam') for f in batch] for batch in batches]
  want_state = (eager.bounds.clone(), eager.metrics.clone())

  graphed = make_processor(td, w, h, ma, transform)
  for b, batch in enumerate(batches):  # call 0 runs eagerly and captures, calls 1 and 2 replay with new input
    got = graphed.process_batch(batch, 'cam', graph=True)
    assert got.shape[0] == 4 and got.dtype == torch.uint8
    for i in range(4):
      assert_same(got[i], want[b][i], f'batch {b} frame {i}: graph vs eager')
  assert graphed._batch_graph is not None
  assert torch.allclose(graphed.bounds, want_state[0], atol=1e-6) and torch.allclose(graphed.metrics, want_state[1], atol=1e-6)

  plain = make_processor(td, w, h, ma, transform)
  for b, batch in enumerate(batches):
    got = plain.process_batch(batch, 'cam', graph=False)
    for i in range(4):
      assert_same(got[i], want[b][i], f'batch {b} frame {i}: batch vs per-frame calls')


def test_batch_input_buffer_is_read_in_place(td):
  """Filling the graph's own input buffer saves the device-to-device copy; the output buffer is reused by the next call."""
  import torch
  h, w = 192, 256
  proc = make_processor(td, w, h, 1.0)
  first = frames_of(h, w, [1, 2])
  out0 = proc.process_batch(first, 'cam').clone()
  buf = proc.batch_input_buffer(2)
  assert buf.data_ptr() == proc._batch_graph[2].data_ptr()
  second = frames_of(h, w, [3, 4])
  buf.copy_(second)
  out1 = proc.process_batch(buf, 'cam')
  ref = make_processor(td, w, h, 1.0)
  assert_same(out1[0], ref.process(second[0], 'cam')), assert_same(out1[1], ref.process(second[1], 'cam'))
  assert not torch.equal(out0, out1)
  with pytest.raises(Exception):
    proc.process_batch(second[:, :-3], 'cam')


def test_state_from_stage_path_carries_into_the_fused_path(td):
  """bounds / metrics set by the stage-by-stage composite (new tensors) are picked up by the in-place state of the fused path."""
  import torch
  h, w = 192, 256
  fr = frames_of(h, w, [5, 6, 7])
  a, b = make_processor(td, w, h, 0.4), make_processor(td, w, h, 0.4)
  a.process_image_set_by_stage({'cam': fr[0]})
  b.process_image_set({'cam': fr[0]})
  ra = a.process_image_set({'cam': fr[1]})['cam']
  rb = b.process_image_set({'cam': fr[1]})['cam']
  d = (ra.to(torch.int16) - rb.to(torch.int16)).abs()
  assert int(d.max()) <= 1 and float((d > 0).float().mean()) <= 1e-3
  assert torch.allclose(a.bounds, b.bounds, atol=2e-6) and torch.allclose(a.metrics, b.metrics, atol=2e-5)


@pytest.mark.parametrize('ma', [1.0, 0.3])
def test_submit_two_frames_in_flight_equals_sequential_process(td, ma):
  """`submit` overlaps consecutive frames on two lanes; the EMA chain must stay that of sequential `process` calls, also when eager
  calls are interleaved (they join the lanes and take the state over) and when the caller's tensors are short-lived."""
  import torch
  h, w = 250, 372
  fr = frames_of(h, w, range(30, 39))
  seq = make_processor(td, w, h, ma, 'none')
  want = [seq.process(f, 'cam') for f in fr]
  ovl = make_processor(td, w, h, ma, 'none')
  got = []
  for i, f in enumerate(fr):
    if i == 4:  # an eager call in the middle
      got.append(ovl.process(f.clone(), 'cam'))
    else:
      res, done = ovl.submit(f.clone(), 'cam')  # the clone dies right after the call: the lane must still read valid bytes
      got.append(res)
  ovl.join()
  torch.cuda.synchronize()
  for i, (g, wv) in enumerate(zip(got, want)):
    assert_same(g, wv, f'frame {i}: submit vs process')
  assert torch.allclose(ovl.bounds, seq.bounds, atol=1e-6) and torch.allclose(ovl.metrics, seq.metrics, atol=1e-6)
  assert ovl.bounds is ovl._bounds_pp[0]  # join() leaves the state in its resting buffers


def test_host_frame_runner_matches_process(td):
  """The three-stream host runner (pinned in / out, two frames in flight on the lanes) against per-frame process calls."""
  import torch
  from torch_darktable.pipeline.batch import HostFrameRunner
  h, w = 250, 372
  fr = frames_of(h, w, range(50, 57))
  ref = make_processor(td, w, h, 0.5, 'rotate_270')
  want = [ref.process(f, 'cam').cpu() for f in fr]
  proc = make_processor(td, w, h, 0.5, 'rotate_270')
  runner = HostFrameRunner(proc)
  host_in = [f.cpu().pin_memory() for f in fr]
  host_out = [torch.empty((w, h, 3), dtype=torch.uint8).pin_memory() for _ in fr]
  runner.run(host_in[:4], host_out[:4])
  runner.run(host_in[4:], host_out[4:], after_caller=False)
  runner.wait()
  torch.cuda.synchronize()
  for i, (g, wv) in enumerate(zip(host_out, want)):
    assert_same(g.cuda(), wv.cuda(), f'frame {i}: host runner vs process')


@pytest.mark.parametrize('kw', [dict(postprocess=False), dict(enable_denoise=False), dict(enable_bilateral=False), dict(enable_denoise=False, enable_bilateral=False)])
def test_batch_graph_with_other_stage_combinations(td, kw):
  """Settings that leave the nine-launch configuration (no post-process: frames run in line; no denoise / no bilateral: shorter lanes)
  go through the same batch entry and graph."""
  import torch
  from torch_darktable.pipeline import ImageProcessingSettings, ImageProcessor, ImageTransform
  from torch_darktable.pipeline.config import Debayer, ToneMapper
  h, w = 192, 256
  base = dict(debayer=Debayer.ppg, tone_mapping=ToneMapper.reinhard, enable_denoise=True, enable_bilateral=True, postprocess=True,
              tone_gamma=1.5, tone_intensity=2.0, light_adapt=0.8, vibrance=0.5, moving_average=0.5)
  base.update(kw)
  mk = lambda: ImageProcessor((w, h), td.BayerPattern.RGGB, td.PackedFormat.Packed12, ImageProcessingSettings(**base), torch.device('cuda:0'),  # noqa: E731
                              (1.8, 1.0, 2.1), ImageTransform.flip_horiz)
  batches = [frames_of(h, w, range(60 + 3 * b, 63 + 3 * b)) for b in range(3)]
  seq, bat = mk(), mk()
  for batch in batches:
    want = [seq.process(f, 'cam') for f in batch]
    got = bat.process_batch(batch, 'cam', graph=True)
    for i in range(3):
      assert_same(got[i], want[i], f'{kw} frame {i}')
  assert torch.allclose(bat.bounds, seq.bounds, atol=1e-6) and torch.allclose(bat.metrics, seq.metrics, atol=1e-6)
