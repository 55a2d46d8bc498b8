"""torch_darktable, B200-native: the RAW -> sRGB ISP hot path as hand-written sm_100a CUDA behind a C ABI.

Drop-in for the Python API of uc-vision/torch-darktable (same module and function names); see extension.py for the
boundary and csrc/ for the kernels.  There is no CPU fallback: importing this package without the built library fails.
"""

import importlib as _importlib

__version__ = '0.2.3+b200'

# public surface: sub-module -> names re-exported at package level (the reference's torch_darktable/__init__.py, plus
# demosaic_packed, the fused ingest this package adds)
_SURFACE = {
  'extension': (),
  'bayer': ('BayerPattern', 'PackedFormat', 'load_as_bayer', 'rgb_to_bayer'),
  'color_conversion': ('color_transform_3x3', 'compute_log_luminance', 'compute_luminance', 'lab_to_rgb', 'lab_to_xyz', 'modify_hsl',
                       'modify_log_luminance', 'modify_luminance', 'modify_vibrance', 'rgb_to_lab', 'rgb_to_xyz', 'xyz_to_lab', 'xyz_to_rgb'),
  'debayer': ('PPG', 'RCD', 'Bilinear5x5', 'PostProcess', 'bilinear5x5_demosaic', 'decode12', 'decode12_float', 'decode12_half',
              'decode12_u16', 'demosaic_packed', 'encode', 'encode12_float', 'encode12_u16'),
  'denoise': ('Wiener', 'estimate_channel_noise'),
  'jpeg': ('InputFormat', 'Jpeg', 'JpegException', 'Subsampling'),
  'local_contrast': ('Bilateral', 'Laplacian', 'LaplacianParams'),
  'tonemap': ('TonemapParameters', 'aces_tonemap', 'compute_image_bounds', 'compute_image_metrics', 'linear_tonemap',
              'metrics_from_dict', 'metrics_to_dict', 'print_metrics', 'reinhard_tonemap'),
  'white_balance': ('apply_white_balance', 'estimate_white_balance'),
}

__all__ = []
for _module_name, _names in _SURFACE.items():
  _module = _importlib.import_module(f'.{_module_name}', __name__)
  globals()[_module_name] = _module
  __all__.append(_module_name)
  for _name in _names:
    globals()[_name] = getattr(_module, _name)
    __all__.append(_name)
__all__.sort()
del _module_name, _names, _module, _name
