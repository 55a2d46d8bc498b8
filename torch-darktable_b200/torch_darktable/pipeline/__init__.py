"""RAW -> sRGB pipeline stage objects (public names of the reference's torch_darktable/pipeline package)."""

from .camera_settings import CameraSettings, load_camera_settings_from_dir, settings_for_file
from .config import Debayer, ImageProcessingSettings, ToneMapper
from .image_processor import ImageProcessor, ImageSizeMismatchError
from .presets import get_preset, presets
from .transform import ImageTransform, transform, transformed_size

__all__ = ['CameraSettings', 'Debayer', 'ImageProcessingSettings', 'ImageProcessor', 'ImageSizeMismatchError', 'ImageTransform',
           'ToneMapper', 'get_preset', 'load_camera_settings_from_dir', 'presets', 'settings_for_file', 'transform',
           'transformed_size']
