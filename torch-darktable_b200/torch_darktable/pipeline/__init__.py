"""RAW -> sRGB pipeline stage objects (public names of the reference's torch_darktable/pipeline package)."""

from .camera_settings import CameraSettings, load_camera_settings_from_dir, settings_for_file
from .config import Debayer, ImageProcessingSettings, ToneMapper
from .image_processor import ImageProcessor, ImageSizeMismatchError
from .presets import get_preset, presets
from .tiled import TiledFrameProcessor, halo_rows, partition_rows
from .transform import ImageTransform, transform, transformed_size

__all__ = ['CameraSettings', 'Debayer', 'ImageProcessingSettings', 'ImageProcessor', 'ImageSizeMismatchError', 'ImageTransform',
           'TiledFrameProcessor', 'ToneMapper', 'get_preset', 'halo_rows', 'load_camera_settings_from_dir', 'partition_rows', 'presets', 'settings_for_file', 'transform',
           'transformed_size']
