"""Per-camera settings (sensor size, CFA, packing, white balance, orientation) and raw-file loading."""

from pathlib import Path
from typing import Annotated, Literal
import warnings

from beartype import beartype
from pydantic import BaseModel
import torch

from .. import bayer, debayer
from .config import EnumValidator, ImageProcessingSettings
from .transform import ImageTransform

warnings.filterwarnings('ignore', category=UserWarning, message='The given buffer is not writable')


@beartype
class CameraSettings(BaseModel, frozen=True):
  type: Literal['camera_settings'] = 'camera_settings'

  name: str
  image_size: tuple[int, int]  # (width, height)
  padding: int = 0             # trailing bytes after the packed frame

  bayer_pattern: Annotated[bayer.BayerPattern, EnumValidator(bayer.BayerPattern, 'Bayer pattern')] = bayer.BayerPattern.RGGB
  packed_format: Annotated[bayer.PackedFormat, EnumValidator(bayer.PackedFormat, 'Packed format')] = bayer.PackedFormat.Packed12
  white_balance: tuple[float, float, float] | None = None
  image_processing: ImageProcessingSettings

  transform: Annotated[ImageTransform | dict[str, ImageTransform], EnumValidator(ImageTransform, 'Image transform')] = ImageTransform.none

  def get_image_transform(self, camera_name: str) -> ImageTransform:
    if isinstance(self.transform, dict):
      return self.transform.get(camera_name, ImageTransform.none)
    return self.transform

  @property
  def bytes(self) -> int:
    width, height = self.image_size
    return (width * height * 3) // 2 + self.padding

  @beartype
  def save_json(self, path: Path) -> None:
    path.write_text(self.model_dump_json(indent=2))

  @classmethod
  @beartype
  def load_json(cls, path: Path) -> 'CameraSettings':
    return cls.model_validate_json(path.read_text())


@beartype
def load_raw_bytes(filepath: Path, device: torch.device = torch.device('cuda:0')):
  """File contents as a uint8 device tensor, undecoded: read straight into pinned host memory, then one asynchronous copy on the
  current stream (the pinned block is recycled by torch's host allocator once that copy has run)."""
  host = torch.empty(filepath.stat().st_size, dtype=torch.uint8, pin_memory=device.type == 'cuda')
  with open(filepath, 'rb') as f:
    f.readinto(host.numpy())
  return host.to(device, non_blocking=True)


@beartype
def load_raw_bytes_stripped(filepath: Path, camera_settings: CameraSettings, device: torch.device = torch.device('cuda:0')):
  raw = load_raw_bytes(filepath, device)
  return raw[: -camera_settings.padding] if camera_settings.padding > 0 else raw


def load_raw_bayer(filepath: Path, camera_settings: CameraSettings | None = None,
                   device: torch.device = torch.device('cuda:0')) -> torch.Tensor:
  if camera_settings is None:
    camera_settings = settings_for_file(filepath)
  raw = load_raw_bytes_stripped(filepath, camera_settings, device)
  decoded = debayer.decode12(raw, output_dtype=torch.float32, format_type=camera_settings.packed_format)
  return decoded.view(-1, camera_settings.image_size[0])


def get_camera_settings_dir() -> Path:
  return Path(__file__).resolve().parent.parent / 'camera_settings'


def load_camera_settings_from_dir(settings_dir: Path | None = None) -> dict[str, CameraSettings]:
  settings_dir = settings_dir or get_camera_settings_dir()
  loaded = (CameraSettings.load_json(p) for p in sorted(settings_dir.glob('*.json')))
  return {s.name: s for s in loaded}


def settings_for_file(file_path: Path) -> CameraSettings:
  """Camera settings by parent directory name, else by file size."""
  known = load_camera_settings_from_dir()
  camera_name = file_path.parent.stem
  if camera_name in known:
    return known[camera_name]
  size = file_path.stat().st_size
  for candidate in known.values():
    if candidate.bytes == size:
      return candidate
  raise ValueError(f'Could not find camera settings for "{file_path}". Directory name "{camera_name}" not recognized and file size '
                   f'{size} bytes does not match any known camera. Available cameras: {list(known.keys())}')


@beartype
def validate_camera_names(settings: CameraSettings, camera_names: list[str]) -> None:
  if isinstance(settings.transform, dict):
    expected, actual = set(settings.transform.keys()), set(camera_names)
    if expected != actual:
      raise ValueError(f'Camera names mismatch: settings expects {sorted(expected)}, got {sorted(actual)}')
