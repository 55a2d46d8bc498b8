"""Pipeline settings (pydantic, frozen) with the same JSON schema as the reference's pipeline/config.py."""

from enum import Enum
from pathlib import Path
from typing import Annotated, Literal, get_args, get_origin

from beartype import beartype
from pydantic import BaseModel, GetCoreSchemaHandler
from pydantic_core import core_schema


class Validator:
  """A field annotation that validates (and documents) one setting."""

  description: str = ''

  def _validate(self, value):
    raise NotImplementedError

  def _serialize(self, value):
    return value

  def __get_pydantic_core_schema__(self, _source_type, _handler: GetCoreSchemaHandler):
    return core_schema.no_info_plain_validator_function(
      self._validate, serialization=core_schema.plain_serializer_function_ser_schema(self._serialize, when_used='always'))


class _Ranged(Validator):
  cast = float

  def __init__(self, range, description: str, step=None):
    self.range, self.description, self.step = range, description, step

  def _validate(self, value):
    value = self.cast(value)
    lo, hi = self.range
    if not lo <= value <= hi:
      raise ValueError(f'{value} not in [{lo}, {hi}]')
    return value


class Float(_Ranged):
  cast = float


class Int(_Ranged):
  cast = int


class Bool(Validator):
  def __init__(self, description: str):
    self.description = description

  def _validate(self, value):
    return bool(value)


class EnumValidator(Validator):
  """Accepts enum members, their names, or a {key: name} mapping; serialises to names."""

  def __init__(self, enum_type, description: str):
    self.enum_type, self.description = enum_type, description

  def _one(self, value):
    if isinstance(value, self.enum_type):
      return value
    if isinstance(value, str):
      return self.enum_type[value]
    raise ValueError(f'{value} is not a {self.enum_type.__name__}')

  def _validate(self, value):
    if isinstance(value, dict):
      return {k: self._one(v) for k, v in value.items()}
    return self._one(value)

  def _serialize(self, value):
    if isinstance(value, dict):
      return {k: v.name for k, v in value.items()}
    return value.name


def get_validator(model: type[BaseModel], field_name: str) -> Validator | None:
  annotation = model.__annotations__.get(field_name)
  if annotation is not None and get_origin(annotation) is Annotated:
    for extra in get_args(annotation)[1:]:
      if isinstance(extra, Validator):
        return extra
  return None


class ToneMapper(Enum):
  linear = 0
  reinhard = 1
  aces = 2
  adaptive_aces = 3


class Debayer(Enum):
  bilinear = 0
  ppg = 1
  rcd = 2


def clamp(x, lower, upper):
  return min(max(x, lower), upper)


class ImageProcessingSettings(BaseModel, frozen=True):
  type: Literal['image_processing_settings'] = 'image_processing_settings'

  tone_gamma: Annotated[float, Float(range=(0.1, 5.0), description='Gamma')] = 0.75
  tone_intensity: Annotated[float, Float(range=(-1.0, 5.0), description='Intensity')] = 2.0
  light_adapt: Annotated[float, Float(range=(0.0, 1.0), description='Light adaptation')] = 1.0
  vibrance: Annotated[float, Float(range=(-1.0, 1.0), description='Vibrance')] = 0.0
  # exponential moving average of bounds / metrics across image sets (1 = no smoothing)
  moving_average: Annotated[float, Float(range=(0.0, 1.0), description='Tonemap moving average')] = 0.02

  debayer: Annotated[Debayer, EnumValidator(Debayer, description='Debayer algorithm')] = Debayer.rcd
  ppg_median_threshold: float = 0.0

  postprocess: Annotated[bool, Bool(description='Postprocess debayer')] = False
  green_eq_threshold: float = 0.04
  color_smoothing_passes: int = 3

  enable_bilateral: Annotated[bool, Bool(description='Enable bilateral constrast enhancement')] = False
  bilateral: Annotated[float, Float(range=(0.0, 1.0), description='Bilateral constrast enhancement amount')] = 0.4
  bil_sigma_spatial: float = 2.0
  bil_sigma_luminance: float = 0.2

  enable_denoise: Annotated[bool, Bool(description='Enable denoise')] = True
  denoise: Annotated[float, Float(range=(0.0, 1.0), description='Denoise amount')] = 0.075

  tone_mapping: Annotated[ToneMapper, EnumValidator(ToneMapper, description='Tonemapping algorithm')] = ToneMapper.reinhard
  resize_width: Annotated[int, Int(range=(0, 4096), description='Resize width')] = 0

  @beartype
  def save_json(self, path: Path) -> None:
    path.write_text(self.model_dump_json(indent=2))

  @classmethod
  @beartype
  def load_json(cls, path: Path) -> 'ImageProcessingSettings':
    return cls.model_validate_json(path.read_text())
