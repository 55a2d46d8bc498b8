"""pytest configuration: marker registration and import paths.

`-m "not gpu"` runs here (no GPU): oracle vs golden vectors, host logic, C-ABI symbol checks, gloo sharding.
`-m gpu` runs on a B200: the parity tests proper, through the C ABI.
"""

from pathlib import Path
import sys

ROOT = Path(__file__).resolve().parents[1]
for p in (ROOT, ROOT / 'tests', ROOT / 'torch-darktable_b200'):
  if str(p) not in sys.path:
    sys.path.insert(0, str(p))


def pytest_configure(config):
  config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')
