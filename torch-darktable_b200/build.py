"""Build libtdb200.so (the C-ABI CUDA library) in-tree for sm_100a with plain nvcc.

  python torch-darktable_b200/build.py [--force] [--verbose]

Objects go to torch-darktable_b200/build/, the library to torch-darktable_b200/torch_darktable/lib/libtdb200.so
(git-ignored, shipped to the GPU box by gpurun).  No torch headers are involved: the library has a C ABI.
"""

from __future__ import annotations

from concurrent.futures import ThreadPoolExecutor
import os
from pathlib import Path
import shutil
import subprocess
import sys

HERE = Path(__file__).resolve().parent
CSRC = HERE / 'csrc'
BUILD = HERE / 'build'
LIB = HERE / 'torch_darktable' / 'lib' / 'libtdb200.so'
INCLUDE = HERE.parent / 'include'

NVCC_FLAGS = [
  '-gencode', 'arch=compute_100a,code=sm_100a',
  '-O3', '--use_fast_math', '-lineinfo', '-std=c++17',
  '--expt-relaxed-constexpr', '--extended-lambda',
  '-Xcompiler', '-fPIC',
  '-I', str(INCLUDE),
]


def nvcc() -> str:
  for cand in (os.environ.get('NVCC'), shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
    if cand and Path(cand).exists():
      return cand
  raise RuntimeError('nvcc not found')


def sources() -> list[Path]:
  return sorted(CSRC.glob('*.cu'))


def _newest_header() -> float:
  hdrs = list(CSRC.glob('*.cuh')) + list(CSRC.glob('*.h')) + list(INCLUDE.glob('*.h'))
  return max(h.stat().st_mtime for h in hdrs)


def build(force: bool = False, verbose: bool = False) -> Path:
  BUILD.mkdir(exist_ok=True)
  LIB.parent.mkdir(parents=True, exist_ok=True)
  hdr_time = _newest_header()
  cc = nvcc()
  todo = []
  objs = []
  for src in sources():
    obj = BUILD / (src.stem + '.o')
    objs.append(obj)
    if force or not obj.exists() or obj.stat().st_mtime < max(src.stat().st_mtime, hdr_time):
      todo.append((src, obj))

  def compile_one(job):
    src, obj = job
    cmd = [cc, *NVCC_FLAGS, '-Xptxas', '-v' if verbose else '-warn-spills', '-c', str(src), '-o', str(obj)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    return src, res

  failed = False
  with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 2)) as ex:
    for src, res in ex.map(compile_one, todo):
      if verbose or res.returncode != 0 or 'warning' in res.stderr:
        sys.stderr.write(f'--- {src.name}\n{res.stdout}{res.stderr}')
      if res.returncode != 0:
        failed = True
  if failed:
    raise RuntimeError('nvcc failed')
  if todo or not LIB.exists():
    cmd = [cc, '-shared', '-o', str(LIB), *map(str, objs), '-gencode', 'arch=compute_100a,code=sm_100a']
    subprocess.run(cmd, check=True)
  return LIB


if __name__ == '__main__':
  path = build(force='--force' in sys.argv, verbose='--verbose' in sys.argv)
  print(path)
