"""Pins the CPU oracle against outputs of the reference extension itself (tests/golden/*.npz).

The reference has no golden vectors of its own; these were produced by
tests/golden/make_golden.py running baseline/_ref on a B200.  Tolerances live in tests/cases.py.
"""

import numpy as np
import pytest

import cases

ORACLE = cases.OracleImpl()
ALL = list(cases.all_cases())


@pytest.mark.parametrize('group,name,op,params,ins,outs', ALL, ids=[c[1] for c in ALL])
def test_oracle_matches_reference(group, name, op, params, ins, outs):
  got = cases.run_case(ORACLE, op, params, ins)
  problems = cases.check_outputs(op, got, outs, oracle=group == 'mid')
  assert not problems, f'{name}: ' + '; '.join(problems)


def test_reference_rcd_depends_on_workspace_history():
  """Documents SURVEY 8a6: the reference's own fresh-vs-used outputs differ inside the margin band only."""
  for _, name, op, _, _, outs in ALL:
    if op != 'rcd_reuse':
      continue
    d = np.abs(outs['out'] - outs['fresh']).max(axis=2)
    assert d.max() < cases.RCD_REUSE_BAND_TOL
    # only a 3-px band just inside the 7-px margin carries history; everything deeper is identical
    assert d[10:-10, 10:-10].max() == 0.0
    assert d[:7].max() == 0.0 and d[:, :7].max() == 0.0  # the PPG-style border never depends on it
