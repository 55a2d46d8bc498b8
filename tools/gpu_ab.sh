# A/B: side streams on / off, three runs each (value = MP/s resident, e2e)
for i in 1 2 3; do
  for mode in on off; do
    if [ $mode = on ]; then export TDB_SIDE_STREAM=1; else unset TDB_SIDE_STREAM; fi
    python bench.py --frames 16 --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$mode', d['value'], d['e2e']['value'])"
  done
done
