for p in 1,3,3,1 1,2,2,2,1 1,2,3,3,2,1 1,1,2,2,1,1 1,2,4,4,4,1 2,3,3; do
  echo "== $p"; BFL_PIPE=$p python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['e2e']['value'], d['value'])"
done
