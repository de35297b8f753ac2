for c in "32,96,128" "24,72,128" "48,80,128" "32,64,128" "16,48,96,96" "40,88,128"; do
  echo -n "chunks $c: "; CB_E2E_CHUNKS=$c timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-c1 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['e2e']['value']), round(d['e2e']['ms_per_step_device_events'],2))"
done
