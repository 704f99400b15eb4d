python bench.py --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['roofline']['stage_ms_one_step'])"
