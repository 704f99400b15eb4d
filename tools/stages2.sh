# device-resident and end-to-end throughput only (stage profile forces one stream)
python bench.py --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('value', d['value'], 'e2e', d['e2e']['value'], 'ms/step', d['ms_per_step'])"
