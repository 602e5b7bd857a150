for ds in 1 2 3 1 2 3; do python bench.py --steps 20 --warmup 3 --no-cpu-baseline --streams 3 --device-streams $ds 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('device-streams', sys.argv[1], round(d['value']/1e6,2), round(d['ms_per_step'],4), round(d['e2e']['value']/1e6,2))" $ds; done
