for s in "32,64,128,256,352,192" "64,128,256,576" "64,160,320,480" "96,224,704" "48,96,192,384,304" "128,256,640" "64,192,384,384"; do
  VB_PAIRS_SCHEDULE=$s python bench.py --steps 6 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$s', round(d['e2e']['value']), round(d['value']))"
done
