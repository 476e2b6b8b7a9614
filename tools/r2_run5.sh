#!/bin/bash
mkdir -p gpurun_out/r2e
O=gpurun_out/r2e
( timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.txt 2>&1; tail -6 $O/pytest_gpu.txt )
for b in 1 0; do
  VRDX_LOOK_BURST=$b timeout 600 python tools/shape_sweep.py --log2n 18 19 20 --algos 1 --shapes 0 2 5 6 --reps 9 > $O/small_burst$b.txt 2>&1
done
VRDX_LOOK_BURST=1 timeout 600 python tools/shape_sweep.py --log2n 21 22 23 24 --algos 1 --shapes 0 2 --reps 9 > $O/mid_burst1.txt 2>&1
VRDX_LOOK_BURST=0 timeout 600 python tools/shape_sweep.py --log2n 21 22 23 24 --algos 1 --shapes 0 2 --reps 9 > $O/mid_burst0.txt 2>&1
grep -H "2^" $O/small_*.txt $O/mid_*.txt | sed 's/gpurun_out.r2e.//' | sort -k7,7 -k8,8 -k3,3n | awk '{print $1,$3,$5,$7,$8,$9,$10,$11,$12}'
timeout 900 python bench.py --steps 5 --warmup 3 --no-cub > $O/bench.json 2> $O/bench.err; tail -c 300 $O/bench.err
python -c "
import json; d=json.load(open('$O/bench.json'))
print('value',d['value'],'e2e',d['e2e']['value'],'roofline',d['roofline']['frac'],d['roofline']['traffic'],d['roofline']['whole_sort']['traffic_over_algorithmic'])
print(d['extra'].get('key_value_2^28'))
"
