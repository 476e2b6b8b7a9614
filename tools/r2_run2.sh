#!/bin/bash
# round-2 GPU call 2: relaxed ranking A/B, full GPU test-suite (new interop / thread / bench_cpp tests), bench.py both arms
mkdir -p gpurun_out/r2b
O=gpurun_out/r2b
( timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.txt 2>&1; tail -5 $O/pytest_gpu.txt )
for r in 1 0; do
  VRDX_RTS_PERSISTENT=0 VRDX_RELAXED=$r timeout 600 python tools/shape_sweep.py --log2n 25 28 --shapes 1 2 --kinds keys --algos 2 > $O/sweep_relaxed$r.txt 2>&1
done
grep -h "2^2[58]" $O/sweep_relaxed*.txt
for p in 1 0; do
  VRDX_RTS_PERSISTENT=$p timeout 600 python tools/shape_sweep.py --log2n 25 26 28 --algos 2 > $O/sweep_persistent$p.txt 2>&1
done
grep -h "2^2[568]" $O/sweep_persistent*.txt | sort -k3,3n -k8,8 -k9,9
ls /usr/share/vulkan/icd.d /etc/vulkan/icd.d 2>&1 | head; ldconfig -p | grep -i vulkan
timeout 900 python bench.py --steps 10 --warmup 3 > $O/bench.json 2> $O/bench.err; tail -c 600 $O/bench.err
python -c "
import json; d=json.load(open('$O/bench.json'))
print('value',d['value'],'e2e',d['e2e']['value'],'roofline',d['roofline']['frac'],d['roofline']['traffic'],d['roofline']['whole_sort'])
for k,v in d['extra'].items(): print(k, v if not isinstance(v,dict) or len(str(v))<400 else str(v)[:400])
"
timeout 900 python bench.py --impl reference --steps 10 --warmup 3 > $O/bench_reference.json 2> $O/bench_reference.err; cat $O/bench_reference.json | cut -c1-600
