#!/bin/bash
# launch list of the bench command under ncu (shares only; never a bench value)
mkdir -p gpurun_out/r2ap
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2ap/launches.csv python bench.py --steps 2 --warmup 1 --no-ncu > gpurun_out/r2ap/b.log 2>&1
python tools/launch_shares.py gpurun_out/r2ap/launches.csv | tee gpurun_out/r2ap/shares.txt
wc -l gpurun_out/r2ap/launches.csv
