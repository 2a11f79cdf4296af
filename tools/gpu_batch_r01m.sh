set -x
mkdir -p gpurun_out
./tools/microbench/pipes > gpurun_out/pipes2.txt 2>&1
for m in 2 3 4; do echo "OLA_QUOT_MINB=$m"; OLA_QUOT_MINB=$m python tools/bench_prove.py 20 2>&1 | tail -1; done > gpurun_out/quot_sweep.txt 2>&1
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 12 -c 60 --csv --log-file gpurun_out/launches_r01m.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --prove-log-n 0 > gpurun_out/ncu_launch.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:tile_ --launch-skip 3 -c 3 -f -o gpurun_out/tile_r01n python bench.py --steps 1 --warmup 3 --no-cpu-baseline --prove-log-n 0 > gpurun_out/ncu_tile.log 2>&1
tail -2 gpurun_out/ncu_tile.log
cat gpurun_out/quot_sweep.txt
python -c "
import json
d=json.load(open('gpurun_out/bench_n1.json')); p=d['prove_all_tables']
print(d['value'], d['ms_per_step'], d['e2e'], d['roofline']['frac'], d['cpu_baseline'], p['seconds'], p['kernel_ms'], p['cpu_baseline'])
print(open('gpurun_out/bench_ref.json').read()[:600])"
