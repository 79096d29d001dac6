mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_semiglobal_gpu.py -x -q 2>&1 | tail -3)
for P in 8192 18944; do
  timeout 300 python bench.py --workload semiglobal --no-cpu-baseline --pairs $P > gpurun_out/bench_sg2_$P.json 2> gpurun_out/bench_sg2_$P.err
  tail -c 300 gpurun_out/bench_sg2_$P.err
  python -c "
import json,sys; d=json.loads(open('gpurun_out/bench_sg2_$P.json').read().strip().splitlines()[-1]); print($P, round(d['value']), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_step'],2), d['verified'])"
done
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 12 --csv --log-file gpurun_out/sg2_launches.csv python bench.py --workload semiglobal --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
grep -o "sg[a-z0-9_]*kernel[^,]*,.*" gpurun_out/sg2_launches.csv | awk -F, '{print $1, $NF}' | head -3
