#!/bin/bash
# Short 1-GPU session: the CLI read-filter tests, the dist A/B (pair list on/off) and the default bench line.
tag=${1:-r02d}
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_cli.py tests/test_gpu_dist.py tests/test_gpu_dist_prefilter.py -m gpu -q -rs 2>&1 | tail -30 ) > gpurun_out/${tag}_pytest_gpu.log
tail -4 gpurun_out/${tag}_pytest_gpu.log
donly="--steps 2 --warmup 1 --units 50 --skip-screen --skip-cpu --skip-e2e --skip-dist5"
MASHGPU_DIST_PAIR_MAX=0 timeout 400 python bench.py $donly > gpurun_out/${tag}_ab_pairs_off.json 2> gpurun_out/${tag}_ab_pairs_off.err
MASHGPU_DIST_PAIR_MAX=4 timeout 400 python bench.py $donly > gpurun_out/${tag}_ab_pairs_4.json 2> gpurun_out/${tag}_ab_pairs_4.err
dsmall="--steps 1 --warmup 1 --units 50 --sketches 100000 --skip-screen --skip-cpu --skip-e2e --skip-dist5"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dist_probe_kernel -s 1 -c 1 -o gpurun_out/${tag}_probe python bench.py $dsmall > gpurun_out/${tag}_ncu_probe.log 2>&1
MASHGPU_TRACE=1 timeout 1200 python bench.py --steps 3 --warmup 3 > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err
echo "bench rc=$?"; tail -3 gpurun_out/${tag}_bench_n1.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02e_ab_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, json.dumps(d.get('dist')))
    except Exception as e: print(f, e)
PY
