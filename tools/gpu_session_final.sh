#!/bin/bash
# Final 1-GPU session of a round: the whole GPU test suite, the bench line (with traces), the launch list of the same command under
# ncu and one --set full capture of the probe kernel.
tag=${1:-r02h}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/${tag}_gpu.txt 2>&1
nproc > gpurun_out/${tag}_host.txt; cat /sys/fs/cgroup/cpu.max >> gpurun_out/${tag}_host.txt 2>/dev/null
( timeout 900 python -m pytest tests -m gpu -q -rs 2>&1 | tail -60 ) > gpurun_out/${tag}_pytest_gpu.log
tail -4 gpurun_out/${tag}_pytest_gpu.log
MASHGPU_TRACE=1 MASHGPU_TRACE_FEED=1 timeout 1200 python bench.py --steps 3 --warmup 3 > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err
echo "bench rc=$?"
MASHGPU_HOST_PACK=1 timeout 400 python bench.py --steps 3 --warmup 3 --skip-dist --skip-dist5 --skip-screen --skip-cpu > gpurun_out/${tag}_bench_feed_packed.json 2> gpurun_out/${tag}_bench_feed_packed.err
small="--steps 1 --warmup 1 --units 400 --sketches 100000 --sketches5 200000 --reads 4000000 --skip-cpu --skip-e2e"
mine='regex:scan_|select_|quirk_|tile_tmax|apply_runs|write_separators|dist_|dict_|list_gather|iota_|screen_|merge_bottom|DeviceRadixSort|DeviceScan|DeviceSegmented|DeviceSelect|reads_replay'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$mine" -c 3000 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py $small > gpurun_out/${tag}_ncu_launches.log 2>&1
dsmall="--steps 1 --warmup 1 --units 50 --sketches 100000 --skip-screen --skip-cpu --skip-e2e --skip-dist5"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dist_probe_kernel -s 1 -c 1 -o gpurun_out/${tag}_probe python bench.py $dsmall > gpurun_out/${tag}_ncu_probe.log 2>&1
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/*_bench_n1.json'))[-1:] + sorted(glob.glob('gpurun_out/*_bench_feed_packed.json'))[-1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value %.1f e2e %.1f' % (d['value'], d['e2e']['value']), (d.get('screen') or {}).get('e2e'))
    except Exception as e:
        print(f, e)
PY
grep "screen packed feed" gpurun_out/${tag}_bench_n1.err | sed -n '30,36p'
