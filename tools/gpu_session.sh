#!/bin/bash
# One GPU session under gpurun (1 GPU): parity tests, the bench line, feed-path A/B, ncu evidence.
#   gpurun --timeout 2400 -- 'bash tools/gpu_session.sh r02a'
# Everything lands in gpurun_out/<tag>_*; summaries worth keeping are copied to profiles/ by hand afterwards.
tag=${1:-r02}
what=${2:-all}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/${tag}_gpu.txt 2>&1
nproc > gpurun_out/${tag}_host.txt; cat /sys/fs/cgroup/cpu.max >> gpurun_out/${tag}_host.txt 2>/dev/null; free -g >> gpurun_out/${tag}_host.txt

if [ "$what" = all ] || [ "$what" = tests ]; then
  ( timeout 900 python -m pytest tests -m gpu -q -rs 2>&1 | tail -80 ) > gpurun_out/${tag}_pytest_gpu.log
  tail -6 gpurun_out/${tag}_pytest_gpu.log
fi
if [ "$what" = all ] || [ "$what" = bench ]; then
  MASHGPU_TRACE=1 timeout 1200 python bench.py --steps 3 --warmup 3 > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err
  echo "bench rc=$?"; tail -3 gpurun_out/${tag}_bench_n1.err
  timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err
  MASHGPU_HOST_PACK=0 timeout 400 python bench.py --steps 3 --warmup 3 --skip-dist --skip-cpu > gpurun_out/${tag}_bench_feed_ascii.json 2> gpurun_out/${tag}_bench_feed_ascii.err
  MASHGPU_HOST_PACK=1 timeout 400 python bench.py --steps 3 --warmup 3 --skip-dist --skip-cpu > gpurun_out/${tag}_bench_feed_packed.json 2> gpurun_out/${tag}_bench_feed_packed.err
fi
if [ "$what" = all ] || [ "$what" = ab ]; then
  # A/B runs of the dist path (dist only, configs[2]): TMA bulk-copy staging of the query rows, pair list on/off
  donly="--steps 2 --warmup 1 --units 50 --skip-screen --skip-cpu --skip-e2e --skip-dist5"
  if [ "$tag" = r02b ]; then
  MASHGPU_DIST_PREFILTER=0 MASHGPU_DIST_BULK=0 timeout 400 python bench.py $donly --sketches 30000 > gpurun_out/${tag}_ab_merge_ldg.json 2> gpurun_out/${tag}_ab_merge_ldg.err
  MASHGPU_DIST_PREFILTER=0 MASHGPU_DIST_BULK=1 timeout 400 python bench.py $donly --sketches 30000 > gpurun_out/${tag}_ab_merge_bulk.json 2> gpurun_out/${tag}_ab_merge_bulk.err
  fi
  MASHGPU_DIST_PAIR_MAX=0 timeout 400 python bench.py $donly > gpurun_out/${tag}_ab_pairs_off.json 2> gpurun_out/${tag}_ab_pairs_off.err
  MASHGPU_DIST_PAIR_MAX=4 timeout 400 python bench.py $donly > gpurun_out/${tag}_ab_pairs_4.json 2> gpurun_out/${tag}_ab_pairs_4.err
fi
if [ "$what" = all ] || [ "$what" = ncu ]; then
  small="--steps 1 --warmup 1 --units 400 --sketches 20000 --reads 4000000 --skip-cpu --skip-e2e --skip-dist5"
  # this library's kernels only (torch's data generation launches thousands of its own)
  mine='regex:scan_|select_|quirk_|tile_tmax|apply_runs|write_separators|dist_|dict_|list_gather|iota_|screen_|merge_bottom|DeviceRadixSort|DeviceScan|DeviceSegmented|DeviceSelect|reads_replay'
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$mine" -c 1500 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py $small > gpurun_out/${tag}_ncu_launches.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_kernel -s 2 -c 1 -o gpurun_out/${tag}_scan python bench.py --steps 1 --warmup 1 --units 400 --skip-dist --skip-cpu --skip-e2e > gpurun_out/${tag}_ncu_scan.log 2>&1
  dsmall="--steps 1 --warmup 1 --units 50 --sketches 100000 --skip-screen --skip-cpu --skip-e2e --skip-dist5"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:dist_probe_kernel -s 1 -c 1 -o gpurun_out/${tag}_probe python bench.py $dsmall > gpurun_out/${tag}_ncu_probe.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:dist_pair_kernel -s 4 -c 1 -o gpurun_out/${tag}_pair python bench.py $dsmall > gpurun_out/${tag}_ncu_pair.log 2>&1
  MASHGPU_DIST_PREFILTER=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:dist_kernel -s 1 -c 1 -o gpurun_out/${tag}_merge python bench.py --steps 1 --warmup 1 --units 50 --sketches 20000 --skip-screen --skip-cpu --skip-e2e --skip-dist5 > gpurun_out/${tag}_ncu_merge.log 2>&1
fi
ls -la gpurun_out | tail -30
