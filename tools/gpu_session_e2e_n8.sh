#!/bin/bash
# 8-GPU A/B of the sketch end-to-end path only (host -> sketches back), 4000 genomes per rank per step:
# NUMA binding of the ranks on/off, hybrid feed against ASCII copies only.  gpurun --gpus 8 -- 'bash tools/gpu_session_e2e_n8.sh r02'
tag=${1:-r02}
n=${2:-8}
mkdir -p gpurun_out
nproc > gpurun_out/${tag}_e2e_n${n}_host.txt; cat /sys/fs/cgroup/cpu.max >> gpurun_out/${tag}_e2e_n${n}_host.txt 2>/dev/null
lscpu | grep -E "NUMA|Model name|Socket" >> gpurun_out/${tag}_e2e_n${n}_host.txt 2>/dev/null
run() {
  name=$1; shift
  env "$@" timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n \
    --steps 2 --warmup 1 --units 4000 --skip-dist --skip-dist5 --skip-screen --skip-cpu > gpurun_out/${tag}_e2e_n${n}_${name}.json 2> gpurun_out/${tag}_e2e_n${n}_${name}.err
  echo "$name rc=$?"
}
run numa_hybrid MASHGPU_BENCH_NUMA=1
run nonuma_hybrid MASHGPU_BENCH_NUMA=0
run numa_ascii MASHGPU_BENCH_NUMA=1 MASHGPU_HOST_PACK=0
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/*_e2e_n*_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value %.1f e2e %.1f packed %.1f' % (d['value'], d['e2e']['value'], d['e2e'].get('packed_collection', {}).get('value', 0)), d['config'].get('host_binding'))
    except Exception as e:
        print(f, e)
PY
