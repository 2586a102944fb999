#!/bin/bash
# One map sharded over 1 / 2 / 4 / 8 GPUs of ONE box (run under `gpurun --gpus 8`): parity of 4 and 8 shards against the oracle,
# then bench.py --parallelism sharded on the given configs.  usage: tools/gpu_scale.sh <tag> [configs...]
tag=$1; shift
cfgs=${@:-D}
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 300 python -m pytest tests/test_sharded.py -q -m gpu > gpurun_out/${tag}_sharded_pytest.log 2>&1; echo "sharded pytest rc=$?"; tail -4 gpurun_out/${tag}_sharded_pytest.log
for cfg in $cfgs; do
  for n in 1 2 4 8; do
    if [ $n = 1 ]; then
      timeout 150 python bench.py --config $cfg --steps 40 --warmup 10 --no-cpu-baseline > gpurun_out/${tag}_${cfg}_n1.json 2> gpurun_out/${tag}_${cfg}_n1.err
    else
      timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29400 + n)) bench.py --gpus $n --config $cfg --steps 40 --warmup 10 \
          --parallelism sharded --no-cpu-baseline > gpurun_out/${tag}_${cfg}_sharded${n}.json 2> gpurun_out/${tag}_${cfg}_sharded${n}.err
    fi
    echo "$cfg n=$n rc=$?"; tail -c 1200 gpurun_out/${tag}_${cfg}_*${n}.json | python -c "
import sys, json
for l in sys.stdin:
    l = l.strip()
    if l.startswith('{'):
        d = json.loads(l); print('   value %.1f e2e %.1f ms/step %.3f stage_ms %s' % (d['value'], d['e2e']['value'], d['ms_per_step'], {k: round(v, 3) for k, v in d['stage_ms'].items()}))
" 2>/dev/null
  done
done
