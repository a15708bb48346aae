#!/bin/bash
# sharded masked dictionary learning with the statistics exchange overlapped: NCCL tests + C4 step with / without
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
[ "$SKIP_TESTS" = 1 ] || { timeout 900 python -m pytest tests/test_dist.py -x -q -m gpu > gpurun_out/r2af_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2af_pytest.log; }
tail -3 gpurun_out/r2af_pytest.log
for ov in 1 0; do
DECOMP_DL_OVERLAP=$ov timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$ov bench.py --gpus $N --steps 20 --warmup 5 --legs configs,parity --skip cpu > gpurun_out/r2af_bench_ov$ov.json 2> gpurun_out/r2af_bench_ov$ov.err; echo "rc=$?"
python - <<PY
import json
b=json.loads(open('gpurun_out/r2af_bench_ov$ov.json').read().strip().splitlines()[-1])
print('N=$N overlap=$ov c4 ms', b['extra_configs']['c4_dl_masked_step']['ms_per_step'], 'parity', b['parity_multi_gpu']['pass'], b['parity_multi_gpu']['worst_error_over_ranks'], b.get('leg_errors'))
PY
done
