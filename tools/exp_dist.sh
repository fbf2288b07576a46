# usage: bash tools/exp_dist.sh N tag   (multi-GPU tests + a short traced bench at N ranks)
N=${1:-2}; TAG=${2:-x}
timeout 600 python -m pytest tests/test_gpu_dist.py tests/test_gpu_cli.py -x -q -m gpu -k "dist or multi" 2>&1 | tail -3 > gpurun_out/${TAG}_tests.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N"
timeout 600 $TR --steps 2 --warmup 2 --no-e2e --no-next > gpurun_out/${TAG}_n$N.json 2> gpurun_out/${TAG}_n$N.err
H10X_TRACE=1 timeout 600 $TR --steps 1 --warmup 1 --no-e2e --no-next > /dev/null 2> gpurun_out/${TAG}_n${N}_trace.err
cat gpurun_out/${TAG}_tests.log
python - gpurun_out/${TAG}_n$N.json <<'P'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['stage_ms'].items()}, d['parity'].get('ok'))
P
grep "h10x-trace d" gpurun_out/${TAG}_n${N}_trace.err | tail -20
