# usage: bash tools/exp_distbench.sh N tag [bench args]   (one bench run at N ranks, full line kept)
N=$1; TAG=$2; shift 2
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N "$@" > gpurun_out/${TAG}_n$N.json 2> gpurun_out/${TAG}_n$N.err
python - gpurun_out/${TAG}_n$N.json <<'P'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], "value %.3e" % d['value'], round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['stage_ms'].items()}, 'parity', d['parity'].get('ok'), 'e2e', (d.get('e2e') or {}).get('ms_per_step'), 'next', {k:(round(v,1) if isinstance(v,float) else v) for k,v in (d.get('next_rows') or {}).items() if k.endswith('_ms')})
P
