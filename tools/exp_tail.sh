# usage: bash tools/exp_tail.sh TAG "ENV1=.. ENV2=.." "ENV.." ...   (one short 1 Gb bench per environment setting; '-' = none)
TAG=$1; shift
B="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-next --no-weak-base --no-clocks"
i=0
for envs in "$@"; do
  i=$((i+1)); f=gpurun_out/${TAG}_$i.json
  if [ "$envs" = "-" ]; then envs=""; fi
  env $envs $B > $f 2> gpurun_out/${TAG}_$i.err || tail -3 gpurun_out/${TAG}_$i.err
  python - $f "$envs" <<'P'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("%-40s %7.2f ms %s parity=%s" % (sys.argv[2] or "(default)", d['ms_per_step'], {k:round(v,1) for k,v in d['stage_ms'].items()}, d['parity'].get('ok')))
except Exception as e:
    print(sys.argv[2], "FAILED", e)
P
done
