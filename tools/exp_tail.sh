set -x
B="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-next --no-weak-base --no-clocks"
$B > gpurun_out/r02t_base.json 2> gpurun_out/r02t_base.err
H10X_SR_THREADS=256 $B > gpurun_out/r02t_t256.json 2>/dev/null
H10X_SR_GROUP=2600 $B > gpurun_out/r02t_g2600.json 2>/dev/null
H10X_SR_GROUP=5120 $B > gpurun_out/r02t_g5120.json 2>/dev/null
H10X_SR_THREADS=256 H10X_SR_GROUP=2600 $B > gpurun_out/r02t_t256g2600.json 2>/dev/null
H10X_SR_THREADS=256 H10X_SR_GROUP=5120 $B > gpurun_out/r02t_t256g5120.json 2>/dev/null
for f in gpurun_out/r02t_*.json; do python - $f <<'P'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['stage_ms'].items()}, d['parity'].get('ok'))
P
done
