N=$1; TAG=$2
H10X_TRACE=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 1 --warmup 1 --no-e2e --no-next > gpurun_out/${TAG}_n${N}_trace.json 2> gpurun_out/${TAG}_n${N}_trace.err
grep "h10x-trace" gpurun_out/${TAG}_n${N}_trace.err | tail -$((N*22)) | sort | awk '{k=$2; v=$(NF-1); s[k]+=v; if(v>m[k])m[k]=v; n[k]++} END{for(k in s) printf "%-14s mean %8.3f max %8.3f (n=%d)\n", k, s[k]/n[k], m[k], n[k]}' | sort
