mkdir -p gpurun_out/r2l
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
$T --nproc-per-node 8 --master-port 29931 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r2l/n8.json 2> gpurun_out/r2l/n8.err
$T --nproc-per-node 8 --master-port 29932 bench.py --gpus 8 --steps 50 --warmup 5 --no-cpu --no-stream --no-config3 > gpurun_out/r2l/n8b.json 2> gpurun_out/r2l/n8b.err
$T --nproc-per-node 4 --master-port 29933 bench.py --gpus 4 --steps 20 --warmup 3 --no-cpu --no-stream --no-config3 > gpurun_out/r2l/n4.json 2> gpurun_out/r2l/n4.err
$T --nproc-per-node 2 --master-port 29934 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu --no-stream --no-config3 > gpurun_out/r2l/n2.json 2> gpurun_out/r2l/n2.err
python bench.py --steps 20 --warmup 3 --no-cpu --no-stream > gpurun_out/r2l/n1.json 2> gpurun_out/r2l/n1.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2l/*.json')):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1]); e=j.get('e2e') or {}
        print(f.split('/')[-1],'value %.4g'%j['value'],'ms %.3f'%j['ms_per_step'],'vad %.3f'%j['kernel_ms']['vad'],'e2e',e.get('ms_per_step'),e.get('transport',{}).get('chunks_packed_12bit'),j.get('allgather_matches_rank_results'),(j.get('config3') or {}).get('ms_per_step'),(j.get('config4_stream') or {}).get('value'))
    except Exception as ex: print(f,'ERR',ex, open(f.replace('.json','.err')).read()[-800:])
PY
