mkdir -p gpurun_out/r2i
# N=1 transport checks on GPU 0
for cs in 1 2; do SR_COPY_STREAMS=$cs python bench.py --steps 20 --warmup 3 --no-cpu --no-stream > gpurun_out/r2i/n1_cs$cs.json 2> gpurun_out/r2i/n1_cs$cs.err; done
SR_PACK12=1 SR_COPY_STREAMS=2 python bench.py --steps 20 --warmup 3 --no-cpu --no-stream > gpurun_out/r2i/n1_pack_cs2.json 2> gpurun_out/r2i/n1_pack_cs2.err
SR_PACK12=1 SR_COPY_STREAMS=1 python bench.py --steps 20 --warmup 3 --no-cpu --no-stream > gpurun_out/r2i/n1_pack_cs1.json 2> gpurun_out/r2i/n1_pack_cs1.err
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
$T --nproc-per-node 8 --master-port 29911 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r2i/n8.json 2> gpurun_out/r2i/n8.err
$T --nproc-per-node 8 --master-port 29912 bench.py --gpus 8 --steps 20 --warmup 3 --no-cpu --no-stream --no-config3 > gpurun_out/r2i/n8b.json 2> gpurun_out/r2i/n8b.err
$T --nproc-per-node 4 --master-port 29913 bench.py --gpus 4 --steps 20 --warmup 3 --no-cpu --no-stream --no-config3 > gpurun_out/r2i/n4.json 2> gpurun_out/r2i/n4.err
$T --nproc-per-node 2 --master-port 29914 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu --no-stream --no-config3 > gpurun_out/r2i/n2.json 2> gpurun_out/r2i/n2.err
$T --nproc-per-node 8 --master-port 29915 bench.py --gpus 8 --steps 10 --warmup 3 --config 3 --no-cpu --no-stream > gpurun_out/r2i/n8_cfg3.json 2> gpurun_out/r2i/n8_cfg3.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2i/*.json')):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1]); e=j.get('e2e') or {}
        print(f.split('/')[-1],'value %.4g'%j['value'],'ms %.3f'%j['ms_per_step'],'e2e',e.get('ms_per_step'),e.get('transport',{}).get('chunks_packed_12bit'),j.get('allgather_matches_rank_results'),(j.get('config3') or {}).get('ms_per_step'),(j.get('config4_stream') or {}).get('value'))
    except Exception as ex: print(f,'ERR',ex, open(f.replace('.json','.err')).read()[-800:])
PY
