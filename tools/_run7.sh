mkdir -p gpurun_out/r2g
bash tools/box_topology.sh > gpurun_out/r2g/topo8.txt 2>&1
for w in 16 20 24; do SR_PACK12=0 SR_MFCC_WARPS=$w python bench.py --steps 20 --warmup 3 --no-cpu --no-stream > gpurun_out/r2g/w$w.json 2> gpurun_out/r2g/w$w.err; python -c "import json;j=json.loads(open('gpurun_out/r2g/w$w.json').read().strip().splitlines()[-1]);print('w$w',j['ms_per_step'],j['kernel_ms']['mfcc'])"; done
SR_NO_BUILD=1 timeout 900 python -m pytest tests -q -m gpu -k "nccl or multi or group or c_host" > gpurun_out/r2g/pytest_mgpu.txt 2>&1; tail -3 gpurun_out/r2g/pytest_mgpu.txt
stm32-speech-recognition_b200/host/spch_host_mgpu 8 1024 > gpurun_out/r2g/c_host_mgpu8.txt 2>&1; tail -1 gpurun_out/r2g/c_host_mgpu8.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r2g/n8.json 2> gpurun_out/r2g/n8.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29712 bench.py --gpus 4 --steps 20 --warmup 3 --no-cpu --no-stream --no-config3 > gpurun_out/r2g/n4.json 2> gpurun_out/r2g/n4.err
SR_PACK12=0 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29713 bench.py --gpus 8 --steps 20 --warmup 3 --no-cpu --no-stream --no-config3 > gpurun_out/r2g/n8_plain.json 2> gpurun_out/r2g/n8_plain.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2g/reference.json 2> gpurun_out/r2g/reference.err
python - <<'PY'
import json
for f in ('n8','n4','n8_plain','reference'):
    try:
        j=json.loads(open('gpurun_out/r2g/%s.json'%f).read().strip().splitlines()[-1])
        e=j.get('e2e',{})
        print(f, 'value %.4g'%j['value'], 'ms %.3f'%j['ms_per_step'], 'e2e_ms', e.get('ms_per_step'), e.get('transport'), e.get('numa'), j.get('allgather_matches_rank_results'), (j.get('config3') or {}).get('value'), (j.get('config3') or {}).get('ms_per_step'), (j.get('config4_stream') or {}).get('results',{}).get('chunk_800',{}).get('latency_ms_p50'), (j.get('config4_stream') or {}).get('results',{}).get('chunk_800',{}).get('latency_ms_p99'), (j.get('cpu_baseline') or {}).get('cores'))
    except Exception as e:
        print(f, 'ERR', e, open('gpurun_out/r2g/%s.err'%f).read()[-1200:])
PY
