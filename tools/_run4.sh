mkdir -p gpurun_out/r2d
nvidia-smi topo -m > gpurun_out/r2d/topo2.txt 2>&1
SR_NO_BUILD=1 timeout 1200 python -m pytest tests -q -m gpu -k "nccl or multi or group or torchrun or c_host" > gpurun_out/r2d/pytest_mgpu.txt 2>&1; tail -4 gpurun_out/r2d/pytest_mgpu.txt
stm32-speech-recognition_b200/host/spch_host_mgpu 2 2048 > gpurun_out/r2d/c_host_mgpu.txt 2>&1; cat gpurun_out/r2d/c_host_mgpu.txt
SR_PACK12=0 python bench.py --steps 20 --warmup 3 --no-cpu --no-stream > gpurun_out/r2d/n1_plain.json 2> gpurun_out/r2d/n1_plain.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2d/n2.json 2> gpurun_out/r2d/n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 10 --warmup 3 --config 3 --no-cpu --no-stream > gpurun_out/r2d/n2_cfg3.json 2> gpurun_out/r2d/n2_cfg3.err
python - <<'PY'
import json
for f in ('n1_plain','n2','n2_cfg3'):
    try:
        j=json.loads(open('gpurun_out/r2d/%s.json'%f).read().strip().splitlines()[-1])
        print(f, 'value %.4g'%j['value'], 'ms %.3f'%j['ms_per_step'], j['kernel_ms'], 'e2e', j['e2e']['ms_per_step'], j['e2e'].get('numa'), j.get('allgather_matches_rank_results'), (j.get('config3') or {}).get('ms_per_step'), (j.get('config4_stream') or {}).get('value'))
    except Exception as e:
        print(f, 'ERR', e, open('gpurun_out/r2d/%s.err'%f).read()[-1500:])
PY
