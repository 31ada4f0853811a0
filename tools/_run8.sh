mkdir -p gpurun_out/r2h
SR_NO_BUILD=1 timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r2h/pytest.txt 2>&1; tail -3 gpurun_out/r2h/pytest.txt
python bench.py --steps 50 --warmup 5 > gpurun_out/r2h/default.json 2> gpurun_out/r2h/default.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29811 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu > gpurun_out/r2h/n2.json 2> gpurun_out/r2h/n2.err
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:"vad_kernel|mfcc_kernel|dtw_kernel" -s 9 -c 3 -o gpurun_out/r2h/prof_step python bench.py --steps 1 --warmup 3 --no-cpu --no-stream > gpurun_out/r2h/ncu_step.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2h/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-stream > gpurun_out/r2h/launch_run.log 2>&1
$NCU -k regex:"stream_step_kernel" -s 40 -c 1 -o gpurun_out/r2h/prof_stream python bench.py --workload stream --templates 20 > gpurun_out/r2h/ncu_stream.log 2>&1
python bench.py --workload mfcc --steps 10 > gpurun_out/r2h/mfcc.json 2> gpurun_out/r2h/mfcc.err
python bench.py --workload dtw --steps 10 > gpurun_out/r2h/dtw.json 2> gpurun_out/r2h/dtw.err
python bench.py --workload dtw_band --steps 5 > gpurun_out/r2h/dtw_band.json 2> gpurun_out/r2h/dtw_band.err
python bench.py --samples 16000 --batch 32768 --templates 80 --steps 20 --no-cpu --no-stream > gpurun_out/r2h/native2s.json 2> gpurun_out/r2h/native2s.err
compute-sanitizer --tool memcheck python -m pytest tests -q -m gpu -k "streaming_ragged or dtw_dynamic or geom_b or mfcc_fixed" -x > gpurun_out/r2h/sanitizer_memcheck.txt 2>&1; tail -3 gpurun_out/r2h/sanitizer_memcheck.txt
python - <<'PY'
import json
for f in ('default','n2','mfcc','dtw','dtw_band','native2s'):
    try:
        j=json.loads(open('gpurun_out/r2h/%s.json'%f).read().strip().splitlines()[-1]); e=j.get('e2e') or {}
        print(f,'value %.4g'%j['value'],'ms %.3f'%j['ms_per_step'],j.get('kernel_ms'),'e2e',e.get('ms_per_step'),e.get('transport'),j.get('allgather_matches_rank_results'),j.get('parity_vs_cpu_sample'))
    except Exception as ex: print(f,'ERR',ex)
PY
