mkdir -p gpurun_out/r2k
SR_NO_BUILD=1 timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r2k/pytest.txt 2>&1; tail -3 gpurun_out/r2k/pytest.txt
python bench.py --steps 50 --warmup 5 > gpurun_out/r2k/default.json 2> gpurun_out/r2k/default.err
python bench.py --workload dtw --steps 10 > gpurun_out/r2k/dtw.json 2> gpurun_out/r2k/dtw.err
python bench.py --workload mfcc --steps 10 > gpurun_out/r2k/mfcc.json 2> gpurun_out/r2k/mfcc.err
python bench.py --batch 65536 --templates 200 --steps 10 --no-cpu --no-stream > gpurun_out/r2k/t200.json 2> gpurun_out/r2k/t200.err
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:"vad_kernel|mfcc_kernel|dtw_kernel" -s 9 -c 3 -o gpurun_out/r2k/prof_step python bench.py --steps 1 --warmup 3 --no-cpu --no-stream > gpurun_out/r2k/ncu_step.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2k/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-stream > gpurun_out/r2k/launch_run.log 2>&1
python - <<'PY'
import json
for f in ('default','dtw','mfcc','t200'):
    try:
        j=json.loads(open('gpurun_out/r2k/%s.json'%f).read().strip().splitlines()[-1]); e=j.get('e2e') or {}
        print(f,'value %.4g'%j['value'],'ms %.3f'%j['ms_per_step'],j.get('kernel_ms'),'e2e',e.get('ms_per_step'),e.get('transport'),j.get('parity_vs_cpu_sample'), (j.get('roofline') or {}).get('int_issue'))
    except Exception as ex: print(f,'ERR',ex)
PY
