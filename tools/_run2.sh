mkdir -p gpurun_out/r2b
SR_NO_BUILD=1 timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/r2b/pytest.txt 2>&1; tail -8 gpurun_out/r2b/pytest.txt
B="python bench.py --steps 20 --warmup 3 --no-cpu --no-stream"
for f in 0 1 2; do SR_PACK12=0 SR_MFCC_FILT=$f $B > gpurun_out/r2b/filt$f.json 2> gpurun_out/r2b/filt$f.err; done
for nt in 0 1; do for c in 8 16 32; do
SR_PACK_NT=$nt SR_CHUNK_MB=$c SR_PACK_THREADS=10 $B > gpurun_out/r2b/nt${nt}_c$c.json 2> gpurun_out/r2b/nt${nt}_c$c.err
done; done
SR_PACK_NT=0 SR_CHUNK_MB=16 SR_PACK_THREADS=6 $B > gpurun_out/r2b/nt0_c16_t6.json 2>&1
SR_PACK_NT=0 SR_CHUNK_MB=16 SR_PACK_THREADS=13 $B > gpurun_out/r2b/nt0_c16_t13.json 2>&1
python bench.py --workload stream --templates 20 > gpurun_out/r2b/stream.json 2> gpurun_out/r2b/stream.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2b/*.json')):
    try: j=json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e: print(f,'ERR'); continue
    e=j.get('e2e') or {}
    print(f.split('/')[-1], 'step %.3f'%j.get('ms_per_step',0), {k:round(v,3) for k,v in (j.get('kernel_ms') or {}).items() if k in ('vad','mfcc','dtw')}, 'e2e', e.get('ms_per_step'), (e.get('transport') or {}).get('chunks_packed_12bit'), (j.get('results') or {}).get('chunk_800',{}).get('latency_ms_p50'), (j.get('results') or {}).get('chunk_80',{}).get('latency_ms_p50'))
PY
