mkdir -p gpurun_out/r2f
B="python bench.py --steps 20 --warmup 3 --no-cpu --no-stream"
run() { name=$1; shift; env SR_PACK12=0 "$@" $B > gpurun_out/r2f/$name.json 2> gpurun_out/r2f/$name.err; }
run w16 SR_MFCC_WARPS=16
run w16pre SR_MFCC_WARPS=16 SR_MFCC_PRE=1
run w20 SR_MFCC_WARPS=20
run w20pre SR_MFCC_WARPS=20 SR_MFCC_PRE=1
run w24 SR_MFCC_WARPS=24
for v in "SR_MFCC_WARPS=16 SR_MFCC_PRE=1" "SR_MFCC_WARPS=20 SR_MFCC_PRE=1" "SR_MFCC_WARPS=24"; do
  env SR_NO_BUILD=1 $v timeout 900 python -m pytest tests -q -m gpu -k "mfcc or recognise or streaming or enrol or reference_named or large_batch" > "gpurun_out/r2f/pytest_$(echo $v | tr ' =' '__').txt" 2>&1; echo "$v: $(tail -1 gpurun_out/r2f/pytest_$(echo $v | tr ' =' '__').txt)"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2f/w*.json')):
    try: j=json.loads(open(f).read().strip().splitlines()[-1]); print(f.split('/')[-1], j['ms_per_step'], {k:round(v,4) for k,v in j['kernel_ms'].items() if k in('vad','mfcc','dtw')})
    except Exception as e: print(f,'ERR',open(f.replace('.json','.err')).read()[-500:])
PY
