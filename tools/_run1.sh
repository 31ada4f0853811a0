mkdir -p gpurun_out/r2a
bash tools/box_topology.sh > gpurun_out/r2a/topo.txt 2>&1
SR_NO_BUILD=1 timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r2a/pytest.txt 2>&1; tail -5 gpurun_out/r2a/pytest.txt
python bench.py --steps 20 --warmup 3 > gpurun_out/r2a/default.json 2> gpurun_out/r2a/default.err
B="python bench.py --steps 20 --warmup 3 --no-cpu --no-stream"
SR_PACK12=0 $B > gpurun_out/r2a/plain.json 2> gpurun_out/r2a/plain.err
for nt in 6 8 10 12 14; do
SR_PACK12=1 SR_PACK_THREADS=$nt $B > gpurun_out/r2a/pack_t$nt.json 2> gpurun_out/r2a/pack_t$nt.err
done
SR_PACK12=1 SR_PACK_THREADS=10 SR_PACK_WC=1 $B > gpurun_out/r2a/pack_t10_wc.json 2>&1
SR_PACK12=1 SR_PACK_THREADS=10 SR_CHUNK_MB=16 $B > gpurun_out/r2a/pack_t10_c16.json 2>&1
SR_PACK12=1 SR_PACK_THREADS=10 SR_CHUNK_MB=64 $B > gpurun_out/r2a/pack_t10_c64.json 2>&1
python bench.py --workload dtw --steps 10 > gpurun_out/r2a/dtw.json 2> gpurun_out/r2a/dtw.err
python bench.py --workload mfcc --steps 10 > gpurun_out/r2a/mfcc.json 2> gpurun_out/r2a/mfcc.err
grep -h -o '"e2e": {[^}]*}[^}]*}' gpurun_out/r2a/*.json
grep -h -o '"kernel_ms": {[^}]*}' gpurun_out/r2a/default.json
python bench.py --workload dtw --steps 10 --dtw-variant 1 > gpurun_out/r2a/dtw_dyn.json 2> gpurun_out/r2a/dtw_dyn.err
python bench.py --steps 20 --warmup 3 --no-cpu --no-stream --dtw-variant 1 > gpurun_out/r2a/default_dyn.json 2> gpurun_out/r2a/default_dyn.err
grep -h -o '"kernel_ms": {[^}]*}' gpurun_out/r2a/default_dyn.json
grep -h -o '"value": [0-9.e+]*' gpurun_out/r2a/dtw.json gpurun_out/r2a/dtw_dyn.json
