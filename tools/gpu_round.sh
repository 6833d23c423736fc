set -x
for k in 2 3 4 6; do GGP_B200_UPLOAD_CHUNKS=$k python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_k$k.json 2>&1; done
python -c "
import json
for k in (2,3,4,6):
    j=json.load(open('gpurun_out/bench_k%d.json'%k)); print(k, j['ms_per_step'], j['e2e']['ms_per_step'], j['e2e']['value'])"
