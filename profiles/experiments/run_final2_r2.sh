timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/t_final2_r2.log 2>&1; echo "rc=$?" >> gpurun_out/t_final2_r2.log
tail -3 gpurun_out/t_final2_r2.log
python bench.py --steps 20 --warmup 5 > gpurun_out/b_final2_r2.json 2> gpurun_out/b_final2_r2.err
for i in 1 2 3; do python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-inference --no-torch-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['final_loss'], d['clocks']['sm_mhz'])"; done
python -c "
import json
d=json.loads(open('gpurun_out/b_final2_r2.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['e2e']['value'], d['final_loss'], d['clocks'], d['inference']['batch1']['cuda_graph'], d['inference']['batch64']['cuda_graph'])"
