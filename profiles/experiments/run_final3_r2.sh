timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/t_final3_r2.log 2>&1; echo "rc=$?" >> gpurun_out/t_final3_r2.log
tail -3 gpurun_out/t_final3_r2.log
python bench.py --steps 20 --warmup 5 > gpurun_out/b_final3_r2.json 2> gpurun_out/b_final3_r2.err
python -c "
import json
d=json.loads(open('gpurun_out/b_final3_r2.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['e2e']['value'], d['final_loss'], d['clocks'], d['roofline']['frac'], d['roofline_aggregate']['frac'], d['inference']['batch1']['cuda_graph'], d['inference']['batch64']['cuda_graph'])"
ARGUS_PROFILE_DETAIL=1 python profiles/profile_detail.py > gpurun_out/detail_final3_r2.log 2>&1; head -3 gpurun_out/detail_final3_r2.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
