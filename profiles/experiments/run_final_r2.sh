set -x
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/t_final_r2.log 2>&1; echo "rc=$?" >> gpurun_out/t_final_r2.log
tail -4 gpurun_out/t_final_r2.log
python bench.py --steps 20 --warmup 5 > gpurun_out/b_final_r2.json 2> gpurun_out/b_final_r2.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/b_final_r2_ref.json 2> gpurun_out/b_final_r2_ref.err
bash profiles/capture.sh r2f > gpurun_out/capture_r2f.log 2>&1
for f in conv wgrad; do
  ncu -i gpurun_out/prof_${f}_r2f.ncu-rep --page raw --csv > gpurun_out/ncu_raw_${f}_r2f.csv 2>/dev/null
  ncu -i gpurun_out/prof_${f}_r2f.ncu-rep --page details > gpurun_out/ncu_details_${f}_r2f.txt 2>/dev/null
done
ls -la gpurun_out/prof_*_r2f.ncu-rep; rm -f gpurun_out/prof_*_r2f.ncu-rep
du -sh gpurun_out
