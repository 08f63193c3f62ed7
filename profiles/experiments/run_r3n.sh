# in-kernel split-K restricted to >= 64 k-blocks, identity tile prefetched by every split
timeout 240 python -m pytest tests/test_inference_gpu.py -x -q -m gpu 2>&1 | tail -2
for i in 1 2; do
timeout 120 python profiles/inference_latency.py 2>&1 | tail -1 | cut -c1-200
ARGUS_EVAL_SPLITK=0 timeout 120 python profiles/inference_latency.py 2>&1 | tail -1 | cut -c1-200
done
timeout 120 python profiles/profile_eval_detail.py 1 2>&1 | grep -E "total|K4608"
timeout 120 python profiles/profile_eval_detail.py 4 2>&1 | grep -E "total|K4608"
ARGUS_EVAL_SPLITK=0 timeout 120 python profiles/profile_eval_detail.py 4 2>&1 | grep -E "total|K4608"
