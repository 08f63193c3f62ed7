F="--steps 4 --warmup 3 --no-cpu-baseline --no-inference --no-torch-baseline"
export ARGUS_BENCH_TRACE=2
for i in 1 2 3 4; do ARGUS_BENCH_TRACE_FILE=gpurun_out/btrace_$i.json python bench.py $F 2>&1 >/dev/null | grep -E "LOSS_TRACE" ; done
python - <<'PY'
import json
r=[json.load(open(f"gpurun_out/btrace_{i}.json")) for i in (1,2,3,4)]
names=r[0]["names"]
for j in (1,2,3):
    for st in range(len(r[0]["sig"])):
        a,b=r[0]["sig"][st],r[j]["sig"][st]
        if a!=b:
            d=[("in","grad","par")[k] for k in range(3) if a[k]!=b[k]]
            pd=[names[k-3] for k in range(3,len(a)) if a[k]!=b[k]]
            print(f"run {j+1} vs 1: first differing step {st}: {d}; {len(pd)} gradients differ; first {pd[:4]} last {pd[-4:]}")
            break
    else:
        print(f"run {j+1} == run 1")
PY
